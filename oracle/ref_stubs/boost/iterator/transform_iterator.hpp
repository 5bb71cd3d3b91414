// Stand-in: modules/bio_base/seqset_merger.h includes this header without using it (test infrastructure).
#pragma once
