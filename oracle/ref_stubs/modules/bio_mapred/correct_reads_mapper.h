// Shadows modules/bio_mapred/correct_reads_mapper.h (test infrastructure): build_seqset/correct_reads.h includes
// the legacy map-reduce mapper only for the read types it pulls in.
#pragma once
#include "modules/bio_base/corrected_read.h"
#include "modules/bio_base/unaligned_read.h"
#include "modules/bio_mapred/correct_reads.h"
