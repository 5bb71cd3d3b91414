// Stand-in for the reference's generated tools/version.h (test infrastructure).  The version strings only reach
// the JSON bookkeeping members of a spiral file, never the payload members the parity tests compare.
#pragma once
#define SPEC_VERSION "1.3.2-dev"
#define SEQSET_VERSION "2.0.0"
#define BIOGRAPH_VERSION "7.1.2-dev"
