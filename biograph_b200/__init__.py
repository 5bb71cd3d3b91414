"""B200-native seqset construction for BioGraph (k-mer count -> correct -> seqset).

The product is the CUDA library `libbgx.so` behind the C ABI in include/bgx.h; this package is
the thin Python binding used by the tests and bench.py.  There is no CPU fallback."""
from .bgx import Bgx, BgxError, Options, assemble_seqset, lib_path, load_library, build_library  # noqa: F401
