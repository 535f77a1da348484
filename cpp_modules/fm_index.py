"""`cpp_modules.fm_index` -- same module name and exports as the SWIG-generated module of the
reference (cpp_modules_linux/fm_index.i:7-21), backed by libfmidx.so."""
from fm_index_constrained_decoding_b200.fm_index import FMIndex, load_FMIndex, IntVector, IntVectorVector  # noqa: F401
