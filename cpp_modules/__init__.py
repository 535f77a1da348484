"""Drop-in package name of the reference's SWIG output directory (`cpp_modules/`, reference
README.md:27-31): `from cpp_modules.fm_index import FMIndex, load_FMIndex` (reference index.py:13-14)
resolves to the B200-native binding."""
