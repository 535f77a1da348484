"""Oracle-backed stand-in of the reference's SWIG package `cpp_modules` (test infrastructure)."""
