"""Name-only stand-in for Biopython (see ../README.md)."""
