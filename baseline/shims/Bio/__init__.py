"""Stand-in for biopython (absent from this image); golden generation only."""
