"""Stand-in of the `logits_processor_zoo` package (not installed in this image; test infrastructure)."""
