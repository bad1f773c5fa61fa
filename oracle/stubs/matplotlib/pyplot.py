"""Empty stand-in (test infrastructure only)."""
