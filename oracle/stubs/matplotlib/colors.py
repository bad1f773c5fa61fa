"""Empty stand-in (test infrastructure only)."""
BASE_COLORS = {}
