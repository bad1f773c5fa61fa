"""Empty stand-in (test infrastructure only)."""
FancyBboxPatch = None
