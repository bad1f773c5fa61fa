"""Empty stand-in so that the reference's plotting imports resolve (test infrastructure only)."""
