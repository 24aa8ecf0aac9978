"""Import shim so the read-only reference can be imported for fixture generation (test infrastructure only)."""
