"""Small helpers used by the flow modules."""
