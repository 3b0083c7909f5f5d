"""Stub package (import only)."""
