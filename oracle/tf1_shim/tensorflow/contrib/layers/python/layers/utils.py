"""Stub for tensorflow.contrib.layers.python.layers.utils (import only)."""
