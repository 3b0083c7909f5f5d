"""Stub: tf.contrib is only imported (never executed) by the golden generator -- see ../__init__.py."""
