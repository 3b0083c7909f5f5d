"""Stub for tensorflow.contrib.layers.python.layers.utils (TEST INFRASTRUCTURE)."""


def convert_collection_to_dict(collection):
    return {}
