"""Stub of `ray` (absent from this image) so the unmodified reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/reference.py).  Only the serial
`use_ray=False` path of the reference is exercised; the few entry points the
reference touches at import / construction time (mbe.py:28,673-693) are
identity functions.
"""


def is_initialized():
    return False


def init(*args, **kwargs):
    return None


def put(obj):
    return obj


def get(obj):
    return obj


def remote(func):
    return func


def wait(workers):
    return workers[:1], workers[1:]
