"""Switching functions (reference mbgdml/switching.py:30-48)."""


def linear_switching(data, factor):
    """y = factor * x with 0 <= factor <= 1 (switching.py:30-48)."""
    assert 0 <= factor <= 1
    return factor * data
