"""Alchemical scaling of n-body contributions (reference mbgdml/alchemy.py:30-62)."""
from .switching import linear_switching


class mbeAlchemyScale:  # pylint: disable=invalid-name
    """Scale the energies and forces of every ``order``-body fragment that contains
    ``entity_id`` by ``switching_func(data, **switching_kwargs)``.

    The device path recognises the reference's only shipped switching function,
    :func:`~mbgdml_b200.switching.linear_switching`, by identity; any other callable
    raises ``NotImplementedError`` at predict time (there is no CPU fallback).
    """

    def __init__(self, entity_id, order, switching_func, switching_kwargs):
        self.entity_id = entity_id
        self.order = order
        self.switching_func = switching_func
        self.switching_kwargs = switching_kwargs

    def scale(self, data, **kwargs):
        return self.switching_func(data, **self.switching_kwargs, **kwargs)

    def native_factor(self):
        """The linear factor the kernels apply (alchemy.py:52-62 + switching.py:30-48)."""
        if self.switching_func is not linear_switching:
            raise NotImplementedError(
                "only mbgdml_b200.switching.linear_switching is implemented on device "
                f"(got {self.switching_func!r})"
            )
        factor = self.switching_kwargs["factor"]
        assert 0 <= factor <= 1
        return float(factor)
