"""Plane-wave basis function (reference pymes/basis_set/planewave.py:3-26)."""
import numpy as np


class BasisFunc:
    """Plane wave with wavevector 2 pi (i, j, k + k_shift) / L and spin +-1.

    ``kinetic`` is evaluated with the reference's expression so that shell membership
    at the cutoff and the stable sort inside degenerate shells come out identical."""
    __slots__ = ("k", "L", "kp", "kinetic", "spin")

    def __init__(self, i, j, k, L, spin, k_shift=(0., 0., 0.)):
        if spin not in (-1, 1):
            raise RuntimeError("spin not +1 or -1")
        self.k = np.array([i, j, k])
        self.L = L
        self.kp = (self.k + k_shift) * 2 * np.pi / L
        self.kinetic = np.dot(self.kp, self.kp) / 2.
        self.spin = spin

    def __repr__(self):
        return repr((self.k, self.kinetic, self.spin))

    def __lt__(self, other):
        return self.kinetic < other.kinetic
