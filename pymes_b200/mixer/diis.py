"""Pulay DIIS on device-resident amplitudes.

Same interface and -- on purpose -- the same bookkeeping as the reference
``pymes.mixer.diis.DIIS`` (pymes/mixer/diis.py:9-112), including what happens
once the subspace is full (diis.py:59-62 carries one row/column too few of the
overlap matrix over, so the second-newest error vector keeps zero overlaps).
Iteration-by-iteration amplitude parity with the reference depends on it.

The O(m) passes over T2-sized data (one fused multi-dot for the new overlap row,
one linear combination) are HBM-bound kernels (``pmb_dots``/``pmb_lincomb``);
the (m+1)x(m+1) Lagrangian system is solved on the host with numpy like the
reference does.  In a sharded run each rank holds its (ab) slice of every
stored vector and the new overlap row is summed over ranks (``allreduce``).
"""
import numpy as np

from .. import backend as bk
from ..log import print_logging_info


class DIIS:
    def __init__(self, dim_space=5, allreduce=None):
        self.dim_space = dim_space
        self.L = np.zeros((1, 1))
        self.error_list = []
        self.amplitude_list = []
        self.allreduce = allreduce      # callable(np.ndarray) -> np.ndarray, or None

    def _overlap_row(self, sharded=None):
        """Re sum_nt <e_i[nt], e_new[nt]> for every stored i (diis.py:65-78).  Tensors flagged
        in ``sharded`` hold only this rank's rows: their partial dots are summed over ranks,
        replicated tensors are counted once."""
        n = len(self.error_list)
        row, row_sh = np.zeros(n), np.zeros(n)
        new = self.error_list[-1]
        for nt in range(len(new)):
            got = bk.dots([self.error_list[i][nt] for i in range(n)], new[nt]).cpu().numpy()
            if sharded is not None and sharded[nt]:
                row_sh += got
            else:
                row += got
        if self.allreduce is not None and sharded is not None and any(sharded):
            row_sh = self.allreduce(row_sh)
        return row + row_sh

    def mix(self, error, amplitude, sharded=None):
        """error / amplitude: lists of device tensors (e.g. [dT1, dT2], [T1, T2]).
        Returns the extrapolated amplitudes as new tensors."""
        error = [bk.asdev(e).contiguous() for e in error]
        amplitude = [bk.asdev(a).contiguous() for a in amplitude]
        full = len(self.error_list) == self.dim_space
        if full:
            self.error_list.pop(0)
            self.amplitude_list.pop(0)
        self.error_list.append(error)
        self.amplitude_list.append(amplitude)
        n = len(self.error_list)

        L = np.zeros((n + 1, n + 1))
        L[-1, :-1] = -1.0
        L[:-1, -1] = -1.0
        if full:
            L[:-3, :-3] = self.L[1:-2, 1:-2]       # sic: reference bookkeeping
        else:
            L[:-2, :-2] = self.L[:-1, :-1]
        L[:n, -2] += self._overlap_row(sharded)
        L[-2, :] = L[:, -2]
        self.L = L.copy()

        rhs = np.zeros(n + 1)
        rhs[-1] = -1.0
        w, U = np.linalg.eigh(self.L)
        if np.any(np.abs(w) < 1e-12):
            print_logging_info("Linear dependence found in DIIS subspace.", level=2)
            ok = np.abs(w) > 1e-12
            c = np.dot(U[:, ok] * (1.0 / w[ok]), np.dot(U[:, ok].T.conj(), rhs))
        else:
            c = np.linalg.inv(self.L).dot(rhs)
        self.last_coefficients = c

        mixed = [bk.lincomb(c[:n], [self.amplitude_list[a][i] for a in range(n)])
                 for i in range(len(amplitude))]
        print_logging_info("diis.mix", level=2)
        print_logging_info("Coefficients for combining amplitudes=", level=3)
        print_logging_info(c[:-1], level=3)
        print_logging_info("Sum of coefficients = {:.8f}".format(np.sum(c[:-1])), level=3)
        print_logging_info("Lagrangian multiplier = {:.8f}".format(c[-1]), level=3)
        return mixed
