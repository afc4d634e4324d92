"""MP2 start amplitudes and energy (reference pymes/solver/mp2.py:9-22).

``solve`` keeps the reference signature (including the ``leve_shift`` spelling
and the swallowed ``**kwargs``).  T2 = V_abij / (e_i + e_j - e_a - e_b + shift) and
the two energy sums are single HBM-bound passes on the device."""
import numpy as np
import torch

from .. import backend as bk


def solve_device(eps_i, eps_a, V_ijab, V_abij, shift=0.0):
    """Device-resident variant: returns (energy: float, T2: cuda tensor)."""
    T2 = bk.mp2_amplitudes(eps_i, eps_a, shift, V_abij)
    scal = bk.zeros(8)
    bk.energy_doubles(T2, V_ijab, scal, mp2_form=True)
    s = scal.cpu().numpy()
    return float(s[0] + s[1]), T2


def solve(t_epsilon_i, t_epsilon_a, t_V_ijab, t_V_abij, leve_shift=0., **kwargs):
    want_numpy = not isinstance(t_V_abij, torch.Tensor)
    e, T2 = solve_device(bk.asdev(t_epsilon_i).contiguous(), bk.asdev(t_epsilon_a).contiguous(),
                         bk.asdev(t_V_ijab), bk.asdev(t_V_abij), leve_shift)
    return [e, bk.tonumpy(T2) if want_numpy else T2]
