"""Closed-shell Hartree-Fock helpers (reference pymes/mean_field/hf.py:5-43).

These are O(no * nP^2) reads of V on small host arrays in every driver of the
reference; they accept numpy arrays or torch tensors and stay outside the
device hot path."""
import numpy as np


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def calc_hf_e(no, e_core, t_h_pq, t_V_pqrs):
    """E_HF = 2 sum_i h_ii + sum_ij (2 V_jiji - V_ijji) + E_core   (hf.py:5-11)."""
    h, occ = _np(t_h_pq), _np(t_V_pqrs[:no, :no, :no, :no])
    direct = 2.0 * np.einsum("jiji->", occ)
    exchange = -1.0 * np.einsum("ijji->", occ)
    return 2.0 * np.trace(h[:no, :no]) + (direct + exchange) + e_core


def construct_hf_matrix(no, t_h_pq, t_V_pqrs):
    """f_pq = h_pq + 2 V_piqi - V_piiq   (hf.py:14-18)."""
    f = np.array(_np(t_h_pq), dtype=np.float64, copy=True)
    f += 2.0 * np.einsum("piqi->pq", _np(t_V_pqrs[:, :no, :, :no]))
    f -= np.einsum("piiq->pq", _np(t_V_pqrs[:, :no, :no, :]))
    return f


def calcOccupiedOrbE(kinetic_G, tV_ijkl, no):
    """hf.py:21-30"""
    occ = _np(tV_ijkl)
    return _np(kinetic_G)[:no] + 2.0 * np.einsum("ijij->i", occ) - np.einsum("ijji->i", occ)


def calcVirtualOrbE(kinetic_G, t_V_aibj, t_V_aijb, no, nv):
    """hf.py:33-43"""
    return (_np(kinetic_G)[no:] + 2.0 * np.einsum("aiai->a", _np(t_V_aibj))
            - np.einsum("aiia->a", _np(t_V_aijb)))
