"""Synthetic non-hermitian two-body integrals for BASELINE.json configs[2] ("synthetic FCIDUMP,
o = 50, v = 500, random permutation-symmetric real integrals, non-hermitian CCSD"; recipe of
SURVEY 8(d)):  V[p,q,r,s] = eps * N(0,1) drawn from a COUNTER-BASED generator keyed on the
canonical representative of the only symmetry a transcorrelated Hamiltonian keeps,
(p,q,r,s) ~ (q,p,s,r) (pymes/util/fcidump.py:147-149), so that any sub-block -- in particular one
rank's row block of V_abcd, 500 GB / N at full size -- can be generated on its own, on any
device, without V_pqrs ever existing.

The generator is a pure function of (seed, canonical index): two rounds of the splitmix64
finaliser give 64 bits, the top 16 select one of 65536 standard-normal quantiles (a table made
once with ``scipy.special.ndtri``).  Integer arithmetic, one look-up and one multiply -- no
transcendental function is evaluated per element, so the numpy code below (host) and the kernel
``pmb_synth_block`` (device, csrc/synth_build.cu) produce the same doubles BIT FOR BIT."""
import ctypes as C

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
TABLE_BITS = 16
_table = {}


def _mix(x):
    """splitmix64 finaliser on a uint64 array (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * _M1
        x = (x ^ (x >> np.uint64(27))) * _M2
        return x ^ (x >> np.uint64(31))


def normal_table():
    """The 65536 quantiles z_k = Phi^-1((k + 1/2) / 65536) (mean 0, variance 1 - 3e-4)."""
    if "host" not in _table:
        from scipy import special
        k = np.arange(1 << TABLE_BITS, dtype=np.float64)
        _table["host"] = np.ascontiguousarray(special.ndtri((k + 0.5) / float(1 << TABLE_BITS)))
    return _table["host"]


def normal_from_index(seed, idx):
    """N(0,1) deviates for an array of uint64 counters ``idx`` (deterministic in (seed, idx))."""
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _mix(_mix(idx * _GOLD + np.uint64(seed) * _M1 + np.uint64(1)))
    return normal_table()[(h >> np.uint64(64 - TABLE_BITS)).astype(np.int64)]


def tc_block_device(n_orb, lo, ext, seed=0, eps=None, out=None):
    """The same block as :func:`tc_block`, written by ``pmb_synth_block`` into a CUDA tensor
    (HBM-bound; one rank's 62.5 GB row block of V_abcd at o = 50, v = 500 takes tens of ms)."""
    import torch
    from .. import _lib, backend as bk
    eps = 1e-2 / n_orb if eps is None else eps
    key = ("dev", bk._device_key()[0])
    if key not in _table:
        _table[key] = torch.from_numpy(normal_table()).to(bk.device())
    if out is None:
        out = bk.empty(*[int(e) for e in ext])
    elif not out.is_contiguous() or tuple(out.shape) != tuple(int(e) for e in ext):
        raise ValueError("out must be a contiguous tensor of the block's shape")
    _lib.check(_lib.load().pmb_synth_block(int(n_orb), C.c_ulonglong(int(seed)), float(eps), bk._ptr(_table[key]),
                                           _lib.I32x4(*[int(x) for x in lo]), _lib.I32x4(*[int(x) for x in ext]),
                                           bk._ptr(out), bk._stream()), "pmb_synth_block")
    return out


def tc_block(n_orb, lo, ext, seed=0, eps=None):
    """Block V[lo0:lo0+ext0, ..., lo3:lo3+ext3] of the synthetic tensor (numpy, C order).
    ``eps`` defaults to 1e-2 / n_orb (small enough for CC to converge)."""
    eps = 1e-2 / n_orb if eps is None else eps
    n = np.uint64(n_orb)
    p, q, r, s = (np.arange(l, l + e, dtype=np.uint64).reshape([-1 if d == k else 1 for d in range(4)])
                  for k, (l, e) in enumerate(zip(lo, ext)))
    with np.errstate(over="ignore"):
        i1 = ((p * n + q) * n + r) * n + s
        i2 = ((q * n + p) * n + s) * n + r
    return eps * normal_from_index(seed, np.minimum(i1, i2))


def tc_integrals(n_orb, seed=0, eps=None):
    """The dense V_pqrs [n_orb]^4 (small systems / tests)."""
    return tc_block(n_orb, (0, 0, 0, 0), (n_orb,) * 4, seed, eps)


def tc_fock(no, nv, seed=0, off_diagonal=1e-3):
    """diag(-2..-1 | +1..+3) plus small NON-symmetric off-diagonal elements."""
    n = no + nv
    f = np.diag(np.concatenate([np.linspace(-2.0, -1.0, no), np.linspace(1.0, 3.0, nv)]))
    idx = np.arange(n * n, dtype=np.uint64).reshape(n, n) + np.uint64(1) << np.uint64(40)
    off = off_diagonal * normal_from_index(seed, idx)
    off[np.diag_indices(n)] = 0.0
    return f + off


def tc_blocks(no, nv, keys, seed=0, eps=None, ranges=None, device=False):
    """Named partition blocks (``integral.partition`` keys) generated directly; ``ranges`` =
    ``{key: {dim: (lo, n)}}`` restricts a dimension to absolute orbital indices (a rank's rows).
    ``device=True``: CUDA tensors written by ``pmb_synth_block`` (bit-identical values)."""
    from ..integral.partition import OCCUPIED
    out, geom = {}, {}
    for key in keys:
        lo = [0 if ch in OCCUPIED else no for ch in key]
        ext = [no if ch in OCCUPIED else nv for ch in key]
        for dim, (r_lo, r_n) in (ranges or {}).get(key, {}).items():
            lo[dim], ext[dim] = int(r_lo), int(r_n)
        geom[key] = (lo, ext)
    pre = {}
    if device and "iabc" in geom and "aibc" in geom and \
            int(np.prod(geom["iabc"][1])) == int(np.prod(geom["aibc"][1])):
        from .. import backend as bk          # back to back: one operand for solver.ccsd.pair_with_tau
        pre["iabc"], pre["aibc"] = bk.empty_stacked(geom["iabc"][1], geom["aibc"][1])
    for key in keys:
        lo, ext = geom[key]
        if device:
            out[key] = tc_block_device(no + nv, lo, ext, seed, eps, out=pre.get(key))
        else:
            out[key] = tc_block(no + nv, lo, ext, seed, eps)
    return out
