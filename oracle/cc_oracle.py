"""CPU oracle for the coupled-cluster amplitude-equation hot path.

TEST INFRASTRUCTURE ONLY.  This is a numpy restatement of the reference
algorithm (nickirk/pymes @ 734974a) and exists so that the CUDA path can be
checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package ``pymes_b200`` never does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
here against fixtures produced by importing the reference itself
(``tests/golden/make_golden.py``) and against the constants typed into the
reference's own tests (LiH/3-21G HF/CCD/CCSD, UEG 14e CCD/DCD, TC-UEG).

Every contraction is written as a row ``(coef, subscripts, operand names)`` in a
table; each table cites the reference lines it restates.  All arithmetic is
float64 ``numpy.einsum`` exactly like the reference (``optimize=True`` for the
multi-operand products, as in ``ccsd.py:11``).
"""
import string

import numpy as np

_OCC = "ijklmn"


def _es(sub, *ops):
    return np.einsum(sub.replace(" ", ""), *ops, optimize=True)


# ``pymes/solver/ccd.py`` calls bare ``np.einsum`` (no ``optimize``: numpy's single-threaded
# c_einsum loop), while ccsd.py / eom_ccsd.py use ``optimize=True`` (ccsd.py:11).  The values
# agree to round-off; the *timing* does not, so the CPU-baseline legs of bench.py switch the
# ccd.py rows to "as_written" to time what the reference really executes.
_MODE = {"ccd": "optimized"}


def set_ccd_einsum_mode(mode):
    assert mode in ("optimized", "as_written")
    _MODE["ccd"] = mode


def _es_ccd(sub, *ops):
    return np.einsum(sub.replace(" ", ""), *ops, optimize=(_MODE["ccd"] == "optimized"))


# --------------------------------------------------------------------------
# integral partition -- reference pymes/integral/partition.py:4-39
# --------------------------------------------------------------------------
BLOCK_KEYS = ("abci iabj iajk aijk klij aibj ijak abic iajb abcd iabc aijb "
              "ijka aibc ijab abij").split()


def partition(no, V):
    """16 zero-copy views of V_pqrs named by their index pattern."""
    def sl(ch):
        return slice(0, no) if ch in _OCC else slice(no, None)
    return {key: V[tuple(sl(ch) for ch in key)] for key in BLOCK_KEYS}


# --------------------------------------------------------------------------
# Hartree-Fock helpers -- reference pymes/mean_field/hf.py:5-18
# --------------------------------------------------------------------------
def hf_energy(no, e_core, h, V):
    occ = V[:no, :no, :no, :no]
    return (2.0 * np.trace(h[:no, :no]) + 2.0 * _es("jiji->", occ)
            - _es("ijji->", occ) + e_core)


def hf_fock(no, h, V):
    f = h.copy()
    f += 2.0 * _es("piqi->pq", V[:, :no, :, :no])
    f -= _es("piiq->pq", V[:, :no, :no, :])
    return f


# --------------------------------------------------------------------------
# denominators and MP2 -- reference pymes/solver/mp2.py:9-22, ccsd.py:152-156
# --------------------------------------------------------------------------
def denominators(eps_i, eps_a, shift=0.0):
    d2 = (eps_i[None, None, :, None] + eps_i[None, None, None, :]
          - eps_a[:, None, None, None] - eps_a[None, :, None, None])
    d1 = eps_i[None, :] - eps_a[:, None]
    return d1 + shift, d2 + shift


def mp2(eps_i, eps_a, V_ijab, V_abij, shift=0.0):
    _, d2 = denominators(eps_i, eps_a, shift)
    T2 = V_abij / d2
    e = 2.0 * _es("abij,ijab->", T2, V_ijab) - _es("abij,jiab->", T2, V_ijab)
    return e, T2


# --------------------------------------------------------------------------
# doubles residual -- reference pymes/solver/ccd.py:164-254
# --------------------------------------------------------------------------
def doubles_residual(no, fock, T2, V_klij, V_ijab, V_abij, V_iajb, V_iabj,
                     V_abcd, is_dcd=False, is_bruekner=False):
    """R_abij of the distinguishable-cluster / coupled-cluster doubles equation.

    Rows flagged ``ccd_only`` are dropped for DCD/DCSD (``ccd.py:179,189,218,237``).
    No hermiticity of V is assumed anywhere (``ccd.py:172,244-249``).
    """
    ccd = not is_dcd
    fab, fij = fock[no:, no:], fock[:no, :no]
    Tt = 2.0 * T2 - T2.transpose(1, 0, 2, 3)                    # ccd.py:199

    I = V_klij.copy()                                           # ccd.py:178-180
    if ccd:
        I = I + _es_ccd("klcd,cdij->klij", V_ijab, T2)
    R = V_abij + _es_ccd("klij,abkl->abij", I, T2)                  # ccd.py:185-186
    R = R + _es_ccd("abcd,cdij->abij", V_abcd, T2)                  # ccd.py:187
    if ccd:                                                     # ccd.py:189-191
        R = R + _es_ccd("alcj,cbil->abij", _es_ccd("klcd,adkj->alcj", V_ijab, T2), T2)
    R = R + _es_ccd("acik,cbkj->abij", Tt,
                _es_ccd("klcd,dblj->cbkj", V_ijab, Tt))             # ccd.py:202-204

    if is_bruekner:                                             # ccd.py:209-211
        # the reference binds VIEWS of the caller's Fock matrix here, so the in-place
        # updates of ccd.py:218-221 below modify t_fock_pq itself, sweep after sweep
        Xac, Xki = fab, fij
    else:                                                       # ccd.py:213-216
        Xac = fab - 0.5 * _es_ccd("adkl,lkdc->ac", Tt, V_ijab)
        Xki = fij + 0.5 * _es_ccd("cdil,lkdc->ki", Tt, V_ijab)
    if ccd:                                                     # ccd.py:218-221 (in place: `-=`, `+=`)
        Xac -= 0.5 * _es_ccd("adkl,lkdc->ac", Tt, V_ijab)
        Xki += 0.5 * _es_ccd("cdil,lkdc->ki", Tt, V_ijab)

    ops = {"Xac": Xac, "Xki": Xki, "T": T2, "Tt": Tt, "iajb": V_iajb,
           "iabj": V_iabj}
    ex_rows = [                                                 # ccd.py:231-235
        (+1.0, "ac,cbij->abij", "Xac", "T", False),
        (-1.0, "ki,abkj->abij", "Xki", "T", False),
        (-1.0, "kaic,cbkj->abij", "iajb", "T", False),
        (-1.0, "kbic,ackj->abij", "iajb", "T", False),
        (+1.0, "acik,kbcj->abij", "Tt", "iabj", False),
        (-1.0, "alci,cblj->abij", "Xp", "T", True),             # ccd.py:239
        (+1.0, "alci,bclj->abij", "Xp", "T", True),             # ccd.py:240
    ]
    if ccd:
        ops["Xp"] = _es_ccd("klcd,daki->alci", V_ijab, T2)          # ccd.py:238
    Ex = np.zeros_like(R)
    for coef, sub, a, b, ccd_only in ex_rows:
        if ccd_only and not ccd:
            continue
        Ex += coef * _es_ccd(sub, ops[a], ops[b])
    return R + Ex + Ex.transpose(1, 0, 3, 2)                    # ccd.py:249-252


def drccd_residual(eps_i, eps_a, T2, V_abij, V_aijb, V_iabj, V_ijab):
    """What ``drccd.get_residual`` (reference drccd.py:10-39) EXECUTES.  The einsum strings of
    drccd.py:34-35 are not the direct-ring equations of the comment above them: in
    ``"kbcj, acij -> abij"`` k is summed over V alone and j is a batch index, and in
    ``"acij, klcd, dblj -> abij"`` k is again summed over V alone with j a batch index.  The
    restatement spells those sums out."""
    R = V_abij + eps_a[:, None, None, None] * T2                # "ad,dbij->abij", diagonal f_ab
    R = R - eps_i[None, None, :, None] * T2                     # "ik,abkj->abij"
    Tp = T2.transpose(1, 0, 3, 2)                               # T[b,a,j,i]
    R = R + eps_a[None, :, None, None] * Tp                     # "bd,daji->abij"
    R = R - eps_i[None, None, None, :] * Tp                     # "jk,baki->abij"
    R = R + _es("akic,cbkj->abij", V_aijb, T2)                  # drccd.py:33
    Vs = V_iabj.sum(axis=0)                                     # [b,c,j]   drccd.py:34
    R = R + _es("bcj,acij->abij", Vs, T2)
    Ws = V_ijab.sum(axis=0)                                     # [l,c,d]   drccd.py:35
    R = R + _es("acij,lcd,dblj->abij", T2, Ws, T2)
    return R


def ccd_energy(T2, V_ijab):
    """(direct, exchange) -- reference ccd.py:256-262."""
    return (2.0 * _es_ccd("abij,ijab->", T2, V_ijab),
            -1.0 * _es_ccd("abij,ijba->", T2, V_ijab))


def ccsd_energy(f_ia, T1, T2, V_ijab):
    """(one-body, direct, exchange) -- reference ccsd.py:458-466."""
    tau = T2 + _es("ai,bj->abij", T1, T1)
    d, x = ccd_energy(tau, V_ijab)
    return 2.0 * _es("ia,ai->", f_ia, T1), d, x


# --------------------------------------------------------------------------
# T1 dressing -- reference pymes/solver/ccsd.py:226-421
# Each row: (coef, subscripts, source block, number of T1 factors)
# --------------------------------------------------------------------------
_FOCK_ROWS = {
    # target block : rows (coef, subscripts, operands); "t" = T1, f?? = fock blocks
    "ov": [(+2.0, "bj,jabi->ia", ("t", "iabj")),                # ccsd.py:257-258
           (-1.0, "bj,jiab->ia", ("t", "ijab"))],
    "vo": [(-1.0, "ji,aj->ai", ("foo", "t")),                   # ccsd.py:260-272
           (+1.0, "ab,bi->ai", ("fvv", "t")),
           (-1.0, "jb,bi,aj->ai", ("fov", "t", "t")),
           (+2.0, "bj,jabi->ai", ("t", "iabj")),
           (-2.0, "bj,jkbi,ak->ai", ("t", "ijak", "t")),
           (+2.0, "bj,jabc,ci->ai", ("t", "iabc", "t")),
           (-2.0, "bj,jkbc,ci,ak->ai", ("t", "ijab", "t", "t")),
           (-1.0, "bj,jaib->ai", ("t", "iajb")),
           (+1.0, "bj,jkib,ak->ai", ("t", "ijka", "t")),
           (-1.0, "bj,jacb,ci->ai", ("t", "iabc", "t")),
           (+1.0, "bj,jkcb,ci,ak->ai", ("t", "ijab", "t", "t"))],
    "oo": [(+2.0, "ck,kicj->ij", ("t", "ijak")),                # ccsd.py:275-279
           (-1.0, "ck,kijc->ij", ("t", "ijka")),
           (+1.0, "ib,bj->ij", ("fov", "t")),
           (+2.0, "ck,kicb,bj->ij", ("t", "ijab", "t")),
           (-1.0, "ck,kibc,bj->ij", ("t", "ijab", "t"))],
    "vv": [(+2.0, "ci,iacb->ab", ("t", "iabc")),                # ccsd.py:282-286
           (-1.0, "ci,iabc->ab", ("t", "iabc")),
           (-1.0, "ib,ai->ab", ("fov", "t")),
           (-2.0, "ck,klcb,al->ab", ("t", "ijab", "t")),
           (+1.0, "ck,kibc,ai->ab", ("t", "ijab", "t"))],
}


def dressed_fock(no, fock, T1, dV):
    """T1-dressed Fock matrix; every source is the undressed input."""
    src = dict(dV)
    src.update(t=T1, foo=fock[:no, :no], fvv=fock[no:, no:], fov=fock[:no, no:])
    out = fock.copy()
    where = {"ov": (slice(0, no), slice(no, None)),
             "vo": (slice(no, None), slice(0, no)),
             "oo": (slice(0, no), slice(0, no)),
             "vv": (slice(no, None), slice(no, None))}
    for blk, rows in _FOCK_ROWS.items():
        acc = np.zeros_like(out[where[blk]])
        for coef, sub, names in rows:
            acc += coef * _es(sub, *[src[n] for n in names])
        out[where[blk]] += acc
    return out


_V_ROWS = {
    "abij": [(-1.0, "kbij,ak->abij", "iajk", 1),                # ccsd.py:322-343
             (+1.0, "abcj,ci->abij", "abci", 1),
             (-1.0, "kbcj,ak,ci->abij", "iabj", 2),
             (-1.0, "alij,bl->abij", "aijk", 1),
             (+1.0, "klij,ak,bl->abij", "klij", 2),
             (-1.0, "alcj,ci,bl->abij", "aibj", 2),
             (+1.0, "klcj,ak,ci,bl->abij", "ijak", 3),
             (+1.0, "abid,dj->abij", "abic", 1),
             (-1.0, "kbid,ak,dj->abij", "iajb", 2),
             (+1.0, "abcd,ci,dj->abij", "abcd", 2),
             (-1.0, "kbcd,ak,ci,dj->abij", "iabc", 3),
             (-1.0, "alid,bl,dj->abij", "aijb", 2),
             (+1.0, "klid,ak,bl,dj->abij", "ijka", 3),
             (-1.0, "alcd,ci,bl,dj->abij", "aibc", 3),
             (+1.0, "klcd,ak,ci,bl,dj->abij", "ijab", 4)],
    "klij": [(+1.0, "klaj,ai->klij", "ijak", 1),                # ccsd.py:346-352
             (+1.0, "klib,bj->klij", "ijka", 1),
             (+1.0, "klab,ai,bj->klij", "ijab", 2)],
    "ijab": [],                                                 # ccsd.py:355-357
    "ijka": [(+1.0, "ijba,bk->ijka", "ijab", 1)],               # ccsd.py:359-362
    "ijak": [(+1.0, "ijab,bk->ijak", "ijab", 1)],               # ccsd.py:364-367
    "iajb": [(+1.0, "iacb,cj->iajb", "iabc", 1),                # ccsd.py:370-375
             (-1.0, "ikjb,ak->iajb", "ijka", 1),
             (-1.0, "ikcb,cj,ak->iajb", "ijab", 2)],
    "iabj": [(-1.0, "ikbj,ak->iabj", "ijak", 1),                # ccsd.py:378-383
             (+1.0, "iabc,cj->iabj", "iabc", 1),
             (-1.0, "ikbc,ak,cj->iabj", "ijab", 2)],
    "iabc": [(-1.0, "ijbc,aj->iabc", "ijab", 1)],               # ccsd.py:385-388
    "abic": [(-1.0, "jbic,aj->abic", "iajb", 1),                # ccsd.py:390-399
             (+1.0, "abdc,di->abic", "abcd", 1),
             (-1.0, "jbdc,aj,di->abic", "iabc", 2),
             (-1.0, "ajic,bj->abic", "aijb", 1),
             (+1.0, "kjic,ak,bj->abic", "ijka", 2),
             (-1.0, "ajdc,di,bj->abic", "aibc", 2),
             (+1.0, "kjdc,ak,di,bj->abic", "ijab", 3)],
    "iajk": [(-1.0, "iljk,al->iajk", "klij", 1),                # ccsd.py:402-411
             (+1.0, "iajb,bk->iajk", "iajb", 1),
             (-1.0, "iljb,al,bk->iajk", "ijka", 2),
             (+1.0, "iabk,bj->iajk", "iabj", 1),
             (-1.0, "ilbk,bj,al->iajk", "ijak", 2),
             (+1.0, "iabc,bj,ck->iajk", "iabc", 2),
             (-1.0, "ilbc,bj,al,ck->iajk", "ijab", 3)],
    "abcd": [(-1.0, "jbcd,aj->abcd", "iabc", 1),                # ccsd.py:414-419
             (-1.0, "aicd,bi->abcd", "aibc", 1),
             (+1.0, "jicd,aj,bi->abcd", "ijab", 2)],
}
DRESSED_KEYS = tuple(_V_ROWS)


def dressed_V(T1, dV, keys=None):
    """Dict of T1-dressed V blocks; keys not dressed by the reference stay None
    (``dict.fromkeys(dict_t_V, None)``, ccsd.py:316-317)."""
    out = dict.fromkeys(dV, None)
    for key in (DRESSED_KEYS if keys is None else keys):
        blk = dV[key].copy()
        for coef, sub, src, nt in _V_ROWS[key]:
            blk += coef * _es(sub, dV[src], *([T1] * nt))
        out[key] = blk
    return out


def singles_residual(no, fock_dressed, T1, T2, dV):
    """R_ai -- reference ccsd.py:423-438 (dressed Fock, undressed V)."""
    Tt = 2.0 * T2 - T2.transpose(0, 1, 3, 2)
    R = fock_dressed[no:, :no].copy()
    R += _es("jb,abij->ai", fock_dressed[:no, no:], Tt)
    R += _es("ajbc,bcij->ai", dV["aibc"], Tt)
    R -= _es("kjbc,ak,bcij->ai", dV["ijab"], T1, Tt)
    R -= _es("jkib,abjk->ai", dV["ijka"], Tt)
    R -= _es("jkcb,ci,abjk->ai", dV["ijab"], T1, Tt)
    return R


# --------------------------------------------------------------------------
# DIIS -- reference pymes/mixer/diis.py:9-112 (bookkeeping quirk included)
# --------------------------------------------------------------------------
class DIIS:
    def __init__(self, dim_space=5):
        self.dim_space = dim_space
        self.L = np.zeros((1, 1))
        self.errs, self.amps = [], []

    def mix(self, error, amplitude):
        full = len(self.errs) == self.dim_space
        if full:
            self.errs.pop(0)
            self.amps.pop(0)
        self.errs.append(error)
        self.amps.append(amplitude)
        n = len(self.errs)
        L = np.zeros((n + 1, n + 1))
        L[-1, :-1] = -1.0
        L[:-1, -1] = -1.0
        if full:
            # diis.py:59-60 -- one row/column too few is carried over; the
            # overlaps of the second-newest vector are left at zero.
            L[:-3, :-3] = self.L[1:-2, 1:-2]
        else:
            L[:-2, :-2] = self.L[:-1, :-1]
        for i in range(n):
            for e_i, e_new in zip(self.errs[i], self.errs[-1]):
                L[i, -2] += np.real(np.sum(e_i * e_new))        # diis.py:74-78
        L[-2, :] = L[:, -2]
        self.L = L.copy()
        rhs = np.zeros(n + 1)
        rhs[-1] = -1.0
        w, U = np.linalg.eigh(self.L)
        if np.any(np.abs(w) < 1e-12):                           # diis.py:87-93
            ok = np.abs(w) > 1e-12
            c = (U[:, ok] * (1.0 / w[ok])) @ (U[:, ok].T.conj() @ rhs)
        else:
            c = np.linalg.inv(self.L).dot(rhs)
        self.last_c = c
        out = [np.zeros_like(x) for x in self.amps[0]]
        for a in range(n):
            for i in range(len(out)):
                out[i] += self.amps[a][i] * c[a]
        return out


# --------------------------------------------------------------------------
# ground-state drivers -- reference ccd.py:24-162, ccsd.py:47-224
# --------------------------------------------------------------------------
def ccd_solve(no, fock, V, level_shift=0.0, is_dcd=False, is_diis=True,
              delta_e=1e-8, max_iter=50, amps=None, trace=None, is_dr_ccd=False, is_bruekner=False):
    """ccd.py:24-162.  ``is_dr_ccd`` / ``is_bruekner`` restate what those branches execute
    (ccd.py:95-98, 104-121), including the two things that make ``is_bruekner`` differ from
    its own comment: the denominator is the PRODUCT e_i e_j e_a e_b (``einsum('i,j,a,b->abij')``,
    ccd.py:118) and ``fock`` is modified in place by ``doubles_residual`` (its diagonal is what
    the first quasi-particle energies start from, because ``t_epsilon_*`` are views of it)."""
    dV = partition(no, V)
    eps_i, eps_a = fock.diagonal()[:no], fock.diagonal()[no:]     # views, as in ccd.py:40-41
    e_mp2, T2 = mp2(eps_i, eps_a, dV["ijab"], dV["abij"], level_shift)
    if amps is not None:
        T2 = amps
    _, d2 = denominators(eps_i, eps_a, level_shift)
    mixer = DIIS(6) if is_diis else None
    dE, e_last, e, it = abs(e_mp2), e_mp2, 0.0, 0
    while abs(dE) > delta_e and it <= max_iter:
        it += 1
        if is_dr_ccd:
            R = drccd_residual(eps_i, eps_a, T2, dV["abij"], dV["aijb"], dV["iabj"], dV["ijab"])
        else:
            R = doubles_residual(no, fock, T2, dV["klij"], dV["ijab"], dV["abij"],
                                 dV["iajb"], dV["iabj"], dV["abcd"], is_dcd, is_bruekner)
        if is_bruekner:                                             # ccd.py:104-119
            Tt = 2.0 * T2 - T2.transpose(1, 0, 2, 3)
            eps_i = eps_i + 0.5 * np.einsum("ilcd,cdil->i", dV["ijab"], Tt)
            eps_a = eps_a - 0.5 * np.einsum("klad,adkl->a", dV["ijab"], Tt)
            d2 = np.einsum("i,j,a,b->abij", eps_i, eps_i, -eps_a, -eps_a) + level_shift
        dT = R / d2
        T2 += dT
        if mixer is not None:
            T2 = mixer.mix([dT], [T2])[0]
        ed, ex = ccd_energy(T2, dV["ijab"])
        e = ed + ex
        dE, e_last = e - e_last, e
        if trace is not None:
            trace.append(dict(e=e, t2=T2.copy(), dt2_norm=np.linalg.norm(dT),
                              t2_norm=np.linalg.norm(T2)))
    return {"e": e, "t2": T2, "dE": dE, "iterations": it, "e_mp2": e_mp2,
            "hole e": eps_i, "particle e": eps_a}


def ccsd_sweep(no, fock, dV, T1, T2, d1, d2, mixer=None, is_dcsd=False):
    """One iteration of the CCSD loop body, reference ccsd.py:159-194."""
    ft = dressed_fock(no, fock, T1, dV)
    dVt = dressed_V(T1, dV)
    R1 = singles_residual(no, ft, T1, T2, dV)                   # ccsd.py:167-168
    R2 = doubles_residual(no, ft, T2, dVt["klij"], dVt["ijab"], dVt["abij"],
                          dVt["iajb"], dVt["iabj"], dVt["abcd"], is_dcsd)
    dT1, dT2 = R1 / d1, R2 / d2
    T1 = T1 + dT1
    T2 = T2 + dT2
    if mixer is not None:
        T1, T2 = mixer.mix([dT1, dT2], [T1, T2])
    return T1, T2, ccsd_energy(fock[:no, no:], T1, T2, dV["ijab"]), dT2


def ccsd_solve(no, fock, V, level_shift=0.0, is_dcsd=False, is_diis=True,
               delta_e=1e-8, max_iter=50, amps=None, trace=None):
    nv = fock.shape[0] - no
    dV = partition(no, V)
    eps_i, eps_a = fock.diagonal()[:no].copy(), fock.diagonal()[no:].copy()
    e_mp2, T2 = mp2(eps_i, eps_a, dV["ijab"], dV["abij"], level_shift)
    T1 = np.zeros((nv, no))
    if amps is not None:
        T1, T2 = amps
    d1, d2 = denominators(eps_i, eps_a, level_shift)
    mixer = DIIS(6) if is_diis else None
    dE, e_last, e, it = abs(e_mp2), e_mp2, 0.0, 0
    while abs(dE) > delta_e and it <= max_iter:
        it += 1
        T1, T2, (e1, ed, ex), dT2 = ccsd_sweep(no, fock, dV, T1, T2, d1, d2, mixer, is_dcsd)
        e = e1 + ed + ex
        dE, e_last = e - e_last, e
        if trace is not None:
            trace.append(dict(e=e, t1=T1.copy(), t2=T2.copy(),
                              dt2_norm=np.linalg.norm(dT2),
                              t2_norm=np.linalg.norm(T2)))
    return {"e": e, "t1": T1, "t2": T2, "dE": dE, "iterations": it,
            "e_mp2": e_mp2}


# --------------------------------------------------------------------------
# EOM-EE-CCSD sigma, diagonals and Davidson -- reference pymes/solver/eom_ccsd.py
# Term tables: (coefficient, einsum string, operand names); fov/foo/fvv = Fock blocks,
# T = ground-state T2, u1/u2 = trial vector, other names = integral-dictionary keys.
# --------------------------------------------------------------------------
SIGMA1_TERMS = [                                                       # eom_ccsd.py:288-308
    (+2., "jb,baji->ai", "fov u2"), (-1., "ji,aj->ai", "foo u1"),
    (-1., "jb,abji->ai", "fov u2"), (+1., "ab,bi->ai", "fvv u1"),
    (+2., "jabi,bj->ai", "iabj u1"), (-1., "jaib,bj->ai", "iajb u1"),
    (-2., "jkib,abjk->ai", "ijka u2"), (+2., "jabc,bcji->ai", "iabc u2"),
    (+1., "jkib,bajk->ai", "ijka u2"), (-1., "jacb,bcji->ai", "iabc u2"),
    (+4., "jkbc,baji,ck->ai", "ijab T u1"), (-2., "jkbc,bajk,ci->ai", "ijab T u1"),
    (-2., "jkbc,bcji,ak->ai", "ijab T u1"), (-2., "jkbc,abji,ck->ai", "ijab T u1"),
    (-2., "jkcb,baji,ck->ai", "ijab T u1"), (+1., "jkbc,abjk,ci->ai", "ijab T u1"),
    (+1., "jkcb,bcji,ak->ai", "ijab T u1"), (+1., "jkcb,abji,ck->ai", "ijab T u1"),
]

SIGMA2_P_TERMS = [                                                     # eom_ccsd.py:332-373
    (-2., "klid,abkj,dl->abij", "ijka T u1"), (-2., "klci,cbkj,al->abij", "ijak T u1"),
    (+2., "kacd,cbkj,di->abij", "iabc T u1"), (+2., "ladc,cbij,dl->abij", "iabc T u1"),
    (-1., "kd,abkj,di->abij", "fov T u1"), (-1., "lc,cbij,al->abij", "fov T u1"),
    (+1., "klid,abkl,dj->abij", "ijka T u1"), (+1., "klic,cbkj,al->abij", "ijka T u1"),
    (+1., "klid,adkj,bl->abij", "ijka T u1"), (-1., "kbij,ak->abij", "iajk u1"),
    (+1., "kldi,bdkj,al->abij", "ijak T u1"), (-1., "kacd,bckj,di->abij", "iabc T u1"),
    (+1., "kldi,abkj,dl->abij", "ijak T u1"), (-1., "kadc,cbkj,di->abij", "iabc T u1"),
    (-1., "kadc,bcki,dj->abij", "iabc T u1"), (-1., "lacd,cdji,bl->abij", "iabc T u1"),
    (-1., "lacd,cbij,dl->abij", "iabc T u1"), (+1., "abic,cj->abij", "abic u1"),
    (+4., "klcd,caki,dblj->abij", "ijab T u2"), (-2., "klcd,cakl,dbij->abij", "ijab T u2"),
    (-2., "klcd,cdki,ablj->abij", "ijab T u2"), (-2., "klcd,caki,bdlj->abij", "ijab T u2"),
    (+2., "kaci,cbkj->abij", "iabj u2"), (-2., "klcd,acki,dblj->abij", "ijab T u2"),
    (-2., "kldc,caki,dblj->abij", "ijab T u2"), (-2., "kldc,abkj,dcil->abij", "ijab T u2"),
    (-2., "lkcd,cbij,adlk->abij", "ijab T u2"), (-1., "ki,abkj->abij", "foo u2"),
    (+1., "ac,cbij->abij", "fvv u2"), (-1., "kaic,cbkj->abij", "iajb u2"),
    (-1., "kbic,ackj->abij", "iajb u2"), (+1., "klcd,ackl,dbij->abij", "ijab T u2"),
    (+1., "kldc,cdki,ablj->abij", "ijab T u2"), (+1., "klcd,acki,bdlj->abij", "ijab T u2"),
    (-1., "kaci,bckj->abij", "iabj u2"), (+1., "kldc,acki,dblj->abij", "ijab T u2"),
    (+1., "kldc,abkj,dcli->abij", "ijab T u2"), (+1., "kldc,caki,dbjl->abij", "ijab T u2"),
    (+1., "kldc,ackj,dbil->abij", "ijab T u2"), (+1., "lkcd,cbij,dalk->abij", "ijab T u2"),
]

SIGMA2_NP_TERMS = [                                                    # eom_ccsd.py:380-383
    (+1., "klij,abkl->abij", "klij u2"), (+1., "kldc,abkl,dcij->abij", "ijab T u2"),
    (+1., "lkcd,cdij,ablk->abij", "ijab T u2"), (+1., "abcd,cdij->abij", "abcd u2"),
]


def _eom_ops(no, fock, dV, u1, u2, T2):
    ops = dict(dV)
    ops.update(foo=fock[:no, :no], fov=fock[:no, no:], fvv=fock[no:, no:], T=T2, u1=u1, u2=u2)
    return ops


def eom_sigma_singles(no, fock, dV, u1, u2, T2):
    """eom_ccsd.py:268-310."""
    ops = _eom_ops(no, fock, dV, u1, u2, T2)
    out = np.zeros(u1.shape, dtype=np.result_type(u1, u2))
    for coef, spec, names in SIGMA1_TERMS:
        out += coef * _es(spec, *[ops[n] for n in names.split()])
    return out


def eom_sigma_doubles(no, fock, dV, u1, u2, T2):
    """eom_ccsd.py:312-385 (explicit + baji permutation at line 377)."""
    ops = _eom_ops(no, fock, dV, u1, u2, T2)
    out = np.zeros(u2.shape, dtype=np.result_type(u1, u2))
    for coef, spec, names in SIGMA2_P_TERMS:
        out += coef * _es(spec, *[ops[n] for n in names.split()])
    out = out + out.transpose(1, 0, 3, 2)
    for coef, spec, names in SIGMA2_NP_TERMS:
        out += coef * _es(spec, *[ops[n] for n in names.split()])
    return out


def eom_diag_singles(no, fock, dV, T2):
    """eom_ccsd.py:169-198."""
    V, d = dV["ijab"], None
    d = -fock[:no, :no].diagonal()[None, :] + fock[no:, no:].diagonal()[:, None]
    d = d + 2. * _es("iaai->ai", dV["iabj"]) - _es("iaia->ai", dV["iajb"])
    d = d + 4. * _es("jiba,baji->ai", V, T2)
    d = d - 2. * _es("jkba,abjk->a", V, T2)[:, None] - 2. * _es("jicb,bcji->i", V, T2)[None, :]
    d = d - 2. * _es("jiba,abji->ai", V, T2) - 2. * _es("jiab,baji->ai", V, T2)
    d = d + _es("jkab,abjk->a", V, T2)[:, None] + _es("jicb,bcji->i", V, T2)[None, :]
    return d + _es("jiab,abji->ai", V, T2)


def eom_diag_doubles(no, fock, dV, T2):
    """eom_ccsd.py:200-266 (including the [a,i] placement of the "ibib->bi" term, line 231)."""
    V, nv = dV["ijab"], T2.shape[0]
    ai = lambda x: x[:, None, :, None]
    d = np.zeros((nv, nv, no, no))
    d += ai(4. * _es("kica,caki->ai", V, T2))
    d += -2. * _es("klca,cakl->a", V, T2)[:, None, None, None]
    d += -2. * _es("kicd,cdki->i", V, T2)[None, None, :, None]
    d += ai(-2. * _es("kica,caki->ai", V, T2))
    d += ai(2. * _es("iaai->ai", dV["iabj"]))
    d += ai(-2. * _es("kica,acki->ai", V, T2))
    d += ai(-2. * _es("kiac,caki->ai", V, T2))
    d += -2. * _es("kjab,abkj->abj", V, T2)[:, :, None, :]
    d += -2. * _es("ijcb,cbij->ij", V, T2)[None, None, :, :]
    d += -fock[:no, :no].diagonal()[None, None, :, None] + fock[no:, no:].diagonal()[:, None, None, None]
    d += ai(-1. * _es("iaia->ai", dV["iajb"]))
    d += ai(-1. * _es("ibib->bi", dV["iajb"]))
    d += _es("klca,ackl->a", V, T2)[:, None, None, None]
    d += _es("kidc,cdki->i", V, T2)[None, None, :, None]
    d += ai(_es("kicb,acki->ai", V, T2))
    d += ai(-1. * _es("iaai->ai", dV["iabj"]))
    d += ai(_es("kiac,acki->ai", V, T2))
    d += _es("kiab,abkj->abij", V, T2)
    d += _es("kjac,caki->aij", V, T2)[:, None, :, :]
    d += _es("kjac,ackj->aj", V, T2)[:, None, None, :]
    d += _es("ijca,cbij->abij", V, T2)
    d = d + d.transpose(1, 0, 3, 2)
    d += _es("ijij->ij", dV["klij"])[None, None, :, :]
    d += _es("klab,abkl->ab", V, T2)[:, :, None, None]
    d += _es("ijcd,cdij->ij", V, T2)[None, None, :, :]
    d += _es("abab->ab", dV["abcd"])[:, :, None, None]
    return d


def eom_davidson(no, fock, dV, T2, n_excit=3, max_iter=500, e_epsilon=1e-8):
    """Block Davidson of eom_ccsd.py:46-167: unit guesses on the smallest eps_a - eps_i, dense QR
    of the whole subspace every sweep, sigma for every vector, plain (non-conjugated) projections,
    scalar preconditioner e_n - D_guess + 1e-5, collapse at 4 n_excit vectors."""
    nv = T2.shape[0]
    eps = fock.diagonal()
    D_ai = -(eps[:no][None, :] - eps[no:][:, None]).ravel()
    guess = np.argsort(D_ai)[:n_excit]
    n1 = nv * no
    U = np.zeros((n1 + T2.size, n_excit))
    U[guess, np.arange(n_excit)] = 1.0
    e_excit = np.zeros(n_excit)
    diff = np.inf
    for it in range(max_iter):
        U, _ = np.linalg.qr(U)                                  # eom_ccsd.py:512-541
        m = U.shape[1]
        W = np.empty_like(U)
        for l in range(m):
            u1, u2 = U[:n1, l].reshape(nv, no), U[n1:, l].reshape(T2.shape)
            W[:n1, l] = eom_sigma_singles(no, fock, dV, u1, u2, T2).ravel()
            W[n1:, l] = eom_sigma_doubles(no, fock, dV, u1, u2, T2).ravel()
        B = U.T @ W                                             # eom_ccsd.py:103-109
        ev, v = np.linalg.eig(B)
        low = ev.argsort()[:n_excit]
        e, v = np.real(ev[low]), np.real(v[:, low])
        if m >= 4 * n_excit:                                    # eom_ccsd.py:122-133
            U = U @ v
        else:                                                   # eom_ccsd.py:134-147
            Y = (W @ v - (U @ v) * e[None, :]) / (e - D_ai[guess] + 1e-5)[None, :]
            U = np.concatenate([U, Y], axis=1)
            diff = np.linalg.norm(e_excit - e)
            e_excit = e
        if diff < e_epsilon:
            break
    return e_excit, it + 1


def flops_doubles_residual(no, nv, is_dcd=False):
    """Algorithmic flop count of one residual (SURVEY 8d / BASELINE.md 3)."""
    o, v = float(no), float(nv)
    if is_dcd:
        return 2*o**2*v**4 + 10*o**3*v**3 + 2*o**4*v**2 + 4*o**2*v**3 + 4*o**3*v**2
    return 2*o**2*v**4 + 20*o**3*v**3 + 4*o**4*v**2 + 4*o**2*v**3 + 4*o**3*v**2
