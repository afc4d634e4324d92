"""CPU oracle for the UEG plane-wave basis and momentum-conserving integrals.

TEST INFRASTRUCTURE ONLY (same rules as ``oracle/cc_oracle.py``): a vectorised numpy
restatement of ``pymes/model/ueg.py`` (nickirk/pymes @ 734974a) used to check
``pmb_ueg_umat`` / ``pmb_ueg_pair_tables`` / ``pmb_ueg_build_block`` and to feed the CPU
baseline of ``bench.py`` with the same Hamiltonian the GPU arm uses.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks the basis, the index
map, Coulomb / TC ``is_only_2b`` / TC ``is_effect_2b`` integrals and ``u_mat`` against
``tests/golden/ueg_coulomb.npz`` and ``ueg_tc.npz``, which were produced by the
reference's own triple loop (``tests/golden/make_golden.py``).

The reference loops ``for p: for r: for q`` in interpreted Python (ueg.py:384-507); here
the same formulas are evaluated for all (p, r) at once and scattered to s = s*(p, q, r).
Sums over occupied orbitals / lattice vectors are reordered, so agreement is to
round-off (1e-12 relative), not bitwise.
"""
import numpy as np


class UEG:
    def __init__(self, n_ele, rs):
        self.n_ele = int(n_ele)
        self.rs = rs
        self.L = rs * ((4 * np.pi * self.n_ele) / 3) ** (1.0 / 3.0)          # ueg.py:66
        self.Omega = self.L ** 3                                             # ueg.py:69
        self.k_cutoff = None
        self.gamma = None

    # ---- basis: ueg.py:128-172, planewave.py:3-26 ---------------------------
    def init_single_basis(self, cutoff, k_shift=(0., 0., 0.)):
        k_shift = np.array(k_shift, dtype=float)
        imax = int(np.ceil(np.sqrt(cutoff + k_shift.dot(k_shift)))) + 1      # ueg.py:153
        self.cutoff, self.imax = cutoff, imax
        kcut = cutoff * (2 * np.pi / self.L) ** 2 / 2.                       # ueg.py:100
        rows = []
        for i in range(-imax, imax + 1):
            for j in range(-imax, imax + 1):
                for k in range(-imax, imax + 1):
                    kint = np.array([i, j, k])
                    kp = (kint + k_shift) * 2 * np.pi / self.L               # planewave.py:15
                    ke = np.dot(kp, kp) / 2.                                 # planewave.py:19
                    if ke <= kcut:
                        rows.append((ke, kint, kp))
        order = sorted(range(len(rows)), key=lambda n: rows[n][0])           # stable, planewave.py:25
        self.kint = np.array([rows[n][1] for n in order], dtype=np.int64)
        self.kp = np.array([rows[n][2] for n in order], dtype=np.float64)
        self.kinetic = np.array([rows[n][0] for n in order])
        self.n_orb = len(order)
        n = 2 * imax + 1
        self.index_map = -np.ones(n ** 3, dtype=np.int64)                    # ueg.py:119-125
        loc = n * n * (self.kint[:, 0] + imax) + n * (self.kint[:, 1] + imax) + self.kint[:, 2] + imax
        self.index_map[loc] = np.arange(self.n_orb)
        return self

    # ---- correlator: ueg.py:772-800 (non-mutating restatement) --------------
    def trunc(self, k_square):
        if self.k_cutoff is None:
            self.k_cutoff = int(np.ceil(np.sqrt(self.cutoff)))
        if self.gamma is None:
            self.gamma = 1.0
        kc2 = (self.k_cutoff * 2 * np.pi / self.L) ** 2
        k2 = np.array(k_square, dtype=np.float64, copy=True)
        k2[k2 <= kc2 * (1 + 0.00001)] = 0.
        out = np.zeros_like(k2)
        np.divide(-4. * np.pi, k2 ** 2, out=out, where=(k2 > 1e-12))
        return out * self.gamma

    # ---- sumNablaUSquare: ueg.py:581-596 ------------------------------------
    def umat(self, q_vecs, correlator, cutoff=30):
        """u_mat for float transfer vectors q_vecs [nq, 3]."""
        r = np.arange(-cutoff, cutoff + 1)
        kprime = np.stack(np.meshgrid(r, r, r, indexing="ij"), axis=-1).reshape(-1, 3)
        k1 = 2 * np.pi * kprime / self.L
        u1 = correlator(np.einsum("ni,ni->n", k1, k1))
        out = np.empty(len(q_vecs))
        for n, q in enumerate(np.asarray(q_vecs, dtype=np.float64)):
            k2 = q - k1
            out[n] = np.sum(np.einsum("ni,ni->n", k1, k2) * u1
                            * correlator(np.einsum("ni,ni->n", k2, k2))) / self.Omega
        return out

    # ---- single contractions over occupied orbitals: ueg.py:518-573 ---------
    def _ex3(self, o_kp, dvec, correlator):
        """contract_exchange_3_body(kp[o], d) for o_kp [..., 3], dvec [..., 3]."""
        occ = self.kp[: self.n_ele // 2]
        pv = o_kp[..., None, :] - occ                                        # [..., n_occ, 3]
        d2 = np.einsum("...i,...i->...", dvec, dvec)
        res = (np.einsum("...ni,...i->...n", pv, dvec) * correlator(d2)[..., None]
               * correlator(np.einsum("...ni,...ni->...n", pv, pv)))
        return res.sum(-1) / self.Omega

    def _pk(self, o_kp, dvec, correlator):
        """contractP_KWithQ(kp[o], d)."""
        occ = self.kp[: self.n_ele // 2]
        v2 = o_kp[..., None, :] - occ
        v1 = v2 - dvec[..., None, :]
        res = (np.einsum("...ni,...ni->...n", v1, v2) * correlator(np.einsum("...ni,...ni->...n", v1, v1))
               * correlator(np.einsum("...ni,...ni->...n", v2, v2)))
        return res.sum(-1) / self.Omega

    # ---- eval_2b_integrals: ueg.py:265-516 ----------------------------------
    def eval_2b_integrals(self, mode="coulomb", correlator=None):
        """Dense V[p,q,r,s]; mode in {coulomb, rpa, only_2b, only_hermi_2b, only_non_hermi_2b,
        effect_2b} (the branch order of ueg.py:411-474)."""
        nP, imax = self.n_orb, self.imax
        n = 2 * imax + 1
        kint, kp = self.kint, self.kp
        dint = kint[None, :, :] - kint[:, None, :]                           # [p, r] = k_r - k_p
        dvec = kp[None, :, :] - kp[:, None, :]
        d2 = np.einsum("pri,pri->pr", dvec, dvec)
        nz = np.abs(d2) > 0.
        safe = np.where(nz, d2, 1.0)
        ks = kint[None, None, :, :] - dint[:, :, None, :]                    # [p, r, q]
        loc = n * n * (ks[..., 0] + imax) + n * (ks[..., 1] + imax) + ks[..., 2] + imax
        ok = (loc >= 0) & (loc < n ** 3)                                     # ueg.py:401
        s = np.where(ok, self.index_map[np.clip(loc, 0, n ** 3 - 1)], -1)
        ok &= (s >= 0) & (s < nP)

        w0 = np.zeros((nP, nP))
        w1 = None
        if mode == "coulomb":                                                # ueg.py:411-413
            w0 = np.where(nz, 4. * np.pi / safe / self.Omega, 0.)
        else:
            ud = correlator(d2)
            if mode == "rpa":                                                # ueg.py:416-423
                w0 = np.where(nz, -self.n_ele * d2 * ud ** 2 / self.Omega / self.Omega, 0.)
            elif mode in ("only_2b", "only_hermi_2b", "only_non_hermi_2b"):  # ueg.py:426-457
                um = 0.
                if mode != "only_non_hermi_2b":
                    uniq, inv = np.unique(dint.reshape(-1, 3), axis=0, return_inverse=True)
                    um = self.umat(2 * np.pi * uniq / self.L, correlator)[inv.reshape(-1)].reshape(nP, nP)
                if mode == "only_non_hermi_2b":
                    w0 = np.where(nz, 4. * np.pi / safe / self.Omega, 0.)
                else:
                    w0 = np.where(nz, (4. * np.pi / safe + um + d2 * ud) / self.Omega, um / self.Omega)
                if mode != "only_hermi_2b":
                    w1 = np.where(nz, -ud / self.Omega, 0.)
            elif mode == "effect_2b":                                        # ueg.py:461-474
                kr = np.broadcast_to(kp[None, :, :], (nP, nP, 3))
                kq = np.broadcast_to(kp[:, None, :], (nP, nP, 3))
                pk = self._pk(kr, dvec, correlator)
                full = (-self.n_ele * d2 * ud ** 2 / self.Omega + 2. * self._ex3(kr, dvec, correlator)
                        - 2. * self._ex3(kq, dvec, correlator) + 2. * pk)
                w0 = np.where(nz, full, 2. * pk) / self.Omega
            else:
                raise ValueError(mode)

        V = np.zeros((nP, nP, nP, nP))
        p, r, q = np.nonzero(ok)
        ss = s[p, r, q]
        w = w0[p, r]
        if w1 is not None:
            rs_dk = kp[r] - kp[ss]
            w = w + w1[p, r] * np.einsum("ni,ni->n", rs_dk, dvec[p, r])
        V[p, q, r, ss] = w
        if mode == "effect_2b":                                              # ueg.py:509-513
            V = 0.5 * (V + V.transpose(1, 0, 3, 2))
        return V

    def tc_hamiltonian(self, no, correlator=None):
        """(fock, V) of the transcorrelated UEG as assembled in
        pymes/test/test_ueg/test_symmetrised_2body_integral.py:84-160 (pure 2-body + effective
        2-body; Fock from the pure 2-body part plus the doubly-contracted 3-body one-body shifts)."""
        correlator = self.trunc if correlator is None else correlator
        V2 = self.eval_2b_integrals("only_2b", correlator)
        f = np.diag(self.kinetic).astype(np.float64)                         # hf.py:14-18
        f += 2.0 * np.einsum("piqi->pq", V2[:, :no, :, :no])
        f -= np.einsum("piiq->pq", V2[:, :no, :no, :])
        V = V2 + self.eval_2b_integrals("effect_2b", correlator)
        f += np.diag(self.double_contractions_in_3_body(correlator))
        return f, V

    def double_contractions_in_3_body(self, correlator):
        """ueg.py:632-733 (one-body energies from doubly contracted 3-body terms)."""
        kp, ki = self.kp, self.kp[: self.n_ele // 2]
        dpi = kp[:, None, :] - ki[None, :, :]
        dpi2 = np.einsum("pij,pij->pi", dpi, dpi)
        upi = correlator(dpi2)
        perl = 2.0 * self.n_ele / self.Omega ** 2 / 2 * np.sum(upi ** 2 * dpi2, axis=1)
        wave = -np.einsum("pik,pjk,pi,pj->p", dpi, dpi, upi, upi) * 2 / self.Omega ** 2 / 2
        dij = ki[:, None, :] - ki[None, :, :]
        dij2 = np.einsum("ijk,ijk->ij", dij, dij)
        uij = correlator(dij2)
        shield = np.ones(len(kp)) * np.sum(uij ** 2 * dij2) * 2 / 2 / self.Omega ** 2
        frog = -np.einsum("ijk,pik,ij,pi->p", dij, -dpi, uij, upi) * 4 / self.Omega ** 2 / 2
        return perl + wave + shield + frog
