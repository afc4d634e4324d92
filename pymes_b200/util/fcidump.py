"""FCIDUMP text reader / writer (input surface of the reference, pymes/util/fcidump.py:59-163).

Host-side by nature (text parsing).  ``read`` returns the same tuple as the
reference and fills the dense tensor with the same symmetry:

* normal integrals: 4-fold (real orbitals, hermitian)   fcidump.py:143-146
* ``is_tc=True``  : only V[p,q,r,s] = V[q,p,s,r]         fcidump.py:147-149

Lines are ``value p r q s`` in chemists' order (pr|qs); V is stored in physicists'
order V[p,q,r,s] = <pq|rs>.  ``write`` is new (the reference's is dead code that
calls a ctf method) and emits files this reader round-trips.
"""
import re

import numpy as np

from ..log import print_logging_info

_TINY = 1e-19


def _header(fh):
    text = fh.readline().strip()
    while "/" not in text and "end" not in text.lower():
        nxt = fh.readline()
        if not nxt:
            break
        text += nxt.strip()
    found = {}
    for key in ("norb", "nelec"):
        m = re.search(key + r"\s*=\s*(\d+)", text, flags=re.IGNORECASE)
        found[key] = int(m.group(1)) if m else 0
    return found["norb"], found["nelec"]


def read(fcidump_file="FCIDUMP", is_tc=False):
    """-> (n_elec, n_orb, e_core, epsilon_p, h_pq, V_pqrs)"""
    print_logging_info("Reading " + fcidump_file + "...", level=1)
    print_logging_info("Using TC integrals: ", is_tc, level=2)
    e_core = 0.0
    with open(fcidump_file, "r") as fh:
        print_logging_info("Parsing header...", level=2)
        n_orb, n_elec = _header(fh)
        eps = np.zeros(n_orb)
        h = np.zeros((n_orb, n_orb))
        V = np.zeros((n_orb,) * 4)
        print_logging_info("Reading integrals...", level=2)
        for line in fh:
            fields = line.split()
            if len(fields) != 5:
                if not fields:
                    continue
                raise ValueError("malformed FCIDUMP line: %r" % line)
            val = float(fields[0])
            p, r, q, s = (int(x) for x in fields[1:])
            if abs(val) < _TINY:
                continue
            if p and q and r and s:
                p, q, r, s = p - 1, q - 1, r - 1, s - 1
                V[p, q, r, s] = val
                if is_tc:
                    V[q, p, s, r] = val
                else:
                    V[r, q, p, s] = val
                    V[r, s, p, q] = val
                    V[p, s, r, q] = val
            elif p == q == r == s == 0:
                e_core = val
            elif p and not (q or r or s):
                eps[p - 1] = val
            elif p and r and not (q or s):
                h[r - 1, p - 1] = val
                h[p - 1, r - 1] = val
    return n_elec, n_orb, e_core, eps, h, V


def read_blocks(fcidump_file="FCIDUMP", is_tc=False, rows=None, sharded_dims=None):
    """Same file, same symmetry fill as :func:`read`, but the two-body integrals go straight into
    the 16 partition blocks of ``integral.partition`` -- V_pqrs (nP^4 doubles: 732 GB for the
    o = 50, v = 500 configuration) never exists.

    ``rows = (lo, n)`` with ``sharded_dims = {key: dim}`` (``parallel.SHARD_DIMS``) keeps, for the
    listed keys, only the virtual-orbital rows [lo, lo+n) along that dimension: what one rank of
    a sharded run holds.  -> (n_elec, n_orb, e_core, epsilon_p, h_pq, {key: ndarray})"""
    from ..integral.partition import KEYS, OCCUPIED
    print_logging_info("Reading " + fcidump_file + " into partition blocks...", level=1)
    e_core = 0.0
    sharded_dims = sharded_dims or {}
    with open(fcidump_file, "r") as fh:
        n_orb, n_elec = _header(fh)
        no = n_elec // 2
        nv = n_orb - no
        eps = np.zeros(n_orb)
        h = np.zeros((n_orb, n_orb))
        pattern, blocks, window = {}, {}, {}
        for key in KEYS:
            pat = tuple(ch in OCCUPIED for ch in key)
            pattern[pat] = key
            shape = [no if o else nv for o in pat]
            window[key] = None
            if rows is not None and key in sharded_dims:
                shape[sharded_dims[key]] = int(rows[1])
                window[key] = (sharded_dims[key], int(rows[0]), int(rows[1]))
            blocks[key] = np.zeros(shape)

        def put(p, q, r, s, val):
            idx = (p, q, r, s)
            pat = tuple(x < no for x in idx)
            key = pattern[pat]
            loc = [x if o else x - no for x, o in zip(idx, pat)]
            w = window[key]
            if w is not None:
                loc[w[0]] -= w[1]
                if not 0 <= loc[w[0]] < w[2]:
                    return
            blocks[key][tuple(loc)] = val

        for line in fh:
            fields = line.split()
            if len(fields) != 5:
                if not fields:
                    continue
                raise ValueError("malformed FCIDUMP line: %r" % line)
            val = float(fields[0])
            p, r, q, s = (int(x) for x in fields[1:])
            if abs(val) < _TINY:
                continue
            if p and q and r and s:
                p, q, r, s = p - 1, q - 1, r - 1, s - 1
                put(p, q, r, s, val)
                if is_tc:
                    put(q, p, s, r, val)
                else:
                    put(r, q, p, s, val)
                    put(r, s, p, q, val)
                    put(p, s, r, q, val)
            elif p == q == r == s == 0:
                e_core = val
            elif p and not (q or r or s):
                eps[p - 1] = val
            elif p and r and not (q or s):
                h[r - 1, p - 1] = val
                h[p - 1, r - 1] = val
    return n_elec, n_orb, e_core, eps, h, blocks


def write(fcidump_file, n_elec, h_pq, V_pqrs, e_core=0.0, is_tc=False, ms2=0, thresh=_TINY):
    """Write integrals so that ``read(file, is_tc)`` reproduces them.

    Only one representative of each symmetry-equivalent set is written (the reader
    restores the others)."""
    h = np.asarray(h_pq)
    V = np.asarray(V_pqrs)
    n = h.shape[0]
    with open(fcidump_file, "w") as fh:
        fh.write(" &FCI NORB=%d,NELEC=%d,MS2=%d,\n  ORBSYM=%s\n  ISYM=1,\n &END\n"
                 % (n, n_elec, ms2, "1," * n))
        for p in range(n):
            for q in range(n):
                for r in range(n):
                    for s in range(n):
                        v = V[p, q, r, s]
                        if abs(v) < thresh:
                            continue
                        if is_tc:
                            if (q, p, s, r) < (p, q, r, s):
                                continue
                        elif min((p, q, r, s), (r, q, p, s), (r, s, p, q), (p, s, r, q)) != (p, q, r, s):
                            continue
                        fh.write("%28.20E %4d %4d %4d %4d\n" % (v, p + 1, r + 1, q + 1, s + 1))
        for p in range(n):
            for r in range(p + 1):
                if abs(h[p, r]) >= thresh:
                    fh.write("%28.20E %4d %4d %4d %4d\n" % (h[p, r], p + 1, r + 1, 0, 0))
        fh.write("%28.20E %4d %4d %4d %4d\n" % (e_core, 0, 0, 0, 0))
