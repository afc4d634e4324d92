"""The ``is_dr_ccd`` branch of the reference CCD solver, reproduced as the reference EXECUTES it.

Call surface of ``pymes.solver.drccd`` (pymes/solver/drccd.py:10-48): ``get_residual(tEpsilon_i,
tEpsilon_a, tT_abij, tV_abij, tV_aijb, tV_iabj, tV_ijab)`` and ``getEnergy(tT_abij, tV_ijab)``.

The einsum strings of drccd.py:34-35 are not the direct-ring equations of the comment above
them: ``"kbcj, acij -> abij"`` sums k over V alone and carries j as a batch index, and
``"acij, klcd, dblj -> abij"`` again sums k over V alone with j as a batch index.  The tier's
bar is "results identical to the reference's", so exactly those sums are evaluated here: the
k-sums as contractions with a vector of ones, the batch-index products with ``pmb_bdot`` (an
index shared by both operands AND the output is not a matrix product)."""
import torch

from .. import backend as bk


def residual_device(eps_i, eps_a, T2, V_abij, V_aijb, V_iabj, V_ijab):
    """Device tensors in, R_abij (new tensor) out.  drccd.py:28-36 line by line."""
    nv, no = T2.shape[0], T2.shape[2]
    ones = torch.ones(no, dtype=bk.F64, device=T2.device)
    R = bk.copy(V_abij)
    bk.bdot("a,abij->abij", eps_a, T2, out=R, alpha=+1.0, beta=1.0)       # "ad,dbij->abij", f_ab diagonal
    bk.bdot("i,abij->abij", eps_i, T2, out=R, alpha=-1.0, beta=1.0)       # "ik,abkj->abij"
    bk.bdot("b,baji->abij", eps_a, T2, out=R, alpha=+1.0, beta=1.0)       # "bd,daji->abij"
    bk.bdot("j,baji->abij", eps_i, T2, out=R, alpha=-1.0, beta=1.0)       # "jk,baki->abij"
    bk.contract("akic,cbkj->abij", V_aijb, T2, out=R, beta=1.0)           # drccd.py:33
    Vs = bk.contract("kbcj,k->bcj", V_iabj, ones)                         # drccd.py:34: k summed on V only
    bk.bdot("bcj,acij->abij", Vs, T2, out=R, beta=1.0)                    #   ... j is a batch index
    Ws = bk.contract("klcd,k->lcd", V_ijab, ones)                         # drccd.py:35: k summed on V only
    X = bk.contract("acij,lcd->aijld", T2, Ws)
    bk.bdot("aijld,dblj->abij", X, T2, out=R, beta=1.0)                   #   ... j is a batch index
    return R


def get_residual(tEpsilon_i, tEpsilon_a, tT_abij, tV_abij, tV_aijb, tV_iabj, tV_ijab):
    want_numpy = not isinstance(tT_abij, torch.Tensor)
    R = residual_device(bk.asdev(tEpsilon_i).contiguous(), bk.asdev(tEpsilon_a).contiguous(),
                        bk.asdev(tT_abij).contiguous(), bk.asdev(tV_abij), bk.asdev(tV_aijb),
                        bk.asdev(tV_iabj), bk.asdev(tV_ijab))
    return bk.tonumpy(R) if want_numpy else R


def getEnergy(tT_abij, tV_ijab):
    """[2 T.V, 0]: drccd.py:41-48 (the exchange part is commented out there)."""
    scal = bk.zeros(8)
    bk.energy_doubles(bk.asdev(tT_abij).contiguous(), bk.asdev(tV_ijab), scal)
    return [float(scal[0].item()), 0.]
