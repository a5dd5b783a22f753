"""Shared helpers for the parity tests: build the same problem for the oracle and the CUDA path."""
import numpy as np
import torch

import closestmolecule_b200 as cm
from oracle import fem, pq


def oracle_mesh(mesh):
    return fem.Mesh(mesh.coords.cpu().numpy(), mesh.cells.cpu().numpy(), mesh.nx, mesh.ny)


def oracle_params(form):
    p = form.params
    return pq.PQParams(
        D1=p.D1, D2=p.D2, inv_dt=p.inv_dt, gkind=p.gkind, gcoef=p.gcoef, geps=p.geps, gs_thr=p.gs_thr,
        gs_cond_scale=p.gs_cond_scale, gs_lin=p.gs_lin, q_react=p.q_react, q_adv=p.q_adv,
        q_x2p=p.q_x2p, q_x2q=p.q_x2q, q_divr=p.q_divr, q_ibp=p.q_ibp, supg_stab=p.supg_stab,
        degree=p.degree, c0ip_gamma=p.c0ip_gamma, c0ip_h=p.c0ip_h, pen_stab2=p.pen_stab2,
        r2vec=(p.r2x, p.r2y))


def np_forcing(form):
    if form.forcing is None:
        return None

    def f(x, y):
        a, b = form.forcing(torch.as_tensor(x), torch.as_tensor(y))
        return a.numpy(), b.numpy()
    return f


def np_qdata(form):
    if form.q_data is None:
        return None
    return lambda x, y: form.q_data(torch.as_tensor(x), torch.as_tensor(y)).numpy()


def oracle_ext(form, om):
    """(facet ids, kinds) in the oracle's facet numbering (identical to the package's)."""
    mesh = form.space.mesh
    fc = mesh.facets()
    kind = np.zeros(fc.num, dtype=np.int64)
    ext = (fc.cell1 < 0).cpu().numpy()
    for mf, marker, bits in form.ext_terms:
        sel = ext if marker == "on_boundary" else (ext & (mf.values.cpu().numpy() == int(marker)))
        kind[sel] |= int(bits)
    idx = np.nonzero(kind)[0]
    return idx, kind[idx]


def random_state(V, seed=0, device="cpu"):
    rng = np.random.default_rng(seed)
    u = np.empty(2 * V)
    u[0::2] = 0.3 + rng.random(V)
    u[1::2] = 1.0 + rng.random(V)
    return u


def rel_row_err(vals, ref, rowptr):
    """max |vals-ref| relative to the row inf-norm of ref."""
    rowmax = np.maximum.reduceat(np.abs(ref), rowptr[:-1])
    rowmax = np.repeat(rowmax, np.diff(rowptr))
    rowmax[rowmax == 0] = 1.0
    return float(np.max(np.abs(vals - ref) / rowmax))
