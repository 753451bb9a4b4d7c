"""Randomised block structures (host logic on the emulation backend): ragged / empty Z_N sectors on every index, real and
complex data, every entry point of one bond update against the dense oracle.  The planner's offset arithmetic is what these
exercise; the -m gpu suite repeats one draw per model through the CUDA kernels (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import chain_b200 as cb
import common
from oracle import dmrg as od


def random_bond(rng, st, cp, dims_l, dims_r, cplx_data=False):
    site, W, qk = common.build_model(st, "p", 6, cp)
    nq = site.nq
    sq = site.charges
    kl, km, kr = qk[2], qk[3], qk[4]
    ql = np.concatenate([np.full(m, q, np.int32) for q, m in enumerate(dims_l)]) if nq > 1 else np.zeros(sum(dims_l), np.int32)
    qr = np.concatenate([np.full(m, q, np.int32) for q, m in enumerate(dims_r)]) if nq > 1 else np.zeros(sum(dims_r), np.int32)
    rng.shuffle(ql); rng.shuffle(qr)                       # external order is arbitrary: the boundary sorts by charge
    cplx = cplx_data or np.iscomplexobj(W[2])
    def rnd(mask):
        x = rng.standard_normal(mask.shape) + (1j * rng.standard_normal(mask.shape) if cplx else 0)
        return x * mask
    Lm = (ql[:, None, None] - kl[None, :, None] - ql[None, None, :]) % nq == 0
    Rm = (qr[:, None, None] - kr[None, :, None] - qr[None, None, :]) % nq == 0
    pm = (ql[:, None, None, None] + sq[None, :, None, None] + sq[None, None, :, None] - qr[None, None, None, :]) % nq == 0
    return dict(site=site, W1=W[2], W2=W[3], kl=kl, km=km, kr=kr, ql=ql, qr=qr, L=rnd(Lm), R=rnd(Rm), phi=rnd(pm), pm=pm)


def check(ctx, b):
    site = b["site"]
    B = cb.Bond(b["L"], b["W1"], b["W2"], b["R"], ql=b["ql"], qr=b["qr"], qkl=b["kl"], qkm=b["km"], qkr=b["kr"],
                site_charges=site.charges, nq=site.nq, ctx=ctx)
    y = B.apply(b["phi"])
    yo = od.heff_apply(b["L"], b["W1"], b["W2"], b["R"], b["phi"])
    assert np.abs(y - yo).max() <= 2e-12 * max(1.0, np.abs(yo).max())
    assert not np.any(y[~b["pm"]])
    phi = b["phi"] / np.linalg.norm(b["phi"])
    for d in ("left", "right"):
        A1, B1, qm, spec, terr = B.truncate(phi, d, 10 ** 6, 1e-13)
        A2, B2, qmo, speco, terro = od.split_bond(phi, b["ql"], b["qr"], site.charges, site.nq, d, 10 ** 6, 1, 1e-13)
        assert np.abs(np.tensordot(A1, B1, axes=([2], [0])) - phi).max() <= 1e-6      # discarded weight <= 1e-13
        keep = min(len(spec), len(speco))
        assert np.abs(np.sort(spec)[::-1][:keep] - np.sort(speco)[::-1][:keep]).max() <= 1e-12
        T = A1 if d == "left" else B1
        E = B.env_update(d, T, qm)
        Eo = od.update_left(b["L"], T, b["W1"]) if d == "left" else od.update_right(b["R"], T, b["W2"])
        assert np.abs(E - Eo).max() <= 2e-12 * max(1.0, np.abs(Eo).max())
    B.close()


@pytest.mark.parametrize("seed", range(8))
def test_random_z3_structures(emu_ctx, seed):
    rng = np.random.default_rng(100 + seed)
    dims_l = rng.integers(0, 6, 3)
    dims_r = rng.integers(0, 6, 3)
    if dims_l.sum() == 0: dims_l[rng.integers(3)] = 2
    if dims_r.sum() == 0: dims_r[rng.integers(3)] = 2
    cp = [[-1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [-0.5, 0.3, 0.2]][seed % 4]
    b = random_bond(rng, "haagerupq", cp, dims_l.tolist(), dims_r.tolist(), cplx_data=(seed % 2 == 1))
    if b["phi"].size == 0 or not np.any(b["phi"]):
        pytest.skip("no allowed block for this draw")
    check(emu_ctx, b)
    check_blocks(emu_ctx, b)


def check_blocks(ctx, b):
    """The block-packed (QDense order) boundary on the same random structure: shuffled labels = many short sectors, empty charges."""
    site = b["site"]
    B = cb.Bond(b["L"], b["W1"], b["W2"], b["R"], ql=b["ql"], qr=b["qr"], qkl=b["kl"], qkm=b["km"], qkr=b["kr"],
                site_charges=site.charges, nq=site.nq, ctx=ctx)
    assert B.phi_blocks_elems == int(b["pm"].sum())
    xb = B.pack(b["phi"])
    assert np.array_equal(B.unpack(xb), np.asarray(b["phi"], dtype=xb.dtype))
    blocks = B.phi_blocks()
    assert sum(int(np.prod(d)) for _, d in blocks) == xb.size
    assert np.array_equal(B.apply_blocks(xb), B.pack(B.apply(b["phi"])))
    B.close()


@pytest.mark.parametrize("st,cp,dl,dr", [("golden", [-1.0], [7], [3]), ("haagerup", [0.0, -1.0, 0.0], [3], [5]), ("golden", [-1.0], [1], [9])])
def test_random_dense_structures(emu_ctx, st, cp, dl, dr):
    rng = np.random.default_rng(sum(dl) * 10 + sum(dr))
    check(emu_ctx, random_bond(rng, st, cp, dl, dr))
