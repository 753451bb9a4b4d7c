"""Edge cases and error behaviour of the boundary (host logic on the emulation backend; the CUDA library shares the engine):
boundary conditions the reference supports (src/golden_site.cpp:68-117: p / o / s / d / sp), ragged and empty charge
sectors, bond dimension 1, invalid input -> status codes (no exception crosses the C ABI, src/dmrg.h:454 contract)."""
import os
import sys

import numpy as np
import pytest

import chain_b200 as cb
from chain_b200 import driver
import common
from oracle import dmrg as od


@pytest.mark.parametrize("st,bc,n,cp", [("golden", "s", 7, [-1.0]), ("golden", "d", 6, [-1.0]), ("haagerupq", "s", 6, [-1.0, 0.0, 0.0]),
                                        ("haagerup", "o", 5, [0.0, 0.0, -1.0])])
def test_boundary_conditions_track_the_oracle(emu_ctx, st, bc, n, cp):
    site, W, qk = common.build_model(st, bc, n, cp)
    H = cb.MPO(W, qk, site.charges, nq=site.nq)
    psi0 = common.warm_start(site, W, n, 0, seed=11) if site.nq > 1 else od.product_mps(site, n, 0, np.random.default_rng(4))
    sw = ([10, 20, 40], [1e-5, 1e-7, 1e-9])
    eo, po = od.dmrg(W, [], psi0, od.Sweeps(6, *sw))
    eg, pg = cb.dmrg(H, [], common.to_cb(psi0), cb.Sweeps(6, *sw), ctx=emu_ctx)
    assert abs(eo - eg) <= 1e-10 * max(1.0, abs(eo))
    assert pg.bond_dims() == po.bond_dims()


def test_sp_hot_start_schedule(emu_ctx, tmp_path):
    # -b sp: SSD warm-up, then the PBC run starts from the SSD state with the second warm-up variant (src/dmrg.h:548-567, :613-621)
    sim = driver.Simulation("golden", "sp", 6, 100, 1e-8, 1e-7, "-1", 2.0, 0, 1, out=str(tmp_path), seed=5, ctx=emu_ctx, log=lambda *a: None)
    ens, states = sim.run()
    assert abs(ens[0] - common.ed_level("golden", "p", 6, [-1.0], None)[0]) <= 1e-7
    lines = open(sim.en_path).read().splitlines()
    assert len(lines) == 2 and lines[0].startswith("{1,") and lines[1].startswith("{2,")   # SSD stage then the PBC state
    warm = [s for s in sim.sweep_log if s["nsweep"] == 5]
    assert len(warm) == 2


def test_ragged_and_empty_sectors(emu_ctx):
    # a Z3 bond whose left link has an EMPTY charge sector and unequal others
    rng = np.random.default_rng(8)
    site, W, qk = common.build_model("haagerupq", "p", 6, [-1.0, 0.0, 0.0])
    ql = np.array([0] * 3 + [2] * 5, np.int32)                 # no charge-1 states on the left link
    qr = np.array([0] * 4 + [1] * 2 + [2] * 1, np.int32)
    sq = site.charges
    kl, km, kr = qk[2], qk[3], qk[4]
    def rnd(shape, mask):
        return rng.standard_normal(shape) * mask
    Lm = (ql[:, None, None] - kl[None, :, None] - ql[None, None, :]) % 3 == 0
    Rm = (qr[:, None, None] - kr[None, :, None] - qr[None, None, :]) % 3 == 0
    L, R = rnd(Lm.shape, Lm), rnd(Rm.shape, Rm)
    pm = (ql[:, None, None, None] + sq[None, :, None, None] + sq[None, None, :, None] - qr[None, None, None, :]) % 3 == 0
    phi = rnd(pm.shape, pm)
    B = cb.Bond(L, W[2], W[3], R, ql=ql, qr=qr, qkl=kl, qkm=km, qkr=kr, site_charges=sq, nq=3, ctx=emu_ctx)
    y = B.apply(phi)
    yo = od.heff_apply(L, W[2], W[3], R, phi)
    assert np.abs(y - yo).max() <= 1e-12 * max(1.0, np.abs(yo).max())
    assert np.abs(y[~pm]).max() == 0.0                          # forbidden blocks stay exactly zero
    phi /= np.linalg.norm(phi)
    A1, B1, qm, spec, terr = B.truncate(phi, "left", 100, 1e-14)
    assert np.abs(np.tensordot(A1, B1, axes=([2], [0])) - phi).max() <= 1e-12
    B.close()


def test_bond_dimension_one(emu_ctx):
    rng = np.random.default_rng(2)
    L = np.ones((1, 1, 1)); R = np.ones((1, 1, 1))
    W = np.zeros((1, 2, 2, 1)); W[0, :, :, 0] = np.array([[0.3, 0.1], [0.1, -0.2]])
    B = cb.Bond(L, W, W, R, ctx=emu_ctx)
    phi = rng.standard_normal((1, 2, 2, 1))
    want = np.einsum("ts,uv,sv->tu", W[0, :, :, 0], W[0, :, :, 0], phi[0, :, :, 0])
    assert np.abs(B.apply(phi)[0, :, :, 0] - want).max() <= 1e-14
    ev, x, nmv = B.davidson(phi, 2)
    assert nmv == 3 and abs(np.linalg.norm(x) - 1) < 1e-13
    B.close()


def test_invalid_inputs_give_status_codes(emu_ctx):
    site, W, qk = common.build_model("golden", "p", 6, [-1.0])
    H = cb.MPO(W, qk, site.charges, nq=1)
    psi = common.to_cb(od.product_mps(site, 6, 0, np.random.default_rng(0)))
    with pytest.raises(cb.ChainB200Error) as e:                 # noise != 0 is not part of Chain's path (src/dmrg.h:350)
        cb.dmrg(H, [], psi, cb.Sweeps(1, 10, 1e-8, noise=1e-6), ctx=emu_ctx)
    assert e.value.code == -2 and "noise" in str(e.value)
    bad = cb.MPS(psi.sites[:5], None, nq=1)                      # length mismatch
    with pytest.raises(cb.ChainB200Error) as e:
        cb.dmrg(H, [], bad, cb.Sweeps(1, 10, 1e-8), ctx=emu_ctx)
    assert e.value.code == -2
    q_bad = [np.zeros(1, np.int32)] + [np.full(a.shape[2], 5, np.int32) for a in psi.sites]   # charge label out of range
    with pytest.raises(cb.ChainB200Error):
        cb.dmrg(H, [], cb.MPS(psi.sites, q_bad, nq=1), cb.Sweeps(1, 10, 1e-8), ctx=emu_ctx)
    # a Z3-violating "MPO" is detected by the planner, not silently contracted
    site3, W3, qk3 = common.build_model("haagerupq", "p", 6, [-1.0, 0.0, 0.0])
    Wb = [w.copy() for w in W3]
    Wb[2][0, 0, 1, 0] = 1.0                                      # q(s_out=0) != q(s_in=1) with neutral links
    c = common.capture_bond("haagerupq", "p", 6, [-1.0, 0.0, 0.0], 0, 3)[3]
    with pytest.raises(cb.ChainB200Error) as e:
        cb.Bond(c["L"], Wb[2], W3[3], c["R"], ql=c["ql"], qr=c["qr"], qkl=qk3[2], qkm=qk3[3], qkr=qk3[4], site_charges=site3.charges, nq=3, ctx=emu_ctx)
    assert e.value.code == -2 and "violates" in str(e.value)
    # the context is still usable after the errors
    en, _ = cb.dmrg(H, [], psi, cb.Sweeps(2, 10, 1e-8), ctx=emu_ctx)
    assert np.isfinite(en)


def test_cut_inside_a_degenerate_pair_dense_keeps_m_qn_drops_the_pair(emu_ctx):
    """ITensor's two truncation conventions (diagHImpl / svdImpl): a tensor WITHOUT QNs keeps exactly the m columns truncate() decided on
    (`reduceCols(UU, m)`), a QN tensor keeps per block the eigenvalues above the pooled threshold docut -- and docut is raised by 1e-3 P(m-1)
    when the cut would split a (near-)degenerate pair, which drops the whole pair.  Designed spectrum (0.4, 0.3, 0.1, 0.1, 0.05, 0.05), maxdim 3."""
    sys_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
    if sys_path not in sys.path:
        sys.path.insert(0, sys_path)
    import fuzz_truncate as ft
    p = np.array([0.4, 0.3, 0.1, 0.1, 0.05, 0.05])
    for nq, want in ((1, 3), (3, 2)):
        rng = np.random.default_rng(4)
        d = 6 if nq == 3 else 2
        sq = np.array([0, 1, 2, 0, 1, 2], np.int32) if nq == 3 else np.zeros(d, np.int32)
        Dl = Dr = 6
        ql = (np.arange(Dl) % nq).astype(np.int32)
        qr = (np.arange(Dr) % nq).astype(np.int32)
        qm = (np.arange(len(p)) % nq).astype(np.int32)
        rowq = ((ql[:, None] + sq[None, :]) % nq).reshape(-1)
        colq = ((qr[None, :] - sq[:, None]) % nq).reshape(-1)
        U, V = ft.block_isometry(rng, rowq, qm, nq, False), ft.block_isometry(rng, colq, qm, nq, False)
        phi = np.asfortranarray(((U * np.sqrt(p)) @ V.T).reshape(Dl, d, d, Dr))
        k = 2
        E = lambda D, q: rng.standard_normal((D, k, D)) * (q[:, None, None] == q[None, None, :])
        W = np.zeros((k, d, d, k))
        for a in range(k):
            W[a, :, :, a] = np.eye(d)
        zk = np.zeros(k, np.int32)
        B = cb.Bond(np.asfortranarray(E(Dl, ql)), W, W, np.asfortranarray(E(Dr, qr)), ql=ql, qr=qr, qkl=zk, qkm=zk, qkr=zk, site_charges=sq, nq=nq, ctx=emu_ctx)
        for direction in ("left", "right"):
            A1, B1, qg, spec, terr = B.truncate(phi, direction, 3, 1e-12)
            A2, B2, qo, speco, terro = od.split_bond(phi, ql, qr, sq, nq, direction, 3, 1, 1e-12)
            assert len(spec) == len(speco) == want, (nq, direction, len(spec), len(speco))
            assert abs(terr - 0.2) <= 1e-12 and abs(terro - 0.2) <= 1e-12          # truncate() itself discards 0.1 + 0.05 + 0.05 either way
            iso = A1 if direction == "left" else B1
            g = np.tensordot(iso.conj(), iso, axes=([0, 1], [0, 1])) if direction == "left" else np.tensordot(iso, iso.conj(), axes=([1, 2], [1, 2]))
            assert np.abs(g - np.eye(want)).max() <= 1e-12
        B.close()
