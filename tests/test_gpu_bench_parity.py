"""-m gpu: parity AT THE BENCHMARKED SIZES (VERDICT r1 "nothing at bench size is parity-checked").

  * bench.py's own cfg2 workload (haagerupq D=1600 Z3(534,533,533) k=26 complex128) and the cfg4 shape (haagerup D=3000 dense f64): a
    sample of output rows of H_eff * phi against a NumPy evaluation of the same contraction -- complete rows (all s1', s2', r') of a few bra
    indices l' per charge sector -- with the 3-multiplication AND the 4-multiplication complex kernels, to 1e-12 * sqrt(K) relative to the
    largest output element (Karatsuba's error is normwise, so the bound is normwise; K = D * k is the longest contraction);
  * one complete truncation at D=1600: the kept spectrum against numpy.linalg.eigvalsh of the three rho blocks;
  * BASELINE configs[2] as written (golden periodic L=24, five states, penalty weight 1000) against the exact levels of SURVEY App. C and
    against the oracle on the same start; configs[3]'s physical run (haagerup open L=12) against SURVEY App. F.
The two-rank versions of the matvec checks are in tests/test_gpu_shard.py (they need two GPUs)."""
import os
import sys

import numpy as np
import pytest

import chain_b200 as cb
import common
from chain_b200 import driver
from conftest import ROOT
from oracle import dmrg as od

sys.path.insert(0, ROOT)
import bench  # noqa: E402

pytestmark = pytest.mark.gpu


def heff_rows_numpy(w, rows):
    """out[l', s1', s2', r'] for l' in rows = sum L[l',a,l] W1[a,s1',s1,b] W2[b,s2',s2,c] R[r',c,r] phi[l,s1,s2,r] (ITensor LocalMPO::product)."""
    L, W1, W2, R, phi = w["L"], w["W1"], w["W2"], w["R"], w["phi"]
    D, d = phi.shape[0], phi.shape[1]
    kl, kr = L.shape[1], R.shape[1]
    Ls = L[rows]                                                       # [S, a, l]
    t1 = (Ls.reshape(-1, D) @ phi.reshape(D, -1)).reshape(len(rows), kl, d, d, phi.shape[3])      # [S, a, s1, s2, r]
    t2 = np.einsum("xasur,atsb->xbtur", t1, W1, optimize=True)         # [S, b, s1', s2, r]
    t3 = np.einsum("xbtur,bvuc->xtvcr", t2, W2, optimize=True)         # [S, s1', s2', c, r]
    Rm = R.transpose(1, 2, 0).reshape(kr * R.shape[2], R.shape[0])     # [(c, r), r']
    out = t3.reshape(len(rows) * d * d, -1) @ Rm
    return out.reshape(len(rows), d, d, R.shape[0])


def _check_rows(ctx, w, rows, tol_scale):
    bond = cb.Bond(w["L"], w["W1"], w["W2"], w["R"], ql=w["ql"], qr=w["qr"], qkl=w["qkl"], qkm=w["qkm"], qkr=w["qkr"],
                   site_charges=w["sq"], nq=w["nq"], ctx=ctx)
    got = bond.apply(w["phi"])
    bond.close()
    want = heff_rows_numpy(w, rows)
    scale = np.abs(got).max()
    K = w["phi"].shape[0] * w["L"].shape[1]
    err = np.abs(got[rows] - want).max()
    # forbidden blocks of the output stay exactly zero
    nq = w["nq"]
    mask = (np.asarray(w["ql"])[:, None, None, None] + np.asarray(w["sq"])[None, :, None, None] + np.asarray(w["sq"])[None, None, :, None]
            - np.asarray(w["qr"])[None, None, None, :]) % nq == 0
    assert not np.any(got[~mask])
    assert err <= tol_scale * 1e-12 * np.sqrt(K) * scale, (err, scale, K)
    return err / scale


@pytest.mark.parametrize("m3", ["96", "0"], ids=["3M-64x96", "4M-64x128"])
def test_cfg2_bench_workload_rows_match_numpy(gpu_ctx, monkeypatch, m3):
    monkeypatch.setenv("CB200_CPLX_3M", m3)
    w = bench.build_workload("cfg2", seed=2000)                         # exactly what bench.py times
    D = w["phi"].shape[0]
    rows = np.array([0, 1, 533, 534, 535, 1066, 1067, D - 1])           # both ends of every Z3 sector of the left link
    rel = _check_rows(gpu_ctx, w, rows, 1.0)
    print("cfg2 %s: max |H phi - numpy| / max|H phi| = %.2e over %d complete rows" % (m3, rel, len(rows)))


def test_cfg4_shape_rows_match_numpy(gpu_ctx):
    w = bench.build_workload("cfg4", seed=2000)                         # haagerup open, D = 3000 dense f64, k = 14
    rows = np.array([0, 1, 127, 128, 1499, 2047, 2998, 2999])           # tile borders of the 128-row GEMM tiles included
    rel = _check_rows(gpu_ctx, w, rows, 1.0)
    print("cfg4: max |H phi - numpy| / max|H phi| = %.2e over %d complete rows" % (rel, len(rows)))


def test_truncation_spectrum_at_bench_size_matches_eigvalsh(gpu_ctx):
    """One svdBond at D = 1600 (BASELINE configs[1] shape): kept eigenvalues of rho vs LAPACK on the three charge blocks."""
    w = bench.build_workload("cfg2", seed=7)
    bond = cb.Bond(w["L"], w["W1"], w["W2"], w["R"], ql=w["ql"], qr=w["qr"], qkl=w["qkl"], qkm=w["qkm"], qkr=w["qkr"],
                   site_charges=w["sq"], nq=w["nq"], ctx=gpu_ctx)
    phi = w["phi"]
    D, d, nq = phi.shape[0], phi.shape[1], w["nq"]
    A, B, qm, spec, terr = bond.truncate(phi, "left", D, 1e-12)
    bond.close()
    ql, sq = np.asarray(w["ql"]), np.asarray(w["sq"])
    X = phi.reshape(D * d, d * phi.shape[3], order="F")                 # rows (l, s1) with l fastest
    qrow = (ql[:, None] + sq[None, :]).reshape(-1, order="F") % nq
    ev = []
    for q in range(nq):
        Xq = X[qrow == q]
        ev.append(np.linalg.eigvalsh(Xq @ Xq.conj().T))
    ev = np.sort(np.concatenate(ev))[::-1]
    m = A.shape[2]
    assert D - 8 <= m <= D          # the degeneracy guard on docut (ITensor truncate) may drop a few states at the cut
    assert np.abs(np.sort(spec)[::-1] - ev[:m]).max() <= 1e-14 * np.sqrt(X.shape[0])
    # ITensor's truncate reports the weight discarded by the maxdim / cutoff rule (n = maxdim states kept); the docut degeneracy guard can
    # then drop the few states at the cut as well (m <= maxdim), SURVEY App. A.5
    assert abs(terr - ev[D:].sum()) <= 1e-12
    # the kept vectors reproduce phi up to the weight actually discarded
    AB = np.tensordot(A, B, axes=([2], [0]))
    ov = abs(np.vdot(AB, phi)) ** 2 / (np.vdot(AB, AB).real * np.vdot(phi, phi).real)
    assert abs(ov - (1.0 - ev[m:].sum())) <= 1e-10


# golden periodic L = 24, K = -1, U = 2 (exact, 2^24-dimensional matrix-free Lanczos)
GOLDEN_L24 = [-18.362169819967, -18.326270264850, -18.266979632791, -17.939685031845, -17.853736702373]


def test_config3_five_states_as_written(gpu_ctx, tmp_path):
    """`-s golden -b p -l 24 -d 800 -n 5` (BASELINE configs[2]) through the stand-in driver: reference schedule (src/dmrg.h:527-653), penalty
    weight 1000, stopping rule at -p 1e-9.  Energies vs the exact levels (the run is truncation-limited at cutoff 1e-8: ~1e-6, SURVEY
    App. F), and the first two states against the oracle driven through the SAME schedule from the SAME start."""
    sim = driver.Simulation("golden", "p", 24, 800, 1e-8, 1e-9, "-1", 2.0, 0, 5, out=str(tmp_path), seed=11, ctx=gpu_ctx, log=lambda *a: None,
                            checkpoint=False)
    ens, states = sim.run()
    assert len(ens) == 5
    assert np.abs(np.array(ens) - np.array(GOLDEN_L24)).max() <= 5e-6, ens
    assert all(b - a > 1e-3 for a, b in zip(ens, ens[1:]))              # five distinct levels in ascending order
    assert max(max(p.bond_dims()) for p in states) <= 800
    # the oracle on the same start and schedule (ground state: same trajectory; the penalised state: compared at convergence)
    site, W, qk = common.build_model("golden", "p", 24, [-1.0])
    c, prec = float(np.float32(1e-8)), float(np.float32(1e-9))
    psi0 = common.to_oracle(sim.sites.init_state(0, np.random.default_rng(11), dtype=float), site)   # the driver's init_state_ (one per job, :516)
    done = []
    for s in range(2):
        en, psi = od.dmrg(W, done, psi0, od.Sweeps(5, [10, 20, 40, 80, 200], [max(x, c) for x in (1e-5, 1e-6, 1e-7, 1e-8, 1e-9)]), weight=1000.0)
        min_en, cnt, bond_dim = 1000000.0, 3, 200
        while cnt > 0:                                                   # src/dmrg.h:576-609
            if bond_dim <= 800 - 200:
                bond_dim += 200
            en, psi = od.dmrg(W, done, psi, od.Sweeps(1, bond_dim, c), weight=1000.0)
            if abs(min_en - en) * 24 < prec:
                cnt -= 1
            else:
                cnt, min_en = 3, en
        done.append(psi)
        tol = 1e-10 if s == 0 else 2e-7
        assert abs(en - ens[s]) <= tol * abs(en), (s, en, ens[s])


def test_config4_physical_run(gpu_ctx, tmp_path):
    """`-s haagerup -b o -l 12 -d 3000` (BASELINE configs[3]'s physical run; D is never reached at cutoff 1e-8): SURVEY App. F measured
    E = -8.6689857637 with bond profile 5,14,44,113,160,188,159,107,43,13,4 for a state of the degenerate open-chain ground manifold."""
    sim = driver.Simulation("haagerup", "o", 12, 3000, 1e-8, 1e-6, "0,-1,0", 2.0, 0, 1, out=str(tmp_path), seed=5, ctx=gpu_ctx,
                            log=lambda *a: None, checkpoint=False)
    ens, states = sim.run()
    assert abs(ens[0] - (-8.6689857637)) <= 2e-6, ens
    prof = states[0].bond_dims()
    assert 150 <= max(prof) <= 260 and prof[0] <= 6 and prof[-1] <= 6, prof
