"""Stand-in driver (chain_b200/driver.py): CLI flags, schedule, file names and .en/.ee formats of the reference
(src/dmrg.cpp:10-23, src/dmrg.h:360, :527-653, src/utility.cpp:30-53), and CalcEE on the device path.
Host-logic run on the emulation backend; tests/test_gpu_parity.py repeats the config-1 run on the GPU."""
import os
import re

import numpy as np
import pytest

import chain_b200 as cb
from chain_b200 import driver
import common
from oracle import dmrg as od


def test_file_stem_matches_reference_format():
    # e.g. haagerupq_p_6_1600_1e-08_0.01_0,-1,0_2_0 (tinyformat streams the coupling string and u by type)
    s = driver.file_stem("haagerupq", "p", 6, 1600, float(np.float32(1e-8)), float(np.float32(1e-2)), "0,-1,0", 2.0, 0)
    assert s == "haagerupq_p_6_1600_1e-08_0.01_0,-1,0_2_0"
    assert driver.parse_couplings("0,-1,0.000001") == [0.0, -1.0, 0.0]


def run_config1(ctx, tmp_path, nstates=2):
    sim = driver.Simulation("golden", "p", 6, 100, 1e-8, 1e-9, "-1", 2.0, 0, nstates, out=str(tmp_path), seed=3, ctx=ctx, log=lambda *a: None)
    ens, states = sim.run()
    want = common.ed_level("golden", "p", 6, [-1.0], None)
    assert np.allclose(ens, want[:nstates], atol=2e-9)
    # .en: "{<NumDoneStates()+1>,<%.10f>},\n" -- the reference numbers the first state 2 (DumpEnergy after NextState, src/dmrg.h:625-633)
    lines = open(sim.en_path).read().splitlines()
    assert len(lines) == nstates
    for k, (ln, e) in enumerate(zip(lines, ens)):
        assert ln == "{%d,%.10f}," % (k + 2, e)
    # .ee: one "{\n{b,  %.10f},\n...},\n" block per rep of the FIRST state only
    txt = open(sim.ee_path).read()
    blocks = re.findall(r"\{\n((?:\{\d+,  -?\d+\.\d{10}\},\n)+)\},\n", txt)
    assert "".join("{\n" + b + "},\n" for b in blocks) == txt
    first_state_reps = sum(1 for s in sim.sweep_log[1:] if s["nsweep"] == 1)
    assert 1 <= len(blocks) <= first_state_reps
    last = [float(x) for x in re.findall(r",  (-?\d+\.\d{10})\}", blocks[-1])]
    golden_ee = [c for c in common.golden()["ee"] if c["site"] == "golden"][0]["ee"]
    assert np.allclose(last, golden_ee, atol=5e-7)
    return sim, states


def test_driver_config1_on_emulation(emu_ctx, tmp_path):
    run_config1(emu_ctx, tmp_path)


@pytest.mark.parametrize("case", [common.DMRG_CASES[0], common.DMRG_CASES[4]], ids=["golden", "haagerupq-complex"])
def test_calc_ee_matches_oracle(emu_ctx, case):
    st, bc, n, cp, q, nst, product = case
    site, W, qk = common.build_model(st, bc, n, cp)
    psi0 = common.warm_start(site, W, n, q, seed=7)
    eo, po = od.dmrg(W, [], psi0, od.Sweeps(4, [10, 20, 40, 80], [1e-5, 1e-6, 1e-7, 1e-8]))
    ee_o = od.calc_ee(po)
    ee_g = cb.calc_ee(common.to_cb(po), site.charges, ctx=emu_ctx)
    assert np.abs(np.array(ee_o) - np.array(ee_g)).max() <= 1e-10
    # a non-canonical copy (centre unknown) must give the same entropies
    m = common.to_cb(po)
    m.center = 0
    assert np.abs(np.array(ee_o) - np.array(cb.calc_ee(m, site.charges, ctx=emu_ctx))).max() <= 1e-10


def test_cli_mode7_needs_no_device_and_bad_site_is_rejected(capsys):
    assert driver.main("-s golden -b p -l 6 -d 100 -c 1e-8 -p 1e-9 -j -1 -u 2 -q 0 -n 3 -m 7".split()) == 0
    got = [float(x) for x in capsys.readouterr().out.strip().strip("{}").split(",")]
    assert np.allclose(got, common.ed_level("golden", "p", 6, [-1.0], None)[:3], atol=1e-9)
    with pytest.raises(SystemExit):      # src/dmrg.cpp:110: "Invalid site type."
        driver.main("-s fibonacci -b p -l 6 -d 100 -c 1e-8 -p 1e-9 -j -1 -u 2 -q 0 -n 1 -m 0".split())


class _Stop(Exception):
    pass


def _run_interrupted(ctx, out, bc, nstates, seed, stop_after):
    """Run the job, killing it right after the `stop_after`-th dmrg() call has been checkpointed; returns #calls made."""
    sim = driver.Simulation("golden", bc, 6, 100, 1e-8, 1e-7, "-1", 2.0, 0, nstates, out=out, seed=seed, ctx=ctx, log=lambda *a: None)
    calls = [0]
    write = driver.DMRGProgress.Write

    def counting_write(self, path, sq, backup=False):
        write(self, path, sq, backup)
        if stop_after != "transition" and self.num_sweeps_vec_[-1] > 0 and not backup:
            calls[0] += 1
            if calls[0] == stop_after:
                raise _Stop()
    driver.DMRGProgress.Write = counting_write
    if stop_after == "transition":                          # -b sp: killed between the SSD stage and the PBC re-warm-up
        inner = sim._dmrg

        def guarded(H, done, psi, sweeps):
            if sim.dmrg_progress_.stage == 0 and sim.dmrg_progress_.num_sweeps_vec_[-1] == 0:
                calls[0] = stop_after
                raise _Stop()
            return inner(H, done, psi, sweeps)
        sim._dmrg = guarded
    try:
        sim.run()
        return sim, None
    except _Stop:
        return sim, calls[0]
    finally:
        driver.DMRGProgress.Write = write


@pytest.mark.parametrize("bc,nstates,stops", [("p", 2, (1, 3, 7)), ("sp", 1, (2, 4, "transition", 6))], ids=["periodic-2-states", "ssd-hot-start"])
def test_checkpoint_resume_reproduces_the_uninterrupted_run(emu_ctx, tmp_path, bc, nstates, stops):
    # DMRGProgress semantics (src/dmrg.h:15-257, :489-517): progress written after every dmrg() call, resume at the next sweep
    ref = driver.Simulation("golden", bc, 6, 100, 1e-8, 1e-7, "-1", 2.0, 0, nstates, out=str(tmp_path / "ref"), seed=4, ctx=emu_ctx, log=lambda *a: None)
    ens_ref, _ = ref.run()
    ncalls_ref = len(ref.sweep_log)
    for stop in stops:
        out = str(tmp_path / ("cut%s" % stop))
        sim1, n1 = _run_interrupted(emu_ctx, out, bc, nstates, 4, stop)
        assert n1 == stop                                   # it really was interrupted
        sim2 = driver.Simulation("golden", bc, 6, 100, 1e-8, 1e-7, "-1", 2.0, 0, nstates, out=out, seed=4, ctx=emu_ctx, log=lambda *a: None)
        ens2, states2 = sim2.run()
        assert len(sim1.sweep_log) + len(sim2.sweep_log) == ncalls_ref      # no sweep repeated, none skipped
        assert ens2 == ens_ref                              # bit-identical: the MPS and the stopping-rule counters round-trip exactly
        assert open(sim2.en_path).read() == open(ref.en_path).read()
        # a third start finds the job complete and does no DMRG at all (src/dmrg.h:492-511)
        sim3 = driver.Simulation("golden", bc, 6, 100, 1e-8, 1e-7, "-1", 2.0, 0, nstates, out=out, seed=4, ctx=emu_ctx, log=lambda *a: None)
        ens3, states3 = sim3.run()
        assert sim3.sweep_log == [] and ens3 == ens_ref and len(states3) == nstates
        assert all(np.array_equal(a, b) for p, q in zip(states2, states3) for a, b in zip(p.sites, q.sites))


def test_checkpoint_roundtrip_block_sparse_complex(tmp_path):
    # site tensors are stored without their forbidden charge blocks; links, centre and dtype survive
    rng = np.random.default_rng(0)
    sq = np.array([0, 1, 2, 0, 1, 2], np.int32)
    q = [np.array([0], np.int32), np.array([0, 1, 1, 2], np.int32), np.array([2, 0, 0, 1, 1], np.int32), np.array([0], np.int32)]
    sites = []
    for j in range(3):
        m = (q[j][:, None, None] + sq[None, :, None] - q[j + 1][None, None, :]) % 3 == 0
        sites.append((rng.standard_normal(m.shape) + 1j * rng.standard_normal(m.shape)) * m)
    prog = driver.DMRGProgress()
    prog.NextState()
    prog.Update(7, 400, -1.25, cb.MPS(sites, q, nq=3, center=2))
    prog.NextState()
    prog.stage, prog.min_en, prog.cnt = -1, -1.25, 2
    prog.Write(str(tmp_path / "job"), sq)
    back = driver.DMRGProgress()
    back.Read(str(tmp_path / "job"))
    assert back.num_sweeps_vec_ == [7, 0] and back.max_dims_ == [400, 0] and back.ens_ == [-1.25, 0.0]
    assert back.NumDoneStates() == 1 and back.psis_[1] is None and (back.stage, back.min_en, back.cnt) == (-1, -1.25, 2)
    p = back.psis_[0]
    assert p.nq == 3 and p.center == 2 and p.is_complex()
    assert all(np.array_equal(a, b) for a, b in zip(p.sites, sites)) and all(np.array_equal(a, b) for a, b in zip(p.link_charges, q))


@pytest.mark.parametrize("st,bc,n,cp", [("golden", "p", 6, [-1.0]), ("golden", "o", 6, [-1.0]), ("golden", "p", 8, [-1.0])])
def test_mode7_exact_levels_of_the_product_side_mpo(st, bc, n, cp):
    # Test(), src/dmrg.h:1283-1300: the reference's own cross-check -- here also a KAT of chain_b200.models against the ED vectors
    from chain_b200 import models
    mpo = models.make_siteset(st, n).hamiltonian(bc, n, 2.0, cp)
    want = common.ed_level(st, bc, n, cp, None)
    got = driver.exact_levels(mpo.sites, len(want))
    assert np.allclose(got, want, atol=1e-9)


def test_analysis_modes_on_the_progress_file(emu_ctx, tmp_path, capsys):
    # modes 1 / 2 / 3 / 4 / 6 of src/dmrg.cpp:44-60 start from the progress file of an earlier run
    from conftest import EMU_LIB
    base = ["-s", "golden", "-b", "p", "-l", "6", "-d", "100", "-c", "1e-8", "-p", "1e-9", "-j", "-1", "-u", "2", "-q", "0",
            "--out", str(tmp_path), "--seed", "3", "--lib", EMU_LIB]
    assert driver.main(base + ["-n", "3", "-m", "1"]) == 0
    assert "No progress file available" in capsys.readouterr().out
    assert driver.main(base + ["-n", "3", "-m", "0"]) == 0
    want = common.ed_level("golden", "p", 6, [-1.0], None)[:3]
    stem = driver.file_stem("golden", "p", 6, 100, float(np.float32(1e-8)), float(np.float32(1e-9)), "-1", 2.0, 0)
    an = str(tmp_path / "an" / (stem + ".an"))

    def table():
        return [[float(x) for x in ln.strip("{},\n").split(",")] for ln in open(an).read().splitlines()[1:-1]]
    capsys.readouterr()
    assert driver.main(base + ["-n", "3", "-m", "1"]) == 0                 # Analyze: {energy, momentum, rho}
    out = capsys.readouterr().out
    assert "Spectrum:  {energy, momentum, rho}" in out
    rows = table()
    assert np.allclose([r[0] for r in rows], want, atol=2e-9)
    assert [r[1] for r in rows] == [0.0, 3.0, 0.0]                         # momenta of the three lowest golden L=6 levels
    phi = (1 + 5 ** 0.5) / 2
    assert all(min(abs(r[2] - phi), abs(r[2] - (1 - phi))) < 2e-6 for r in rows)      # rho eigenvalues (src/dmrg.h:1100-1105)
    assert driver.main(base + ["-n", "2", "-m", "2"]) == 0                 # AnalyzeWithoutRho on the first two states only
    rows2 = table()
    assert len(rows2) == 2 and all(len(r) == 2 for r in rows2) and np.allclose([r[0] for r in rows2], want[:2], atol=2e-9)
    en_lines = open(str(tmp_path / "en" / (stem + ".en"))).read().splitlines()
    assert len(en_lines) == 3 and en_lines[0].startswith("{1,")            # Measure("energy") lists every finished state, 1-based
    capsys.readouterr()
    assert driver.main(base + ["-n", "3", "-m", "3"]) == 0                 # Energies
    out = capsys.readouterr().out
    assert "{{-1},{%s}}," % ",".join("%.10f" % e for e in [r[0] for r in table()]) in out
    assert driver.main(base + ["-n", "3", "-m", "4"]) == 0                 # NormalizeEnergies
    out = capsys.readouterr().out
    gap = want[1] - want[0]
    assert ("{{-1},%g,{0.0000000000,1.0000000000,%.10f}}," % (gap, (want[2] - want[0]) / gap))[:24] in out
    # Repair: the primary file is lost, the backup of the last finished state brings everything back
    os.remove(str(tmp_path / "pgs" / (stem + ".pgs")))
    assert driver.main(base + ["-n", "3", "-m", "6"]) == 0
    assert "3 finished state(s)" in capsys.readouterr().out
    assert driver.main(base + ["-n", "3", "-m", "3"]) == 0
    assert np.allclose([r[0] for r in table()], want, atol=2e-9)
