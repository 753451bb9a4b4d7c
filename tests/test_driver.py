"""Stand-in driver (chain_b200/driver.py): CLI flags, schedule, file names and .en/.ee formats of the reference
(src/dmrg.cpp:10-23, src/dmrg.h:360, :527-653, src/utility.cpp:30-53), and CalcEE on the device path.
Host-logic run on the emulation backend; tests/test_gpu_parity.py repeats the config-1 run on the GPU."""
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


def test_cli_rejects_unbuilt_modes():
    with pytest.raises(SystemExit):
        driver.main("-s golden -b p -l 6 -d 100 -c 1e-8 -p 1e-9 -j -1 -u 2 -q 0 -n 1 -m 1".split())
