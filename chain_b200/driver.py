"""Stand-in for the reference's driver (/root/reference/src/dmrg.cpp + DMRG<SiteSet>::SimulateRaw, src/dmrg.h:474-655).

The real driver stays unchanged and reaches this library through the adapter of INTEGRATION.md; it cannot be compiled
where ITensor is absent, so this module restates the part of it that surrounds the hot path -- the CLI flags
(src/dmrg.cpp:10-23), the warm-up / ramp / stopping schedule (src/dmrg.h:527-609), the file names (src/dmrg.h:360-365)
and the .en / .ee writers (src/utility.cpp:30-53) -- and calls cb200_dmrg for every `dmrg(...)` of the reference.
Nothing here does tensor arithmetic: Hamiltonians come from chain_b200.models (host, O(L)), DMRG and CalcEE run on the GPU.

    python -m chain_b200.driver -s golden -b p -l 6 -d 100 -c 1e-8 -p 1e-9 -j -1 -u 2 -q 0 -n 1 -m 0 [--out DIR] [--seed N]

Differences from the reference, all deliberate: output goes under --out (the reference compiles the parent of the cwd in,
src/path_template.h:8); --seed makes the random product start reproducible (the reference never seeds); there is no
pgs/ checkpoint (ITensor's private binary format); mode 5 runs the simulation and then the energy /
translation / rho measurements and the {energy, momentum, rho} table of Analyze (AnalyzeWithoutRho for charged sectors).
"""
import argparse
import os
import sys
import time

import numpy as np

from . import api as cb
from . import models


def parse_couplings(s):
    """ParseCouplings, src/dmrg.h:369-382: comma list, |x| < 1e-5 -> 0."""
    out = []
    for tok in s.split(","):
        v = float(tok)
        out.append(0.0 if abs(v) < 1e-5 else v)
    return out


def file_stem(site, bc, n, maxdim, cutoff, precision, coupling_str, u, charge):
    """src/dmrg.h:360: format("%s_%s_%d_%d_%g_%g_%g_%s_%d", ...) -- tinyformat prints the coupling string under %g and u under %s."""
    return "%s_%s_%d_%d_%g_%g_%s_%g_%d" % (site, bc, n, maxdim, cutoff, precision, coupling_str, u, charge)


def dump_energy(path, state_order, en):
    """DumpEnergy, src/utility.cpp:43-53."""
    with open(path, "a") as f:
        f.write("{%d,%.10f},\n" % (state_order, en))


def dump_ee(path, n, svns):
    """DumpEE, src/utility.cpp:30-40."""
    with open(path, "a") as f:
        f.write("{\n")
        for b in range(1, n):
            f.write("{%d,  %.10f},\n" % (b, svns[b - 1]))
        f.write("},\n")


def dump_matrix(path, rows):
    """DumpMatrix, src/utility.cpp:81-113 (.an, overwritten)."""
    with open(path, "w") as f:
        f.write("{\n")
        f.write(",\n".join("{" + ",".join("%.10f" % v for v in r) + "}" for r in rows))
        f.write("\n}\n")


def dump_mathematica(path, name, M):
    """DumpMathematicaSingle, src/utility.cpp:56-78 (.m, appended)."""
    M = np.asarray(M, dtype=complex)
    n = M.shape[0]
    with open(path, "a") as f:
        f.write("%s = {\n" % name)
        for i in range(n):
            f.write("{" + ",  ".join("%.10f + I * %.10f" % (M[i, j].real, M[i, j].imag) for j in range(n)))
            f.write("},\n" if i != n - 1 else "}\n};\n\n")


def analyze_without_rho(energies, T, n):
    """AnalyzeWithoutRho, src/dmrg.h:1177-1218: m = E/L + T; eigen(m); rotate E and T into its eigenbasis;
    momentum_i = round(Re(log(T_ii) / (2 pi i) * L) * 1e5) / 1e5 with -L/2 -> L/2; rows {energy, momentum} sorted by energy.
    (Host-side post-processing of an n_states x n_states matrix, as in the reference.)"""
    E = np.diag(np.asarray(energies, dtype=complex))
    w, U = np.linalg.eig(E / n + np.asarray(T, dtype=complex))
    Ui = np.linalg.inv(U)
    Er, Tr = Ui @ E @ U, Ui @ T @ U
    rows = []
    for i in range(len(energies)):
        mom = np.round((np.log(Tr[i, i]) / (2j * np.pi) * n).real * 1e5) / 1e5
        if mom == -(n // 2):
            mom = n // 2
        rows.append((float(Er[i, i].real), float(mom)))
    return sorted(rows)


def state_name(k):
    return {0: "1st state", 1: "2nd state", 2: "3rd state"}.get(k, "%dth state" % (k + 1))


class Simulation:
    """SimulateRaw (src/dmrg.h:474-655) with the same member names where they exist."""

    def __init__(self, site, bc, n, maxdim, cutoff, precision, coupling_str, u, charge, nstates, out=".", seed=None, ctx=None,
                 log=print):
        if bc not in ("p", "o", "s", "d", "sp"):
            raise ValueError("Invalid boundary condition")
        self.site_type, self.bc, self.n = site, bc, n
        self.max_bond_dim = int(maxdim)
        # -c, -p, -u are parsed as float and widened (src/dmrg.cpp:15-18, src/dmrg.h:321-326)
        self.svd_cutoff = float(np.float32(cutoff))
        self.precision = float(np.float32(precision))
        self.u = float(np.float32(u))
        self.coupling_str = coupling_str
        self.couplings = parse_couplings(coupling_str)
        self.charge, self.num_states = charge, nstates
        self.noise, self.orthogonal_weight = 0.0, 1000.0             # src/dmrg.h:350-351
        self.init_num_sweeps_per_rep, self.num_sweeps_per_rep, self.num_reps_til_stable = 5, 1, 3   # :354-356
        self.sites = models.make_siteset(site, n)
        self.ctx = ctx
        self.rng = np.random.default_rng(seed)
        self.log = log
        stem = file_stem(site, bc, n, self.max_bond_dim, self.svd_cutoff, self.precision, coupling_str, self.u, charge)
        self.stem = stem
        for sub in ("ee", "en", "m", "an"):
            os.makedirs(os.path.join(out, sub), exist_ok=True)
        self.ee_path = os.path.join(out, "ee", stem + ".ee")
        self.en_path = os.path.join(out, "en", stem + ".en")
        self.m_path = os.path.join(out, "m", stem + ".m")
        self.an_path = os.path.join(out, "an", stem + ".an")
        self.energies, self.states, self.sweep_log = [], [], []

    def hamiltonian(self, bc):
        mpo = self.sites.hamiltonian(bc, self.n, self.u, self.couplings)
        return cb.MPO(mpo.sites, mpo.link_charges, self.sites.charges, nq=self.sites.nq)

    def default_init_state(self):
        dt = complex if self._complex else float
        return self.sites.init_state(self.charge, self.rng, dtype=dt)

    def _dmrg(self, H, done, psi, sweeps):
        t0 = time.perf_counter()
        info = cb.DmrgInfo()
        en, psi = cb.dmrg(H, done, psi, sweeps, weight=self.orthogonal_weight, ctx=self.ctx, info=info)
        self.sweep_log.append(dict(nsweep=sweeps.nsweep, seconds=time.perf_counter() - t0, energy=en, maxdim=max(psi.bond_dims()),
                                   timers=info.timers))
        return en, psi

    def run(self):
        n, c = self.n, self.svd_cutoff
        H = self.hamiltonian(self.bc)
        self._complex = H.is_complex()
        init_state = self.default_init_state()
        hot_start = 1 if self.bc == "sp" else -1
        done = []                    # dmrg_progress_.DoneStates()
        while True:
            name = state_name(len(done))
            self.log("\n> Simulation: %s" % name)
            min_en, cnt, num_sweeps = 1000000.0, self.num_reps_til_stable, 0          # ResetDMRG
            bond_dim = 200
            if hot_start != 0:       # warm-up ladder, src/dmrg.h:538-547
                sw = cb.Sweeps(self.init_num_sweeps_per_rep, [10, 20, 40, 80, 200],
                               [max(1e-5, c), max(1e-6, c), max(1e-7, c), max(1e-8, c), max(1e-9, c)], niter=2, noise=self.noise)
            else:                    # SSD -> PBC re-warm-up, :548-567
                sw = cb.Sweeps(self.init_num_sweeps_per_rep, bond_dim, c, niter=2, noise=self.noise)
            en, psi = self._dmrg(H, done, init_state, sw)
            num_sweeps += self.init_num_sweeps_per_rep
            while True:              # :576-609
                if len(done) == 0:   # IsSimulatingFirstState
                    dump_ee(self.ee_path, n, cb.calc_ee(psi, self.sites.charges, ctx=self.ctx))
                if bond_dim <= self.max_bond_dim - 200:
                    bond_dim += 200
                en, psi = self._dmrg(H, done, psi, cb.Sweeps(self.num_sweeps_per_rep, bond_dim, c, niter=2, noise=self.noise))
                if abs(min_en - en) * n < self.precision:
                    cnt -= 1
                else:
                    cnt, min_en = self.num_reps_til_stable, en
                num_sweeps += self.num_sweeps_per_rep
                self.log("  > Sweep #%d  %s  E=%.12f  bond_dim=%d (reached %d)" % (num_sweeps, name, en, bond_dim, max(psi.bond_dims())))
                if cnt == 0:
                    break
            if hot_start == 1:       # :613-621
                H = self.hamiltonian("p")
                init_state = psi
                done = []
                hot_start = 0
                ndone_after = 0      # the progress object was replaced before NextState()
            else:
                if hot_start == 0:
                    init_state = self.default_init_state()
                    hot_start = -1
                done.append(psi)
                self.states.append(psi)
                self.energies.append(en)
                ndone_after = len(done)
            # DumpEnergy(NumDoneStates()+1, en) is called AFTER NextState(): the first state is written with index 2 (:625-633)
            dump_energy(self.en_path, ndone_after + 1, en)
            self.log("\n> Energy of %s: %g\n" % (name, en))
            if ndone_after >= self.num_states:
                break
        return self.energies, self.states


def analyze_full(energies, T, R, n, bc):
    """Analyze, src/dmrg.h:1118-1174: m = E/L + T + R (periodic) | E/L + R (Dirichlet); eigen(m); rotate E, T, R into its eigenbasis;
    momentum_i = round(Re(log(T_ii)/(2 pi i) L) 1e6)/1e6 with -L/2 -> L/2, rho_i = round(Re R_ii 1e6)/1e6; rows sorted by energy."""
    E = np.diag(np.asarray(energies, dtype=complex))
    m = E / n + (np.asarray(T) + np.asarray(R) if bc == "p" else np.asarray(R))
    w, U = np.linalg.eig(m)
    Ui = np.linalg.inv(U)
    Er, Tr, Rr = Ui @ E @ U, Ui @ T @ U, Ui @ R @ U
    rows = []
    for i in range(len(energies)):
        mom = np.round((np.log(Tr[i, i]) / (2j * np.pi) * n).real * 1e6) / 1e6
        if mom == -(n // 2):
            mom = n // 2
        rho = np.round(Rr[i, i].real * 1e6) / 1e6 + 0.0           # "+ 0.0": -0 -> 0 as in the reference
        rows.append((float(Er[i, i].real), float(mom), float(rho)))
    return sorted(rows)


def rho_matrix(sim, states):
    """Measure("rho"), src/dmrg.h:871-1047 + :1068-1072: R_ij = <psi_i| rho |psi_j>.  The reference applies RhoOp approximately
    (applyMPO "Fit", cutoff 1e-5) and then takes the overlap; here the matrix element itself is contracted on the device
    (cb200_mps_mpo_element) with the ring operator written as a sum of d^2-bond MPOs (models.rho_defect_mpos)."""
    if sim.site_type == "haagerupq":
        states = [models.to_haagerup_basis(p) for p in states]
    mpos = [cb.MPO(m, None, None, nq=1) for m in models.rho_defect_mpos(sim.site_type, sim.n, periodic=(sim.bc in ("p", "sp")))]
    return np.array([[sum(cb.mpo_element(a, O, b, ctx=sim.ctx) for O in mpos) for b in states] for a in states])


def analyze(sim, log=print):
    """What mode 5 runs after the simulation (src/dmrg.h:637-648): Measure("energy") rewrites .en with 1-based state numbers
    (:752-755) and dumps the energy matrix to .m; periodic chains get Measure("translation") (L-1 swaps per state on the device,
    kCutoff = 1e-5; T_ij = innerC(psi_i, T psi_j)); charge-0 periodic / Dirichlet chains get Measure("rho") and the
    {energy, momentum, rho} table of Analyze; charged sectors the {energy, momentum} table of AnalyzeWithoutRho; every other
    boundary condition Energies()."""
    n, ens, states = sim.n, sim.energies, sim.states
    bc = "p" if sim.bc == "sp" else sim.bc
    if os.path.exists(sim.en_path):
        os.remove(sim.en_path)
    for i, e in enumerate(ens):
        dump_energy(sim.en_path, i + 1, e)
    if os.path.exists(sim.m_path):
        os.remove(sim.m_path)
    dump_mathematica(sim.m_path, "energy", np.diag(np.asarray(ens, dtype=complex)))
    if bc not in ("p", "d"):
        rows = [(e,) for e in ens]                               # Energies(), src/dmrg.h:1221-1244
        dump_matrix(sim.an_path, rows)
        return rows
    sq = sim.sites.charges
    nst = len(states)
    if bc == "p":
        acted = [cb.translate(p, sq, 1e-5, ctx=sim.ctx) for p in states]
        T = np.array([[cb.overlap(a, t, sq, ctx=sim.ctx) for t in acted] for a in states])
    else:
        T = np.eye(nst, dtype=complex)
    dump_mathematica(sim.m_path, "translation", T)
    if sim.charge == 0:
        R = rho_matrix(sim, states)
        dump_mathematica(sim.m_path, "rho", R)
        rows = analyze_full(ens, T, R, n, bc)
        head = "{energy, momentum, rho}"
    else:
        rows = analyze_without_rho(ens, T, n)
        head = "{energy, momentum}"
    dump_matrix(sim.an_path, rows)
    log("\nSpectrum:  %s" % head)
    for r in rows:
        log("{%s}" % ",".join("%.10f" % v for v in r))
    return rows


def main(argv=None):
    ap = argparse.ArgumentParser(prog="chain_b200.driver", description="Simulates DMRG on anyon chains (stand-in for ./Chain).", add_help=False)
    ap.add_argument("-s", "--site", required=True, choices=["golden", "haagerup", "haagerupq"])
    ap.add_argument("-b", "--bc", required=True)
    ap.add_argument("-l", "--length", type=int, required=True)
    ap.add_argument("-d", type=int, required=True, dest="maxdim")
    ap.add_argument("-c", "--cutoff", type=float, required=True)
    ap.add_argument("-p", "--precision", type=float, required=True)
    ap.add_argument("-j", "--couplings", required=True)
    ap.add_argument("-u", "--penalty", type=float, required=True)
    ap.add_argument("-q", "--charge_", type=int, required=True, dest="charge")
    ap.add_argument("-n", "--nstates", type=int, required=True)
    ap.add_argument("-m", "--mode", type=int, required=True)
    ap.add_argument("-h", "--help", action="help")
    ap.add_argument("--out", default=".")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--lib", default=None, help="path of the C-ABI library (default: chain_b200/libchain_b200.so)")
    a = ap.parse_args(argv)
    if a.mode not in (0, 5):
        sys.exit("mode %d is not available in the stand-in driver: 1/2/3/4/6/7 need the measurement operators of src/chain.cpp "
                 "(SURVEY 8f row f2) or ITensor's checkpoint files; the hot path is modes 0 and 5" % a.mode)
    ctx = cb.Context(device=a.device, lib_path=a.lib)
    sim = Simulation(a.site, a.bc, a.length, a.maxdim, a.cutoff, a.precision, a.couplings, a.penalty, a.charge, a.nstates,
                     out=a.out, seed=a.seed, ctx=ctx)
    print("\n> %s %s:  L=%d  bond_dim=%d  svd_cutoff=%g  precision=%g  penalty=%g  couplings=%s  charge=%d" % (
        a.site, a.bc, a.length, a.maxdim, sim.svd_cutoff, sim.precision, sim.u, a.couplings, a.charge))
    t0 = time.perf_counter()
    ens, _ = sim.run()
    wall = time.perf_counter() - t0
    nsw = sum(s["nsweep"] for s in sim.sweep_log)
    print("> %d states, %d sweeps, %.2f s in dmrg() (%.3f s/sweep); energies: %s" % (
        len(ens), nsw, sum(s["seconds"] for s in sim.sweep_log), sum(s["seconds"] for s in sim.sweep_log) / max(nsw, 1),
        ", ".join("%.10f" % e for e in ens)))
    print("> wrote %s and %s (wall %.2f s)" % (sim.en_path, sim.ee_path, wall))
    if a.mode == 5:
        analyze(sim)
        print("> wrote %s and %s" % (sim.m_path, sim.an_path))
    return 0


if __name__ == "__main__":
    sys.exit(main())
