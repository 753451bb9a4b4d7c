"""Stand-in for the reference's driver (/root/reference/src/dmrg.cpp + DMRG<SiteSet>::SimulateRaw, src/dmrg.h:474-655).

The real driver stays unchanged and reaches this library through the adapter of INTEGRATION.md; it cannot be compiled
where ITensor is absent, so this module restates the part of it that surrounds the hot path -- the CLI flags
(src/dmrg.cpp:10-23), the warm-up / ramp / stopping schedule (src/dmrg.h:527-609), the file names (src/dmrg.h:360-365)
and the .en / .ee writers (src/utility.cpp:30-53) -- and calls cb200_dmrg for every `dmrg(...)` of the reference.
Nothing here does tensor arithmetic: Hamiltonians come from chain_b200.models (host, O(L)), DMRG and CalcEE run on the GPU.

    python -m chain_b200.driver -s golden -b p -l 6 -d 100 -c 1e-8 -p 1e-9 -j -1 -u 2 -q 0 -n 1 -m 0 [--out DIR] [--seed N]

Modes as in src/dmrg.cpp:44-60: 0 Simulate, 5 Simulate + analysis, 1 / 2 / 3 Analyze / AnalyzeWithoutRho / Energies and 4 NormalizeEnergies
on the progress file of an earlier run, 6 Repair (RecoverStates), 7 Test (exact levels of the Hamiltonian MPO, host-side as in the reference).

Differences from the reference, all deliberate: output goes under --out (the reference compiles the parent of the cwd in,
src/path_template.h:8); --seed makes the random product start reproducible (the reference never seeds); the pgs/ checkpoint
has DMRGProgress's semantics (src/dmrg.h:15-257: written after every dmrg() call, a finished job is recognised, an interrupted
one resumes at the sweep it stopped before) but its own container -- one NumPy .npz per job instead of ITensor's private binary
streams, SURVEY 8f row f4 -- and also stores the stopping-rule counters the reference leaves uninitialised on resume; mode 5
runs the simulation and then the energy / translation / rho measurements and the {energy, momentum, rho} table of Analyze
(AnalyzeWithoutRho for charged sectors).
"""
import argparse
import io
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


class DMRGProgress:
    """DMRGProgress<SiteSet>, src/dmrg.h:15-257, with the same member names: per state the latest (times swept, bond dimension,
    energy, MPS); the last entry is the state being simulated, everything before it is done.  The Hamiltonian is not stored (the
    reference keeps a copy per state, :24; it is a pure function of the job parameters, which are in the file name) -- `stage`
    records which Hamiltonian a `-b sp` job is on instead (hot_start of src/dmrg.h:521-524), and (min_en, cnt) the stopping rule.
    File: <stem>.pgs = one .npz, written to a temporary and renamed so that a kill during the write cannot corrupt it (the
    reference answers that case with .bk copies and RecoverStates, :164-199); site tensors are stored without their forbidden
    charge blocks."""

    def __init__(self):
        self.num_sweeps_vec_, self.max_dims_, self.ens_, self.psis_ = [], [], [], []
        self.stage, self.min_en, self.cnt = None, 1000000.0, 0
        self.sp_seed = None          # `-b sp` only: the SSD result that seeds the PBC re-warm-up (init_state_ of src/dmrg.h:616)

    def Update(self, times_swept, max_dim, en, psi):
        self.num_sweeps_vec_[-1], self.max_dims_[-1], self.ens_[-1], self.psis_[-1] = int(times_swept), int(max_dim), float(en), psi

    def NextState(self):
        self.num_sweeps_vec_.append(0)
        self.max_dims_.append(0)
        self.ens_.append(0.0)
        self.psis_.append(None)

    def All(self):
        return self.num_sweeps_vec_[-1], self.max_dims_[-1], self.ens_[-1], self.psis_[-1]

    def DoneStates(self):
        return list(self.psis_[:-1])

    def Energies(self):
        return list(self.ens_[:-1])

    def NumDoneStates(self):
        return len(self.psis_) - 1

    @staticmethod
    def _mask(psi, j, site_charges):
        ql, qr = np.asarray(psi.link_charges[j]), np.asarray(psi.link_charges[j + 1])
        return (ql[:, None, None] + np.asarray(site_charges)[None, :, None] - qr[None, None, :]) % psi.nq == 0

    def Write(self, path, site_charges, backup=False):
        z = {"num_sweeps_vec": np.asarray(self.num_sweeps_vec_, np.int64), "max_dims": np.asarray(self.max_dims_, np.int64),
             "ens": np.asarray(self.ens_, np.float64), "stage": np.asarray(-2 if self.stage is None else self.stage, np.int64),
             "min_en": np.asarray(self.min_en), "cnt": np.asarray(self.cnt, np.int64), "site_charges": np.asarray(site_charges, np.int32),
             "present": np.asarray([p is not None for p in self.psis_], np.int8)}
        for i, psi in list(enumerate(self.psis_)) + [("eed", self.sp_seed)]:
            if psi is None:
                continue
            z["s%s_meta" % i] = np.asarray([psi.n, psi.nq, psi.center or 0, int(psi.is_complex())], np.int64)
            for j in range(psi.n + 1):
                z["s%s_q%d" % (i, j)] = np.asarray(psi.link_charges[j], np.int32)
            for j in range(psi.n):
                z["s%s_a%d" % (i, j)] = np.asarray(psi.sites[j])[self._mask(psi, j, site_charges)]
        buf = io.BytesIO()
        np.savez(buf, **z)
        dst = path + (".pgs.bk" if backup else ".pgs")
        with open(dst + ".tmp", "wb") as f:
            f.write(buf.getvalue())
            f.flush()
            os.fsync(f.fileno())
        os.replace(dst + ".tmp", dst)

    def Read(self, path, backup=False):
        with np.load(path + (".pgs.bk" if backup else ".pgs")) as z:
            self.num_sweeps_vec_ = [int(v) for v in z["num_sweeps_vec"]]
            self.max_dims_ = [int(v) for v in z["max_dims"]]
            self.ens_ = [float(v) for v in z["ens"]]
            st = int(z["stage"])
            self.stage, self.min_en, self.cnt = (None if st == -2 else st), float(z["min_en"]), int(z["cnt"])
            sq = z["site_charges"]
            def load(i):
                n, nq, center, cplx = (int(v) for v in z["s%s_meta" % i])
                q = [z["s%s_q%d" % (i, j)] for j in range(n + 1)]
                sites = []
                for j in range(n):
                    m = (q[j][:, None, None] + sq[None, :, None] - q[j + 1][None, None, :]) % nq == 0
                    a = np.zeros(m.shape, complex if cplx else float)
                    a[m] = z["s%s_a%d" % (i, j)]
                    sites.append(a)
                return cb.MPS(sites, q, nq=nq, center=center)
            self.psis_ = [load(i) if present else None for i, present in enumerate(z["present"])]
            self.sp_seed = load("eed") if "seed_meta" in z.files else None
        if not (len(self.num_sweeps_vec_) == len(self.max_dims_) == len(self.ens_) == len(self.psis_) >= 1):
            raise IndexError("corrupt progress file %s" % path)      # the std::out_of_range Simulate catches, src/dmrg.h:454


class Simulation:
    """SimulateRaw (src/dmrg.h:474-655) with the same member names where they exist."""

    def __init__(self, site, bc, n, maxdim, cutoff, precision, coupling_str, u, charge, nstates, out=".", seed=None, ctx=None,
                 log=print, checkpoint=True):
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
        self.checkpoint = bool(checkpoint)
        for sub in ("ee", "en", "m", "an") + (("pgs",) if checkpoint else ()):
            os.makedirs(os.path.join(out, sub), exist_ok=True)
        self.ee_path = os.path.join(out, "ee", stem + ".ee")
        self.en_path = os.path.join(out, "en", stem + ".en")
        self.m_path = os.path.join(out, "m", stem + ".m")
        self.an_path = os.path.join(out, "an", stem + ".an")
        self.progress_path = os.path.join(out, "pgs", stem)       # progress_path_, src/dmrg.h:361
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

    def load_progress(self):
        """What every analysis mode starts with (Measure, src/dmrg.h:733-741): the finished states of the progress file."""
        prog = self.dmrg_progress_ = DMRGProgress()
        prog.Read(self.progress_path)
        self.energies, self.states = prog.Energies(), prog.DoneStates()
        return prog.NumDoneStates()

    def run(self):
        """SimulateRaw, src/dmrg.h:474-655, including its progress file: every dmrg() call is followed by Update + Write, a job whose
        file already holds the requested number of states returns at once, an interrupted job continues with the sweep it was
        stopped before (same energies as the uninterrupted run: the stopping-rule counters are part of the file)."""
        n, c = self.n, self.svd_cutoff
        prog = self.dmrg_progress_ = DMRGProgress()
        sq = self.sites.charges
        ck = self.checkpoint
        save = (lambda backup=False: prog.Write(self.progress_path, sq, backup)) if ck else (lambda backup=False: None)
        hot_start = 1 if self.bc == "sp" else -1
        if ck and os.path.exists(self.progress_path + ".pgs"):
            try:
                prog.Read(self.progress_path)                # job was run before, read progress (:489-491)
            except Exception as err:                         # Simulate's answer to a corrupt file: the backup copy (:454-462)
                self.log("%s\nReplace progress files with backup progress files." % err)
                prog.Read(self.progress_path, backup=True)
                save()
            if prog.stage is not None:
                hot_start = prog.stage
            if prog.NumDoneStates() >= self.num_states:
                self.log("\n> Job already complete\n> Requested number of states: %d \n> Completed number of states: %d\n" % (
                    self.num_states, prog.NumDoneStates()))
                self.energies, self.states = prog.Energies(), prog.DoneStates()
                return self.energies, self.states
        else:                                                # new job (:513-517)
            prog.NextState()
            prog.stage = hot_start
            save()
        H = self.hamiltonian("p" if (self.bc == "sp" and hot_start != 1) else self.bc)
        self._complex = H.is_complex()
        init_state = self.default_init_state()
        if hot_start == 0 and prog.sp_seed is not None:      # resumed between the SSD stage and the PBC re-warm-up
            init_state = prog.sp_seed
        while True:
            name = state_name(prog.NumDoneStates())
            self.log("\n> Simulation: %s" % name)
            if prog.num_sweeps_vec_[-1] == 0:
                prog.min_en, prog.cnt, num_sweeps = 1000000.0, self.num_reps_til_stable, 0          # ResetDMRG
                bond_dim = 200
                if hot_start != 0:       # warm-up ladder, src/dmrg.h:538-547
                    sw = cb.Sweeps(self.init_num_sweeps_per_rep, [10, 20, 40, 80, 200],
                                   [max(1e-5, c), max(1e-6, c), max(1e-7, c), max(1e-8, c), max(1e-9, c)], niter=2, noise=self.noise)
                else:                    # SSD -> PBC re-warm-up, :548-567
                    sw = cb.Sweeps(self.init_num_sweeps_per_rep, bond_dim, c, niter=2, noise=self.noise)
                en, psi = self._dmrg(H, prog.DoneStates(), init_state, sw)
                num_sweeps += self.init_num_sweeps_per_rep
                prog.Update(num_sweeps, bond_dim, en, psi)
                save()
            # else: warm-up previously completed, no more warm-up (:568-572)
            while prog.cnt > 0:      # :576-609 (a job killed right after its last rep was saved resumes past the loop)
                num_sweeps, bond_dim, en, psi = prog.All()
                if prog.NumDoneStates() == 0:   # IsSimulatingFirstState
                    dump_ee(self.ee_path, n, cb.calc_ee(psi, sq, ctx=self.ctx))
                if bond_dim <= self.max_bond_dim - 200:
                    bond_dim += 200
                en, psi = self._dmrg(H, prog.DoneStates(), psi, cb.Sweeps(self.num_sweeps_per_rep, bond_dim, c, niter=2, noise=self.noise))
                if abs(prog.min_en - en) * n < self.precision:
                    prog.cnt -= 1
                else:
                    prog.cnt, prog.min_en = self.num_reps_til_stable, en
                num_sweeps += self.num_sweeps_per_rep
                prog.Update(num_sweeps, bond_dim, en, psi)
                save()
                self.log("  > Sweep #%d  %s  E=%.12f  bond_dim=%d (reached %d)" % (num_sweeps, name, en, bond_dim, max(psi.bond_dims())))
            num_sweeps, bond_dim, en, psi = prog.All()
            if hot_start == 1:       # :613-621: the SSD result seeds the PBC run and the progress object is replaced
                H = self.hamiltonian("p")
                init_state = psi
                prog = self.dmrg_progress_ = DMRGProgress()
                prog.sp_seed = psi
                save = (lambda backup=False: prog.Write(self.progress_path, sq, backup)) if ck else (lambda backup=False: None)
                hot_start = 0
            elif hot_start == 0:
                init_state = self.default_init_state()
                prog.sp_seed = None
                hot_start = -1
            prog.NextState()         # even if there is no next state, this marks completion of the current one (:623-629)
            prog.stage = hot_start
            save()
            save(backup=True)
            # DumpEnergy(NumDoneStates()+1, en) is called AFTER NextState(): the first state is written with index 2 (:625-633)
            dump_energy(self.en_path, prog.NumDoneStates() + 1, en)
            self.log("\n> Energy of %s: %g\n" % (name, en))
            if prog.NumDoneStates() >= self.num_states:
                break
        self.energies, self.states = prog.Energies(), prog.DoneStates()
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


def analyze(sim, log=print, which=None):
    """`which` = None: what mode 5 runs after the simulation (src/dmrg.h:637-648): Measure("energy") rewrites .en with 1-based state numbers
    (:752-755) and dumps the energy matrix to .m; periodic chains get Measure("translation") (L-1 swaps per state on the device,
    kCutoff = 1e-5; T_ij = innerC(psi_i, T psi_j)); charge-0 periodic / Dirichlet chains get Measure("rho") and the
    {energy, momentum, rho} table of Analyze; charged sectors the {energy, momentum} table of AnalyzeWithoutRho; every other
    boundary condition Energies().  `which` = "full" | "without_rho" | "energies": modes 1 / 2 / 3 of the CLI (Analyze,
    AnalyzeWithoutRho, Energies: src/dmrg.h:1118-1244) on the states of the progress file.  Only the first -n states are analysed
    (num_states = min(NumDoneStates, num_states_), :741), the .en file lists every finished state (:752-755)."""
    n = sim.n
    bc = "p" if sim.bc == "sp" else sim.bc
    if os.path.exists(sim.en_path):
        os.remove(sim.en_path)
    for i, e in enumerate(sim.energies):
        dump_energy(sim.en_path, i + 1, e)
    k = min(len(sim.states), sim.num_states)
    ens, states = list(sim.energies[:k]), list(sim.states[:k])
    if os.path.exists(sim.m_path):
        os.remove(sim.m_path)
    dump_mathematica(sim.m_path, "energy", np.diag(np.asarray(ens, dtype=complex)))
    if which is None:
        which = "energies" if bc not in ("p", "d") else ("full" if sim.charge == 0 else "without_rho")
    if which == "energies":
        rows = [(e,) for e in ens]                               # Energies(), src/dmrg.h:1221-1244
        log("{{%s},{%s}}," % (sim.coupling_str, ",".join("%.10f" % e for e in ens)))
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
    if which == "full":
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


def normalize_energies(sim, log=print):
    """NormalizeEnergies (mode 4), src/dmrg.h:1247-1276: (E_i - E_0) / (E_1 - E_0) of the sorted energies of the progress file."""
    ens = sorted(sim.energies[:min(len(sim.energies), sim.num_states)])
    if len(ens) <= 2:
        log("Fewer than 3 states available")
        return None
    gap = ens[1] - ens[0]
    out = [0.0, 1.0] + [(e - ens[0]) / gap for e in ens[2:]]
    log("{{%s},%g,{%s}}," % (sim.coupling_str, gap, ",".join("%.10f" % v for v in out)))
    return gap, out


def repair(sim, log=print):
    """Repair (mode 6) = DMRGProgress::RecoverStates, src/dmrg.h:164-199: re-read the progress file (the backup if the file itself does
    not load), touch every finished state (innerC(psi, psi) on the device) and write a clean file holding exactly the states that
    survived.  With the atomic .npz container a torn write cannot happen any more; what is left to repair is a missing / unreadable
    primary file."""
    prog = DMRGProgress()
    try:
        prog.Read(sim.progress_path)
    except Exception as err:
        log("%s\nReplace progress files with backup progress files." % err)
        prog.Read(sim.progress_path, backup=True)
    sq = sim.sites.charges
    good = DMRGProgress()
    good.NextState()
    log("Recover states from corrupt vector.")
    ndone = prog.NumDoneStates()
    for i in range(ndone):
        nrm = cb.overlap(prog.psis_[i], prog.psis_[i], sq, ctx=sim.ctx)
        if not np.isfinite(abs(nrm)) or abs(nrm) == 0:
            break
        log("State %d/%d is OK." % (i + 1, ndone))
        good.Update(prog.num_sweeps_vec_[i], prog.max_dims_[i], prog.ens_[i], prog.psis_[i])
        good.NextState()
    if good.NumDoneStates() == ndone and prog.psis_[-1] is not None:      # the state in progress survives as well
        good.Update(*prog.All())
        good.min_en, good.cnt = prog.min_en, prog.cnt
    good.stage = prog.stage if prog.stage is not None else (1 if sim.bc == "sp" else -1)
    good.sp_seed = prog.sp_seed
    good.Write(sim.progress_path, sq)
    good.Write(sim.progress_path, sq, backup=True)
    return good


def exact_levels(mpo_sites, k):
    """Test (mode 7), src/dmrg.h:1283-1300: the reference contracts the whole MPO into one tensor and reads the lowest levels off
    svd(exp(-H)) -- its only built-in cross-check of a Hamiltonian.  Same numbers here from the same object (the product-side MPO of
    chain_b200.models, QNs ignored): dense eigvalsh up to dimension 4096, Lanczos on the MPO-vector product above.  Host-side, like
    the reference's; this is a check of the MPO, not part of the hot path."""
    n, d = len(mpo_sites), mpo_sites[0].shape[1]
    dim = d ** n

    def matvec(x):
        x = np.asarray(x).reshape((1,) + (d,) * n)                  # [a, s_1 .. s_n], s_1 slowest
        for j, W in enumerate(mpo_sites):                           # W[a, s_out, s_in, a']
            x = np.tensordot(W, x, axes=([0, 2], [0, j + 1]))       # -> [s_out, a', rest...]
            x = np.moveaxis(x, 0, j + 1)                            # a' first again, s_out back in place j+1
        return x.reshape(dim)
    cplx = any(np.iscomplexobj(W) and np.abs(np.imag(W)).max() > 0 for W in mpo_sites)
    dt = complex if cplx else float
    if dim <= 4096:
        Hm = np.stack([matvec(np.eye(dim, dtype=dt)[:, i]) for i in range(dim)], axis=1)
        return np.linalg.eigvalsh((Hm + Hm.conj().T) / 2)[:k].tolist()
    # Lanczos (ARPACK) on the MPO-vector product; k + 4 levels are converged and the lowest k returned.  A Krylov method sees one vector
    # per degenerate level unless rounding splits it, so a multiplet can come back with fewer copies than it has -- the level VALUES are
    # exact (that is what the cross-check compares)
    from scipy.sparse.linalg import LinearOperator, eigsh
    op = LinearOperator((dim, dim), matvec=lambda v: matvec(np.asarray(v, dtype=dt).reshape(dim)), dtype=dt)
    kk = min(k + 4, dim - 2)
    w = eigsh(op, k=kk, which="SA", tol=1e-12, ncv=max(4 * kk, 48), return_eigenvectors=False, v0=np.random.default_rng(0).standard_normal(dim).astype(dt))
    return sorted(np.real(w).tolist())[:k]


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
    ap.add_argument("--no-checkpoint", action="store_true", help="do not write / read the pgs/ progress file")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--lib", default=None, help="path of the C-ABI library (default: chain_b200/libchain_b200.so)")
    a = ap.parse_args(argv)
    if a.mode == 7:                                            # Test(): needs neither a device nor a progress file
        ss = models.make_siteset(a.site, a.length)
        mpo = ss.hamiltonian(a.bc, a.length, float(np.float32(a.penalty)), parse_couplings(a.couplings))
        print("{%s}" % ",".join("%.10f" % e for e in exact_levels(mpo.sites, a.nstates)))
        return 0
    ctx = cb.Context(device=a.device, lib_path=a.lib)
    sim = Simulation(a.site, a.bc, a.length, a.maxdim, a.cutoff, a.precision, a.couplings, a.penalty, a.charge, a.nstates,
                     out=a.out, seed=a.seed, ctx=ctx, checkpoint=not a.no_checkpoint)
    print("\n> %s %s:  L=%d  bond_dim=%d  svd_cutoff=%g  precision=%g  penalty=%g  couplings=%s  charge=%d" % (
        a.site, a.bc, a.length, a.maxdim, sim.svd_cutoff, sim.precision, sim.u, a.couplings, a.charge))
    if a.mode in (1, 2, 3, 4, 6):                              # measurements and analysis on the progress file (src/dmrg.cpp:44-60)
        if not os.path.exists(sim.progress_path + ".pgs") and not (a.mode == 6 and os.path.exists(sim.progress_path + ".pgs.bk")):
            print("\n> No progress file available")
            return 0
        if a.mode == 6:
            good = repair(sim)
            print("> %d finished state(s) kept in %s.pgs" % (good.NumDoneStates(), sim.progress_path))
            return 0
        sim.load_progress()
        if a.mode == 4:
            normalize_energies(sim)
        else:
            analyze(sim, which={1: "full", 2: "without_rho", 3: "energies"}[a.mode])
            print("> wrote %s and %s" % (sim.m_path, sim.an_path))
        return 0
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
