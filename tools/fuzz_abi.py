"""Misuse fuzz of the C ABI (include/chain_b200.h) under AddressSanitizer + UBSan: every call of a caller that gets a struct field wrong must
come back with a status code -- no crash, no out-of-bounds access, no hang -- and leave the context usable.  The sources under test are the
product's own host side (capi.cpp / engine_io.cpp / engine.cpp: argument validation lives there, in front of any kernel), linked against the
host-emulation backend and instrumented (tools/asan_engine.py builds it under /tmp).

A valid `cb200_dmrg` / `cb200_mps_entropies` / `cb200_mps_overlap` / `cb200_mps_mpo_element` / `cb200_mps_translate` call is built for a dense (golden)
and a Z3 (haagerupq, complex) chain; one field at a time is then corrupted: counts (n, nq, d, nprev, nsweep) zero / negative / mismatched, tensor rank,
dtype, dims zero / negative / inconsistent with the neighbour or with the link index, charge labels out of range, NULL site / link / data / charge /
sweep / output pointers, ortho_center and layout out of range, sweep parameters (maxdim, mindim, niter <= 0, cutoff / noise / weight negative or NaN),
NaN / Inf / flux-violating tensor data, an all-zero start state.  Tensor buffers are over-allocated so that a corrupted dimension never makes a
CORRECT library read outside what the caller owns.  Each case runs in the instrumented child with an alarm; the parent attributes a crash or a
sanitizer report to the case that was running.

    python tools/fuzz_abi.py            # prints one line per failing case and `cases: N  failures: 0`
    python tools/fuzz_abi.py --quick    # every 5th case on the un-instrumented host-emulation library (tests/test_fuzz.py; ~15 s)
"""
import ctypes as C
import os
import signal
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

SLACK = 1 << 16          # elements every tensor buffer really has


def build_valid(kind):
    """Valid C structs (and the numpy arrays behind them, padded to SLACK elements) for one model."""
    import numpy as np
    import chain_b200 as cb
    from chain_b200 import _lib
    import common
    from oracle import dmrg as od
    st, cp = ("golden", [-1.0]) if kind == "dense" else ("haagerupq", [0.0, -1.0, 0.3])
    n = 4
    site, W, qk = common.build_model(st, "p", n, cp)
    H = cb.MPO(W, qk, site.charges, nq=site.nq)
    psi = common.warm_start(site, W, n, 0, seed=3, maxdim=6)
    mps = common.to_cb(psi)
    keep = []

    def padded(t, dt):
        f = _lib.fcol(t, dt)
        buf = np.zeros(max(SLACK, f.size), dt)
        buf[:f.size] = f.reshape(-1, order="F")
        keep.append(buf)
        return buf

    Hc, Pc = H._to_c(keep), mps._to_c(keep)
    for j in range(n):           # re-point at padded buffers
        Hc.site[j].data = padded(H.sites[j] if H.is_complex() else np.real(H.sites[j]), np.complex128 if H.is_complex() else np.float64).ctypes.data
        Pc.site[j].data = padded(mps.sites[j], np.complex128 if mps.is_complex() else np.float64).ctypes.data
    # charge arrays padded too (a corrupted link dim must stay inside them)
    for arr, nl in ((Hc.link, n + 1), (Pc.link, n + 1)):
        for j in range(nl):
            q = np.zeros(256, np.int32)
            src = np.ctypeslib.as_array(arr[j].charge, shape=(arr[j].dim,)) if arr[j].charge else np.zeros(arr[j].dim, np.int32)
            q[:arr[j].dim] = src
            keep.append(q)
            arr[j].charge = q.ctypes.data_as(C.POINTER(C.c_int32))
    sc = np.zeros(64, np.int32)
    sc[:H.d] = H.site_charges
    keep.append(sc)
    Hc.site_charge = sc.ctypes.data_as(C.POINTER(C.c_int32))
    sw = cb.Sweeps(2, [4, 8], 1e-8).to_c()
    return dict(H=Hc, psi=Pc, sw=sw, nsweep=2, keep=keep, d=H.d, nq=site.nq, n=n, sc=Hc.site_charge)


def mutations(v):
    """(name, function(ctx-independent state dict) -> None) -- each corrupts ONE thing of a fresh valid state."""
    n = v["n"]
    nan = float("nan")
    out = []

    def m(name, fn):
        out.append((name, fn))

    for who in ("H", "psi"):
        for val in (0, 1, -1, n - 1, -2 ** 31):          # (n larger than the arrays the caller owns cannot be detected by any library)
            m("%s.n=%d" % (who, val), lambda s, who=who, val=val: setattr(s[who], "n", val))
        for val in (0, -1, 2, 5, 1000):
            m("%s.nq=%d" % (who, val), lambda s, who=who, val=val: setattr(s[who], "nq", val))
        m("%s.site=NULL" % who, lambda s, who=who: setattr(s[who], "site", None))
        m("%s.link=NULL" % who, lambda s, who=who: setattr(s[who], "link", None))
        for j in (0, 1, n - 1):
            for val in (0, 1, 2, 5, -3):
                m("%s.site[%d].rank=%d" % (who, j, val), lambda s, who=who, j=j, val=val: setattr(s[who].site[j], "rank", val))
            for val in (-1, 2, 7):
                m("%s.site[%d].dtype=%d" % (who, j, val), lambda s, who=who, j=j, val=val: setattr(s[who].site[j], "dtype", val))
            m("%s.site[%d].dtype flipped" % (who, j), lambda s, who=who, j=j: setattr(s[who].site[j], "dtype", 1 - s[who].site[j].dtype))
            m("%s.site[%d].data=NULL" % (who, j), lambda s, who=who, j=j: setattr(s[who].site[j], "data", None))
            for k in range(4):
                for val in (0, -1, 1, 3, 7, 40):
                    def f(s, who=who, j=j, k=k, val=val):
                        s[who].site[j].dims[k] = val
                    m("%s.site[%d].dims[%d]=%d" % (who, j, k, val), f)
        for j in (0, 1, n):
            for val in (0, -1, 1, 3, 40, 200):
                m("%s.link[%d].dim=%d" % (who, j, val), lambda s, who=who, j=j, val=val: setattr(s[who].link[j], "dim", val))
            m("%s.link[%d].charge=NULL" % (who, j), lambda s, who=who, j=j: setattr(s[who].link[j], "charge", None))
            for val in (-1, 3, 99):
                def f(s, who=who, j=j, val=val):
                    s[who].link[j].charge[0] = val
                m("%s.link[%d].charge[0]=%d" % (who, j, val), f)
            def g(s, who=who, j=j):
                for i in range(s[who].link[j].dim):
                    s[who].link[j].charge[i] = (s[who].link[j].charge[i] + 1) % max(s[who].nq, 2)
            m("%s.link[%d].charges shifted" % (who, j), g)
    for val in (0, -1, 1, 3, 7, 64):
        m("H.d=%d" % val, lambda s, val=val: setattr(s["H"], "d", val))
    m("H.site_charge=NULL", lambda s: setattr(s["H"], "site_charge", None))
    for val in (-1, 3, 50):
        def f(s, val=val):
            s["H"].site_charge[0] = val
        m("H.site_charge[0]=%d" % val, f)
    for val in (-1, n + 1, 2 ** 30):
        m("psi.ortho_center=%d" % val, lambda s, val=val: setattr(s["psi"], "ortho_center", val))
    for val in (-1, 2, 9):
        m("psi.layout=%d" % val, lambda s, val=val: setattr(s["psi"], "layout", val))
    m("psi.layout=1 with dense data", lambda s: setattr(s["psi"], "layout", 1))
    for val in (0, -1, -2 ** 31):
        m("nsweep=%d" % val, lambda s, val=val: s.__setitem__("nsweep", val))
    m("sweeps=NULL", lambda s: s.__setitem__("sw", None))
    for fld, vals in (("maxdim", (0, -1, -2 ** 31)), ("mindim", (0, -1, 100, 2 ** 31 - 1)), ("niter", (0, -1, 1, 50)),
                      ("cutoff", (-1.0, nan, float("inf"), 2.0, 0.0, 1e-300)), ("noise", (nan, -1.0, 1e-3))):
        for val in vals:
            for i in (0, 1):
                m("sweeps[%d].%s=%r" % (i, fld, val), lambda s, fld=fld, val=val, i=i: setattr(s["sw"][i], fld, val))
    for val in (nan, float("inf"), -5.0, 0.0, 1e300):
        m("weight=%r" % val, lambda s, val=val: s.__setitem__("weight", val))
    for val in (-1, 1, 3):
        m("nprev=%d with prev=NULL" % val, lambda s, val=val: (s.__setitem__("nprev", val), s.__setitem__("prev", None)))
    m("nprev=1 (psi itself)", lambda s: (s.__setitem__("nprev", 1), s.__setitem__("prev", C.pointer(s["psi"]))))
    m("energy=NULL", lambda s: s.__setitem__("energy_null", True))
    m("out=NULL", lambda s: s.__setitem__("out_null", True))
    m("ctx=NULL", lambda s: s.__setitem__("ctx_null", True))
    m("H=NULL", lambda s: s.__setitem__("H_null", True))
    m("psi0=NULL", lambda s: s.__setitem__("psi_null", True))

    def poke(who, j, idx, val):
        def f(s):
            t = s[who].site[j]
            p = C.cast(t.data, C.POINTER(C.c_double))
            p[idx] = val
        return f
    for who in ("H", "psi"):
        for val, nm in ((nan, "NaN"), (float("inf"), "Inf"), (1e308, "1e308")):
            m("%s.site[1].data[0]=%s" % (who, nm), poke(who, 1, 0, val))
    # flux violation: fill the whole tensor with ones (forbidden blocks become non-zero for Z3)
    def ones(who):
        def f(s):
            for j in range(n):
                t = s[who].site[j]
                cnt = 1
                for k in range(t.rank):
                    cnt *= t.dims[k]
                p = C.cast(t.data, C.POINTER(C.c_double))
                for i in range(cnt * (2 if t.dtype == 1 else 1)):
                    p[i] = 1.0
        return f
    m("psi data all ones (flux-violating for Z3)", ones("psi"))
    m("H data all ones (flux-violating for Z3)", ones("H"))

    def zeros(s):
        for j in range(n):
            t = s["psi"].site[j]
            cnt = 1
            for k in range(t.rank):
                cnt *= t.dims[k]
            C.memset(t.data, 0, cnt * (16 if t.dtype == 1 else 8))
    m("psi data all zero", zeros)
    return out


def call_all(lib, ctx, s):
    """Every MPS/MPO-taking entry point with the (possibly corrupted) state; returns the status codes."""
    from chain_b200._lib import MpsC
    energy, res = C.c_double(), C.c_void_p()
    h = None if s.get("ctx_null") else ctx.h
    Hp = None if s.get("H_null") else C.byref(s["H"])
    pp = None if s.get("psi_null") else C.byref(s["psi"])
    prev = s.get("prev", (MpsC * 1)())
    rcs = []
    rc = lib.cb200_dmrg(h, Hp, prev, s.get("nprev", 0), pp, s["sw"], s["nsweep"], s.get("weight", 1000.0),
                        None if s.get("energy_null") else C.byref(energy), None if s.get("out_null") else C.byref(res))
    rcs.append(rc)
    if rc == 0 and res.value:
        lib.cb200_result_free(res)
    n = max(min(s["psi"].n, 64), 2)
    ee = (C.c_double * 64)()
    if not s.get("out_null"):
        rcs.append(lib.cb200_mps_entropies(h, pp, s["H"].d if Hp is not None else 2, s["H"].site_charge if Hp is not None else None, ee) if n <= 64 else 0)
    ov = (C.c_double * 2)()
    rcs.append(lib.cb200_mps_overlap(h, pp, pp, s["H"].d, s["H"].site_charge, ov))
    rcs.append(lib.cb200_mps_mpo_element(h, pp, Hp, pp, ov))
    tr = C.c_void_p()
    rc = lib.cb200_mps_translate(h, pp, s["H"].d, s["H"].site_charge, 1e-5, C.byref(tr))
    rcs.append(rc)
    if rc == 0 and tr.value:
        lib.cb200_result_free(tr)
    return rcs


def child():
    import chain_b200 as cb
    if os.environ.get("CB200_FUZZ_QUICK") == "1":
        from conftest import EMU_LIB as LIB
    else:
        from asan_engine import LIB
    stride = 5 if os.environ.get("CB200_FUZZ_QUICK") == "1" else 1
    ctx = cb.Context(lib_path=LIB)
    lib = ctx.lib
    total = 0
    for kind in ("dense", "z3"):
        base = build_valid(kind)
        rc0 = call_all(lib, ctx, dict(base))
        want0 = [0, 0, 0, 0 if kind == "dense" else -2, 0]          # cb200_mps_mpo_element takes dense (nq = 1) operands only
        assert rc0 == want0, ("the valid call must succeed", kind, rc0, lib.cb200_last_error(ctx.h))
        names = [nm for nm, _ in mutations(base)]
        only = os.environ.get("CB200_FUZZ_ONLY")
        skip = set(filter(None, os.environ.get("CB200_FUZZ_SKIP", "").split(",")))      # cases that crashed a previous child
        for i, nm in enumerate(names):
            if (only and only != "%s:%d" % (kind, i)) or "%s:%d" % (kind, i) in skip or i % stride:
                continue
            s = build_valid(kind)                     # fresh structs and buffers for every case
            fn = dict(mutations(s))[nm]
            print("CASE %s:%d %s" % (kind, i, nm), flush=True)
            fn(s)
            signal.alarm(60)
            rcs = call_all(lib, ctx, s)
            signal.alarm(0)
            print("RC %s" % rcs, flush=True)
            total += 1
        again = call_all(lib, ctx, build_valid(kind))  # the context survived all of it
        assert again == want0, ("context unusable after the misuse", kind, again)
    print("DONE %d" % total, flush=True)


def main():
    if "--quick" in sys.argv:
        env = dict(os.environ, CB200_FUZZ_CHILD="1", CB200_FUZZ_QUICK="1")
    else:
        import asan_engine
        asan_engine.build()
        gcc = lambda f: subprocess.check_output(["gcc", "-print-file-name=" + f], text=True).strip()
        env = dict(os.environ, CB200_FUZZ_CHILD="1", LD_PRELOAD=gcc("libasan.so") + " " + gcc("libstdc++.so.6"),
                   ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    failures, cases, skip = [], 0, set()
    # a crash ends the child; restart it past the offending case until the list is exhausted
    for attempt in range(200):
        env["CB200_FUZZ_SKIP"] = ",".join(sorted(skip))
        p = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        lines = p.stdout.splitlines()
        done = any(l.startswith("DONE") for l in lines)
        last = [l for l in lines if l.startswith("CASE ")]
        if done and p.returncode == 0:
            cases = int([l for l in lines if l.startswith("DONE")][0].split()[1]) + len(skip)
            # corrupted calls that every entry point still accepted: each must be a corruption that leaves a VALID problem
            acc = [lines[i][5:] for i in range(len(lines) - 1) if lines[i].startswith("CASE ") and lines[i + 1] in ("RC [0, 0, 0, 0, 0]", "RC [0, 0, 0, -2, 0]")]
            print("accepted by every entry point (%d):" % len(acc))
            for a in acc:
                print("   ", a)
            break
        culprit = last[-1] if last else "(before the first case)"
        if last and lines.index(last[-1]) + 1 < len(lines):      # the last case did return: the child died later (an assertion of this script)
            print("child failed outside a case:\n%s" % (p.stderr or "")[-1500:])
            failures.append(("(outside a case)", p.returncode, ""))
            break
        err = (p.stderr or "").splitlines()
        top = [l for l in err if "ERROR" in l or "runtime error" in l or "csrc/" in l or "hostemu/" in l or "Error" in l][:8]
        failures.append((culprit, p.returncode, top))
        print("FAIL %s rc=%d\n    %s" % (culprit, p.returncode, "\n    ".join(t.strip()[:230] for t in top)), flush=True)
        if not last:
            break
        skip.add(culprit.split()[1])
    print("cases: %d  failures: %d" % (cases, len(failures)))
    return 1 if failures else 0


if __name__ == "__main__":
    if os.environ.get("CB200_FUZZ_CHILD") == "1":
        child()
    else:
        sys.exit(main())
