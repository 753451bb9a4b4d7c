"""Prints the headline fields of a bench.py JSON line (used by tools/gpu_run.sh)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print("value=%.2f %s ms=%.2f e2e=%.2f n_gpus=%s frac=%s stage_ms=%s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"],
      ("%.3f" % r["frac"]) if r else None, {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}))
if d.get("roofline_hbm"):
    for k in d["roofline_hbm"]["kernels"]:
        print("  hbm %-40s %8.1f GB/s frac %.2f (%.3f ms)" % (k["kernel"], k["GB/s"] or 0, k["frac"] or 0, k["ms"]))
if d.get("bond_update"):
    b = d["bond_update"]
    print("  bond_update: matvec %.3f level1 %.3f trunc %.3f env %.3f total %.3f m=%s %s" % (b["matvec_s"], b["davidson_level1_s"], b["truncation_s"], b["env_update_s"], b["total_s"], b["kept_states"], b.get("truncation_phases")))
if d.get("sweep"):
    s = d["sweep"]
    print("  sweep: wall %.3f matvec %.3f trunc %.3f env %.3f cpu_sweep_s=%s" % (s["wall_s"], s["matvec_s"], s["truncation_s"], s["env_update_s"], s.get("cpu_sweep_s")))
if d.get("cpu_baseline"):
    print("  cpu:", {k: v for k, v in d["cpu_baseline"].items() if k not in ("sweep_how",)})
print("  clocks:", d.get("clocks"))
