"""Summarises `ncu --page raw --csv` exports of the DMMA GEMM launches into profiles/rNN_ncu_gemm_<workload>_summary.json
(the file bench.py reads `roofline.traffic` from).   python tools/ncu_summary.py OUT.json WORKLOAD raw1.csv [raw2.csv ...]"""
import csv
import json
import sys

WANT = {
    "gpu_time_ms": "gpu__time_duration.sum",
    "dram_read_GB": "dram__bytes_read.sum",
    "dram_write_GB": "dram__bytes_write.sum",
    "dmma_pipe_active_pct": "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_active_realtime_pct": "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "fp64_pipe_active_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
}


def num(x):
    try:
        v = float(x.replace(",", ""))
        return None if v != v else v
    except Exception:
        return None


out_path, workload = sys.argv[1], sys.argv[2]
launches = []
for path in sys.argv[3:]:
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        e = {"kernel": r[ci["Kernel Name"]].split("(")[0].replace("void ", ""), "source": path}
        for k, col in WANT.items():
            if col in ci:
                v = num(r[ci[col]])
                if v is not None and units[ci[col]] in ("Mbyte",) and k.endswith("_GB"):
                    v /= 1e3
                if v is not None and units[ci[col]] in ("Kbyte",) and k.endswith("_GB"):
                    v /= 1e6
                if v is not None and units[ci[col]] in ("us",) and k == "gpu_time_ms":
                    v /= 1e3
                e[k] = v
        launches.append(e)
tot, covered = 0.0, []
for e in launches:
    if e.get("dram_read_GB") is not None and e.get("dram_write_GB") is not None:
        tot += (e["dram_read_GB"] + e["dram_write_GB"]) * 1e9
        covered.append(e["kernel"])
js = {"workload": workload, "launches": launches, "dram_bytes_per_matvec_gemm_launches": tot if covered else None,
      "dram_bytes_covers": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, of: " + "; ".join(covered)}
json.dump(js, open(out_path, "w"), indent=1)
print(json.dumps(js, indent=1)[:3000])
