"""FP64 calibration on the GPU box (MEASURED_PEAKS.json has no FP64 entry): DMMA / DFMA issue-rate loops, the DMMA
GEMM kernel on square problems, and cuBLAS DGEMM (via torch) as an outside reference.  Writes gpurun_out/fp64_peak.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import chain_b200 as cb  # noqa: E402

out = {"gpu": torch.cuda.get_device_name(0)}
ctx = cb.Context(0)
out["dmma_pipe_tflops"] = max(cb.k_fp64_peak(0, 40000, ctx=ctx) for _ in range(3))
out["dfma_pipe_tflops"] = max(cb.k_fp64_peak(1, 40000, ctx=ctx) for _ in range(3))
for n in (2048, 4096, 8192):
    out["gemm_f64_n%d_tflops" % n] = cb.k_gemm_peak(n, 5, False, ctx=ctx)
for n in (2048, 4096):
    out["gemm_c128_n%d_tflops" % n] = cb.k_gemm_peak(n, 5, True, ctx=ctx)
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.matmul(a, b)
torch.cuda.synchronize()
best = 0.0
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
out["cublas_dgemm_8192_tflops"] = best
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
