"""Static evidence that needs no GPU: registers / spill bytes of every kernel from the ptxas log of the build (build/ptxas_backend.log, written by
`make`), and -- for the DMMA GEMM kernels -- where the spill instructions sit relative to the DMMA main loop in the SASS of the built library
(cuobjdump -sass): ptxas reports spills per kernel, but what costs time is a local-memory access between the first and the last DMMA of the k loop.

    python tools/ptxas_summary.py > profiles/r02_ptxas_registers_spills.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "build", "ptxas_backend.log")
LIB = os.path.join(ROOT, "chain_b200", "libchain_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", d).replace("void ", "") for d in out]


def main():
    t = open(LOG).read()
    ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Function properties for \1\n"
                      r"\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", t)
    names = demangle([e[0] for e in ents])
    print("# %d kernels in chain_b200/libchain_b200.so (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v)" % len(ents))
    print("# kernel | registers | stack B | spill stores B | spill loads B")
    for e, d in sorted(zip(ents, names), key=lambda x: x[1]):
        print("%-95s %4s %5s %5s %5s" % (d[:95], e[4], e[1], e[2], e[3]))
    nsp = sum(1 for e in ents if int(e[2]) or int(e[3]))
    print("\n# %d of %d kernels spill at all.  The GEMM kernels are compiled for the 168-register launch budget (384 threads) and re-balance with" % (nsp, len(ents)))
    print("# setmaxnreg (40 producer / 232 DMMA warps), so ptxas accounts their spills against 168.  Position of the local-memory instructions in the SASS:")
    print("# kernel | DMMA count | SASS lines of first / last DMMA | STL / LDL inside that range | STL / LDL outside")
    for e, d in sorted(zip(ents, names), key=lambda x: x[1]):
        if "gemm_grouped" not in d:
            continue
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", e[0], LIB], capture_output=True, text=True).stdout.splitlines()
        dm = [i for i, l in enumerate(sass) if "DMMA.8x8x4" in l]
        loc = [i for i, l in enumerate(sass) if re.search(r"\b(STL|LDL)\b|\bSTL\.|\bLDL\.", l)]
        if not dm:
            continue
        inside = [i for i in loc if dm[0] <= i <= dm[-1]]
        print("%-70s %4d  %6d / %6d  %3d  %3d" % (d[:70], len(dm), dm[0], dm[-1], len(inside), len(loc) - len(inside)))


if __name__ == "__main__":
    sys.exit(main())
