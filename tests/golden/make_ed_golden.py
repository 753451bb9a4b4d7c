"""Regenerates tests/golden/ed_levels.json: exact-diagonalisation known answers for the restated Hamiltonians.

The reference ships no golden vectors; its only cross-check is mode 7 (src/dmrg.h:1283-1300),
an exact diagonalisation printed to stdout.  These values play that role.  They were cross-checked against the
values computed independently during the survey -- see tests/test_oracle.py.
Run from the repo root:  python tests/golden/make_ed_golden.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import ed  # noqa: E402

CASES = [
    ("golden", "p", 6, [-1.0], None, 5),
    ("golden", "p", 8, [-1.0], None, 5),
    ("golden", "p", 6, [1.0], None, 4),
    ("golden", "o", 6, [-1.0], None, 4),
    ("haagerup", "p", 6, [0.0, -1.0, 0.0], None, 6),
    ("haagerupq", "p", 6, [0.0, -1.0, 0.0], 0, 5),
    ("haagerupq", "p", 6, [0.0, -1.0, 0.0], 1, 5),
    ("haagerupq", "p", 6, [0.0, -1.0, 0.0], 2, 5),
    ("haagerupq", "p", 6, [-1.0, 0.0, 0.0], 0, 5),
    ("haagerupq", "p", 6, [-1.0, 0.0, 0.0], 1, 4),
    ("haagerupq", "p", 6, [0.0, 0.0, -1.0], 0, 4),
    ("haagerupq", "s", 6, [0.0, -1.0, 0.0], 0, 6),   # highly degenerate: only distinct values are meaningful
    ("haagerupq", "p", 8, [-1.0, 0.0, 0.0], 0, 3),   # BASELINE configs[4]'s physical chain (6^8/3 states per charge, ~40 s each)
    ("haagerupq", "p", 8, [-1.0, 0.0, 0.0], 1, 3),
]
EE_CASES = [("golden", "p", 6, [-1.0], None), ("haagerupq", "p", 6, [0.0, -1.0, 0.0], 0), ("haagerupq", "p", 6, [-1.0, 0.0, 0.0], 0)]

out = {"U": 2.0, "levels": [], "ee": []}
for st, bc, n, cp, q, k in CASES:
    w = ed.lowest_levels(st, bc, n, 2.0, cp, k=k, charge=q)
    out["levels"].append(dict(site=st, bc=bc, n=n, couplings=cp, charge=q, levels=[float(x) for x in w]))
    print(st, bc, n, cp, q, w)
for st, bc, n, cp, q in EE_CASES:
    e0, vec, site = ed.ground_state(st, bc, n, 2.0, cp, charge=q)
    ee = ed.entanglement_entropies(vec, site.d, n)
    out["ee"].append(dict(site=st, bc=bc, n=n, couplings=cp, charge=q, e0=float(e0), ee=ee))
    print(st, bc, n, cp, q, e0, ee)
with open(os.path.join(os.path.dirname(__file__), "ed_levels.json"), "w") as f:
    json.dump(out, f, indent=1)
