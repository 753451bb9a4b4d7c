"""CPU oracle for the Chain DMRG hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the algorithm the reference runs on its
hot path (ITensor C++ v3 `dmrg()` called at /root/reference/src/dmrg.h:544,563,596)
plus the model definitions that feed it (src/golden_site.cpp, src/haagerup_site.cpp,
src/haagerup_q_site.cpp, src/f_data.cpp).

PINNING.  MODELS (site operators, F-symbols, Hamiltonian term lists, state labels, Z3 sectors): pinned to outputs of the
REFERENCE'S OWN CODE -- src/{golden_site,haagerup_site,haagerup_q_site,f_data}.cpp compiled unmodified against a miniature
ITensor stand-in (oracle/Makefile -> oracle/_ref/ref_models; tests/golden/from_reference.py), committed as
tests/golden/reference_models.json and compared operator by operator / term by term in tests/test_reference_pin.py.
ENGINE (ITensor's dmrg/davidson/svdBond restated in oracle/dmrg.py from SURVEY App. A): PARITY UNPINNED by reference vectors --
the reference ships no golden vectors, tests or fixtures for this path, and ITensor (github.com/ITensor/ITensor, branch v3,
no pinned version) is absent from /root/reference and cannot be built here.  The engine restatement is pinned by
(i) exact diagonalisation of Hamiltonians assembled from the reference dump (tests/golden/ed_levels_reference.json ==
tests/golden/ed_levels.json), which the restated DMRG must reproduce, (ii) the rho / translation eigenvalue sets the reference
documents (src/dmrg.h:661-666,1100-1105), and (iii) the haagerup == union_q haagerupq cross-check.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; it is the checker, never the product path.
"""
