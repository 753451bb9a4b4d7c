"""ctypes binding of include/chain_b200.h.

The product library is chain_b200/libchain_b200.so (built in-tree by `make` / __graft_entry__.build()).
There is no fallback: if the library is missing, or no sm_100 device is usable, the error is raised to the caller.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libchain_b200.so")


class Index(C.Structure):
    _fields_ = [("dim", C.c_int64), ("charge", C.POINTER(C.c_int32))]


class Tensor(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("rank", C.c_int32), ("dims", C.c_int64 * 4), ("data", C.c_void_p)]


class Sweep(C.Structure):
    _fields_ = [("maxdim", C.c_int32), ("mindim", C.c_int32), ("niter", C.c_int32), ("pad_", C.c_int32),
                ("cutoff", C.c_double), ("noise", C.c_double)]


class BondStat(C.Structure):
    _fields_ = [("sweep", C.c_int32), ("bond", C.c_int32), ("dir", C.c_int32), ("m", C.c_int32),
                ("energy", C.c_double), ("truncerr", C.c_double)]


class MpsC(C.Structure):
    _fields_ = [("n", C.c_int32), ("nq", C.c_int32), ("site", C.POINTER(Tensor)), ("link", C.POINTER(Index)),
                ("ortho_center", C.c_int32), ("layout", C.c_int32)]


class MpoC(C.Structure):
    _fields_ = [("n", C.c_int32), ("nq", C.c_int32), ("site", C.POINTER(Tensor)), ("link", C.POINTER(Index)),
                ("site_charge", C.POINTER(C.c_int32)), ("d", C.c_int32), ("pad_", C.c_int32)]


# every symbol include/chain_b200.h declares: (name, restype, argtypes)
vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
SYMBOLS = [
    ("cb200_create", C.c_int, [C.POINTER(vp), C.c_int]),
    ("cb200_destroy", None, [vp]),
    ("cb200_last_error", C.c_char_p, [vp]),
    ("cb200_backend_name", C.c_char_p, []),
    ("cb200_launch_count", C.c_uint64, [vp]),
    ("cb200_nccl_unique_id", C.c_int, [vp, vp]),
    ("cb200_shard_alloc", C.c_int, [vp, i64, vp]),
    ("cb200_shard_setup", C.c_int, [vp, i32, i32, i32, vp, vp, i64]),
    ("cb200_dmrg", C.c_int, [vp, C.POINTER(MpoC), C.POINTER(MpsC), i32, C.POINTER(MpsC), C.POINTER(Sweep), i32, dbl,
                             C.POINTER(dbl), C.POINTER(vp)]),
    ("cb200_mps_entropies", C.c_int, [vp, C.POINTER(MpsC), i32, C.POINTER(i32), C.POINTER(dbl)]),
    ("cb200_mps_translate", C.c_int, [vp, C.POINTER(MpsC), i32, C.POINTER(i32), dbl, C.POINTER(vp)]),
    ("cb200_mps_overlap", C.c_int, [vp, C.POINTER(MpsC), C.POINTER(MpsC), i32, C.POINTER(i32), C.POINTER(dbl)]),
    ("cb200_mps_mpo_element", C.c_int, [vp, C.POINTER(MpsC), C.POINTER(MpoC), C.POINTER(MpsC), C.POINTER(dbl)]),
    ("cb200_result_nsites", i32, [vp]),
    ("cb200_result_dtype", i32, [vp]),
    ("cb200_result_link_dim", i64, [vp, i32]),
    ("cb200_result_link_charges", C.c_int, [vp, i32, C.POINTER(i32)]),
    ("cb200_result_site", C.c_int, [vp, i32, vp]),
    ("cb200_site_blocks_elems", i64, [i32, i32, C.POINTER(i32), C.POINTER(Index), C.POINTER(Index)]),
    ("cb200_result_site_blocks_elems", i64, [vp, i32]),
    ("cb200_result_site_blocks", C.c_int, [vp, i32, vp]),
    ("cb200_result_nstats", i32, [vp]),
    ("cb200_result_stats", C.c_int, [vp, C.POINTER(BondStat)]),
    ("cb200_result_timers", C.c_int, [vp, C.POINTER(dbl)]),
    ("cb200_result_free", None, [vp]),
    ("cb200_bond_create", C.c_int, [vp, i32, i32, i32, C.POINTER(i32), C.POINTER(Index), C.POINTER(Index), C.POINTER(Index),
                                    C.POINTER(Index), C.POINTER(Index), vp, vp, vp, vp, C.POINTER(vp)]),
    ("cb200_bond_free", None, [vp]),
    ("cb200_bond_set_penalty", C.c_int, [vp, i32, C.POINTER(vp), dbl]),
    ("cb200_bond_phi_elems", i64, [vp]),
    ("cb200_bond_matvec_flops", dbl, [vp]),
    ("cb200_bond_apply", C.c_int, [vp, vp, vp]),
    ("cb200_bond_phi_blocks_elems", i64, [vp]),
    ("cb200_bond_phi_blocks", C.c_int, [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]),
    ("cb200_bond_apply_blocks", C.c_int, [vp, vp, vp]),
    ("cb200_bond_stage_bytes", C.c_int, [vp, C.POINTER(dbl)]),
    ("cb200_bond_phi_blocks_slice", C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
    ("cb200_bond_apply_blocks_sharded", C.c_int, [vp, vp, vp]),
    ("cb200_bond_apply_timed", C.c_int, [vp, vp, i32, i32, C.POINTER(dbl)]),
    ("cb200_bond_stage_times", C.c_int, [vp, vp, i32, C.POINTER(dbl), C.POINTER(dbl)]),
    ("cb200_bond_update_timed", C.c_int, [vp, vp, i32, i32, dbl, i32, C.POINTER(dbl), C.POINTER(i32), C.POINTER(i32), C.POINTER(dbl)]),
    ("cb200_bond_davidson", C.c_int, [vp, vp, i32, C.POINTER(dbl), C.POINTER(i32)]),
    ("cb200_bond_truncate", C.c_int, [vp, vp, i32, i32, i32, dbl, C.POINTER(i32), C.POINTER(i32), C.POINTER(dbl),
                                      C.POINTER(dbl), vp, vp]),
    ("cb200_bond_env_update", C.c_int, [vp, i32, i32, C.POINTER(i32), vp, vp]),
    ("cb200_k_gemm", C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(i32), vp, i64, C.POINTER(i64), i64,
                               vp, i64, C.POINTER(i64), i64, vp, i64, i64, i64, i32, C.POINTER(dbl)]),
    ("cb200_k_jacobi_eig", C.c_int, [vp, i32, i32, vp, vp, dbl, C.POINTER(i64)]),
    ("cb200_k_svd", C.c_int, [vp, i32, i64, i64, vp, vp, C.POINTER(dbl), C.POINTER(i32)]),
    ("cb200_k_heig", C.c_int, [vp, i32, i64, vp, i64, C.POINTER(dbl), vp, C.POINTER(i32)]),
    ("cb200_k_heig_batched", C.c_int, [vp, i32, i32, C.POINTER(i64), C.POINTER(vp), C.POINTER(i64), C.POINTER(C.POINTER(dbl)), C.POINTER(vp)]),
    ("cb200_k_dots", C.c_int, [vp, i32, i32, i64, vp, vp, C.POINTER(dbl)]),
    ("cb200_k_lincomb", C.c_int, [vp, i32, i32, i64, vp, C.POINTER(dbl), dbl, dbl, vp]),
    ("cb200_k_level1_timed", C.c_int, [vp, i32, i32, i64, i32, C.POINTER(dbl)]),
    ("cb200_k_gemm_peak", C.c_int, [vp, i32, i32, i32, C.POINTER(dbl)]),
    ("cb200_k_fp64_peak", C.c_int, [vp, i32, i32, C.POINTER(dbl)]),
]


class ChainB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("chain_b200 error %d: %s" % (code, msg))
        self.code = code


_cache = {}


def load(path=None):
    """Load the C-ABI library and declare every prototype. Raises if the file or any symbol is missing."""
    path = os.path.abspath(path or DEFAULT_LIB)   # no environment override: the product library is the only default
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise ChainB200Error(-1, "%s not found: build it with `make` (or __graft_entry__.build()); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _cache[path] = lib
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def i32ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def fcol(a, dtype):
    """Column-major copy of an ndarray in the requested dtype (what crosses the boundary)."""
    return np.asfortranarray(np.asarray(a, dtype=dtype))
