"""chain_b200: B200-native (sm_100a) implementation of the DMRG hot path of tzu-chen/Chain.

The product is the C-ABI shared library chain_b200/libchain_b200.so (include/chain_b200.h); this package is the
host-side mirror of the reference's `dmrg(H, psis, psi0, sweeps, args)` call plus thin wrappers for the
fine-grained entry points. Importing the package does not load the library; the first call does, and fails
loudly if the library or a usable device is missing.
"""
from .api import (Context, Sweeps, MPS, PackedMPS, pack_site_blocks, unpack_site_blocks, MPO, DmrgInfo, dmrg, calc_ee, translate, overlap, mpo_element, Bond, default_context,  # noqa: F401
                  k_gemm, k_jacobi_eig, k_svd, k_heig, k_heig_batched, k_dots, pack_phi_blocks, unpack_phi_blocks, k_lincomb, k_gemm_peak, k_fp64_peak, k_level1_timed)
from ._lib import ChainB200Error, DEFAULT_LIB  # noqa: F401
