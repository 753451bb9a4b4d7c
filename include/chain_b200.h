/* chain_b200 -- C ABI of the B200-native DMRG hot path for tzu-chen/Chain.
 *
 * The reference has no FFI of its own: its seam is the C++ call
 *     std::tie(en_, psi_) = dmrg(H, dmrg_progress_.DoneStates(), psi, sweeps, {"Quiet",true,"Weight=",1000.0});
 * at /root/reference/src/dmrg.h:544, :563 and :596 (ITensor's itensor::dmrg).  cb200_dmrg() below is what an adapter
 * for that call binds (INTEGRATION.md shows the adapter); the finer-grained entry points expose the individual
 * steps of that call (LocalMPO::product, davidson, MPS::svdBond, LocalMPO::position) for tests and profiling.
 *
 * Conventions
 *   - all arrays are column-major (first index fastest), exactly like ITensor's dense storage;
 *   - dtype 0 = float64, 1 = complex128 (interleaved re,im);
 *   - MPS site tensor A[l, s, r]; MPO site tensor W[a, s_out, s_in, a']; (s_in is the ket index ITensor calls s,
 *     s_out the primed index s');
 *   - Z_N symmetry is described by one charge label per index value (cb200_index), the block structure is
 *     derived from the labels: A[l,s,r] != 0 only if q(l)+q(s) = q(r), W != 0 only if q(a)+q(s_out) = q(s_in)+q(a')
 *     ... all mod N.  N = 1 (all labels 0) is the dense case (golden, haagerup); N = 3 is haagerupq;
 *   - every function returns 0 on success, a negative cb200_status otherwise; cb200_last_error() gives the text.
 *     No exception crosses this boundary and there is no CPU fallback: without a usable sm_100 device
 *     cb200_create() fails.
 */
#ifndef CHAIN_B200_H
#define CHAIN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb200_ctx cb200_ctx;

/* Argument limits (anything outside is refused with CB200_ERR_INVALID before any work is done, as are NULL arrays, ranks / dtypes / dimensions
 * that contradict each other, charge labels outside [0, nq), maxdim < 1, a NaN or negative cutoff ...: tools/fuzz_abi.py). */
#define CB200_MAX_NQ 64            /* order N of the Z_N symmetry (Chain uses 1 and 3) */
#define CB200_MAX_SITES 65536      /* chain length */
#define CB200_MAX_SITE_DIM 1024    /* physical dimension d (Chain: 2 and 6) */

enum cb200_status {
  CB200_OK = 0,
  CB200_ERR_NO_DEVICE = -1,
  CB200_ERR_INVALID = -2,
  CB200_ERR_CUDA = -3,
  CB200_ERR_NOMEM = -4,
  CB200_ERR_NOCONV = -5,
  CB200_ERR_RANGE = -6   /* the adapter rethrows this one as std::out_of_range, which DMRG::Simulate catches (src/dmrg.h:454) */
};

/* ---- context ------------------------------------------------------------------------------------------------ */
/* One device per process: every context of a process must name the same device (kernel attributes and the current-device setting are
 * process-wide); a second context on a different device fails with CB200_ERR_INVALID / CB200_ERR_NO_DEVICE.  Multi-GPU = one process per GPU. */
int cb200_create(cb200_ctx** out, int device);
void cb200_destroy(cb200_ctx* ctx);
const char* cb200_last_error(const cb200_ctx* ctx);   /* ctx may be NULL: error of the last failed cb200_create */
const char* cb200_backend_name(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t cb200_launch_count(const cb200_ctx* ctx);

/* ---- multi-GPU: the H_eff matvec row-sharded over the GPUs of one box (one process and one context per GPU) ------
 * Rank g applies H_eff for its own slice of the bra rows l' of every charge sector (1/world of the work) and the ranks
 * exchange output rows; everything else of a bond update is replicated and stays bit-identical across ranks.
 *   mode 0: no exchange -- rows a rank does not own come back as zeros, the caller sums the vectors (CPU/gloo tests);
 *   mode 1: ncclAllReduce(sum, ncclDouble) of the zero-padded partial vectors (libnccl.so.2 is dlopen'ed on first use);
 *   mode 2: all-gather fused into the R-stage GEMM epilogue: tiles are stored into every rank's landing buffer through
 *           CUDA-IPC peer mappings over NVLink, then a stream-ordered flag barrier.
 * Only bonds whose left link has dimension >= min_dim are sharded (default 1600, BASELINE north_star); pass -1 to keep it.
 * Setup: every rank calls cb200_shard_alloc (mode 2; landing_bytes >= the largest phi in bytes) and the host side
 * all-gathers the 64-byte handles; rank 0 calls cb200_nccl_unique_id (mode 1) and broadcasts the 128 bytes; then every
 * rank calls cb200_shard_setup. */
int cb200_nccl_unique_id(cb200_ctx* ctx, void* out128);
int cb200_shard_alloc(cb200_ctx* ctx, int64_t landing_bytes, void* handle_out64);
int cb200_shard_setup(cb200_ctx* ctx, int32_t rank, int32_t world, int32_t mode, const void* nccl_uid128 /* mode 1 */,
                      const void* handles /* mode 2: world x 64 bytes */, int64_t min_dim);

/* ---- problem description -------------------------------------------------------------------------------------- */
typedef struct cb200_index {
  int64_t dim;
  const int32_t* charge;      /* dim labels in [0, nq); NULL means all zero */
} cb200_index;

typedef struct cb200_tensor {
  int32_t dtype;              /* 0 f64, 1 c128 */
  int32_t rank;               /* 3 (MPS) or 4 (MPO) */
  int64_t dims[4];
  const void* data;           /* column-major, dense with exact zeros in forbidden blocks (elements of an MPS tensor that sit in a
                               * charge-forbidden block are ignored, not diagnosed: ITensor's QN storage cannot hold them; an MPO tensor
                               * with such elements is refused) */
} cb200_tensor;

/* per-sweep parameters: what itensor::Sweeps holds (src/dmrg.h:538-543, :589-593) */
typedef struct cb200_sweep {
  int32_t maxdim, mindim, niter;
  int32_t pad_;
  double cutoff, noise;
} cb200_sweep;

typedef struct cb200_bond_stat {
  int32_t sweep, bond, dir, m;       /* dir 0 = Fromleft, 1 = Fromright */
  double energy, truncerr;
} cb200_bond_stat;

/* An MPS handed across the boundary: n site tensors + n+1 link indices (link[0] and link[n] have dim 1). */
typedef struct cb200_mps {
  int32_t n, nq;
  const cb200_tensor* site;          /* n tensors [Dl, d, Dr] */
  const cb200_index* link;           /* n+1 */
  int32_t ortho_center;              /* 1-based; 0 = unknown (the library orthogonalises to site 1) */
  int32_t layout;                    /* 0: site[j].data is the dense [Dl,d,Dr] array; 1: it is BLOCK-PACKED, the way itensor::QDense<T> holds a
                                      * QN tensor: every index cut into sectors (maximal runs of equal charge label), the zero-flux sector
                                      * triples q(l)+q(s) = q(r) in increasing block number i_l + n_l*(i_s + n_s*i_r), each block dense
                                      * column-major [len_l, len_s, len_r], no gaps (cb200_site_blocks_elems elements).  dims stay (Dl, d, Dr). */
} cb200_mps;

typedef struct cb200_mpo {
  int32_t n, nq;
  const cb200_tensor* site;          /* n tensors [kl, d, d, kr] */
  const cb200_index* link;           /* n+1 MPO link indices */
  const int32_t* site_charge;        /* d labels of the physical index (same on every site) */
  int32_t d, pad_;
} cb200_mpo;

/* ---- coarse entry point: replaces itensor::dmrg(H, psis, psi0, sweeps, {"Weight",w}) -------------------------- */
typedef struct cb200_result cb200_result;   /* owns the optimised MPS until cb200_result_free */

int cb200_dmrg(cb200_ctx* ctx, const cb200_mpo* H, const cb200_mps* prev, int32_t nprev, const cb200_mps* psi0,
               const cb200_sweep* sweeps, int32_t nsweep, double weight, double* energy, cb200_result** out);
/* CalcEE of /root/reference/src/utility.cpp:5-27: von Neumann entropy of every cut b = 1..n-1 (p > 1e-12 rule, natural log),
 * Schmidt spectra from the device truncation path; out has n-1 entries */
int cb200_mps_entropies(cb200_ctx* ctx, const cb200_mps* psi, int32_t d, const int32_t* site_charge, double* out);
/* Translation by one site as the reference measures it (Measure("translation"), /root/reference/src/dmrg.h:855-858): L-1 swap
 * gates (Swap, src/chain.cpp:486-507), each followed by an SVD with relative cutoff `cutoff` (kCutoff = 1e-5 there); the
 * translated MPS comes back in a result object: read it with cb200_result_site, cb200_result_link_dim, cb200_result_link_charges */
int cb200_mps_translate(cb200_ctx* ctx, const cb200_mps* psi, int32_t d, const int32_t* site_charge, double cutoff, cb200_result** out);
/* innerC(a, b) = <a|b> (src/dmrg.h:1063); re_im[2] */
int cb200_mps_overlap(cb200_ctx* ctx, const cb200_mps* a, const cb200_mps* b, int32_t d, const int32_t* site_charge, double* re_im);
/* <bra| O |ket> for an MPO and two different dense (nq = 1) MPS: the matrix elements of the rho-defect operator that
 * Measure("rho") needs (/root/reference/src/dmrg.h:871-1047, :1068-1072; the reference removes the QNs first, :781-784) */
int cb200_mps_mpo_element(cb200_ctx* ctx, const cb200_mps* bra, const cb200_mpo* O, const cb200_mps* ket, double* re_im);
int32_t cb200_result_nsites(const cb200_result* r);
int32_t cb200_result_dtype(const cb200_result* r);                                  /* 0 f64, 1 c128 */
int64_t cb200_result_link_dim(const cb200_result* r, int32_t link);                 /* link in [0, n] */
int cb200_result_link_charges(const cb200_result* r, int32_t link, int32_t* out);  /* dim labels */
/* dense [Dl,d,Dr], caller-sized.  The optimised MPS stays on the device until it is asked for: this call packs one site tensor to
 * the dense external form on the device and copies it straight into `out` (every element is written, zeros included), so a result
 * holds device memory and must be freed before its context is destroyed */
int cb200_result_site(const cb200_result* r, int32_t site, void* out);
/* block-packed counterparts (cb200_mps.layout = 1): element count of a site tensor for given link labels, and the result's tensors in
 * that layout (links of a result are charge-sorted: one sector per charge) */
int64_t cb200_site_blocks_elems(int32_t nq, int32_t d, const int32_t* site_charge, const cb200_index* l, const cb200_index* r);
int64_t cb200_result_site_blocks_elems(const cb200_result* r, int32_t site);
int cb200_result_site_blocks(const cb200_result* r, int32_t site, void* out);
int32_t cb200_result_nstats(const cb200_result* r);
int cb200_result_stats(const cb200_result* r, cb200_bond_stat* out);
/* seconds spent in: [0] env update, [1] H_eff matvec, [2] Davidson level-1, [3] truncation, [4] other; [5] matvec count, [6] matvec flop */
int cb200_result_timers(const cb200_result* r, double* out7);
void cb200_result_free(cb200_result* r);

/* ---- fine-grained entry points (one bond of the path; used by tests, bench.py and ncu) ------------------------- */
/* A "bond problem" keeps L, W_b, W_{b+1}, R (and penalty vectors) resident on the device. */
typedef struct cb200_bond cb200_bond;

int cb200_bond_create(cb200_ctx* ctx, int32_t dtype, int32_t nq, int32_t d, const int32_t* site_charge,
                      const cb200_index* l, const cb200_index* r,          /* MPS links left / right of the two sites */
                      const cb200_index* kl, const cb200_index* km, const cb200_index* kr,   /* MPO links */
                      const void* L,   /* [Dl, kl, Dl]  (bra, mpo, ket) */
                      const void* W1,  /* [kl, d, d, km] */
                      const void* W2,  /* [km, d, d, kr] */
                      const void* R,   /* [Dr, kr, Dr] */
                      cb200_bond** out);
void cb200_bond_free(cb200_bond* b);
/* penalty vectors o_j (phi-shaped, dense [Dl,d,d,Dr]) with weight: H_eff += weight * sum_j |o_j><o_j| */
int cb200_bond_set_penalty(cb200_bond* b, int32_t nprev, const void* const* o, double weight);
int64_t cb200_bond_phi_elems(const cb200_bond* b);   /* Dl*d*d*Dr */
double cb200_bond_matvec_flops(const cb200_bond* b); /* F_alg of SURVEY 8(d): block-sparse, sparse-W */
/* LocalMPO(_MPS)::product: host phi in, host H_eff*phi out (copies included) */
int cb200_bond_apply(cb200_bond* b, const void* phi, void* out);
/* LocalMPO(_MPS)::product on ITensor's own QN storage: phi and the result cross the boundary BLOCK-PACKED ("QDense order"),
 * the way itensor::QDense<T> holds a QN tensor (a list of dense blocks over one contiguous buffer), instead of as a dense array
 * with exact zeros -- 1/nq of the bytes over PCIe and no scatter pass in the adapter.  Layout: every index is cut into sectors =
 * maximal runs of equal charge label (an ITensor QN index is such a list of (QN, dim) sectors); the stored blocks are the sector
 * tuples (i_l, i_s1, i_s2, i_r) with q(l)+q(s1)+q(s2) = q(r), in increasing block number i_l + n_l*(i_s1 + n_s*(i_s2 + n_s*i_r));
 * each block is dense column-major [len_l, len_s1, len_s2, len_r]; no gaps.  cb200_bond_phi_blocks_elems = total element count;
 * cb200_bond_phi_blocks lists the blocks (call with offsets = dims4 = NULL for the count).  For nq = 1 this is the dense array. */
int64_t cb200_bond_phi_blocks_elems(const cb200_bond* b);
int cb200_bond_phi_blocks(const cb200_bond* b, int32_t* nblocks, int64_t* offsets, int64_t* dims4 /* 4 per block */);
int cb200_bond_apply_blocks(cb200_bond* b, const void* phi_blocks, void* out_blocks);
/* The same product on a sharded context (cb200_shard_setup, world > 1) without redundant PCIe traffic: rank g passes only ITS contiguous
 * slice [begin, begin+count) of the block-packed vector (cb200_bond_phi_blocks_slice), the slices are all-gathered between the GPUs
 * (NVLink peer stores / ncclBroadcast), the row-sharded product runs, and rank g receives the same slice of the result -- the union of
 * the ranks' out_slice buffers is the whole vector.  With world = 1 it is cb200_bond_apply_blocks. */
int cb200_bond_phi_blocks_slice(const cb200_bond* b, int64_t* begin, int64_t* count);
int cb200_bond_apply_blocks_sharded(cb200_bond* b, const void* phi_slice, void* out_slice);
/* same product, device-resident, repeated `reps` times; returns mean milliseconds (CUDA events) */
int cb200_bond_apply_timed(cb200_bond* b, const void* phi, int32_t reps, int32_t warmup, double* ms_mean);
/* per-stage device time of the product (ms, mean of reps) and algorithmic flops: [0] L-stage GEMM, [1] W (x) W mix, [2] R-stage GEMM */
int cb200_bond_stage_times(cb200_bond* b, const void* phi, int32_t reps, double* ms3, double* flops3);
/* algorithmic HBM bytes of the same three stages (operands read once + result written once): the HBM-roofline numerator of the mix */
int cb200_bond_stage_bytes(const cb200_bond* b, double* bytes3);
/* one complete bond update resident on the device (davidson -> svdBond -> makeL/makeR), phase wall times in seconds:
 * [0] H_eff matvecs, [1] Davidson level-1, [2] truncation, [3] environment update */
int cb200_bond_update_timed(cb200_bond* b, const void* phi, int32_t dir, int32_t maxdim, double cutoff, int32_t niter,
                            double* energy, int32_t* m, int32_t* jacobi_sweeps, double* seconds4);
/* itensor::davidson (MaxIter = niter): phi in/out (host), eigenvalue out */
int cb200_bond_davidson(cb200_bond* b, void* phi, int32_t niter, double* eigenvalue, int32_t* nmatvec);
/* MPS::svdBond: phi (host) -> A [Dl,d,m], B [m,d,Dr]; call with A=B=NULL first to get m and charges_mid
 * (caller-sized to min(Dl*d, d*Dr)), then again with buffers.  dir 0 Fromleft, 1 Fromright.  Branch selection is ITensor's
 * (MPS::svdBond): cutoff >= 1e-12 takes the denmatDecomp path (rho = phi phi^H, Hermitian eigensolver, truncate on the eigenvalues);
 * cutoff < 1e-12 (noise is always 0 here) takes the true-SVD path, which resolves singular values down to eps*sigma_max instead of
 * sqrt(eps)*sigma_max (one-sided Jacobi on the QR-reduced block; truncate applied to sigma^2). */
int cb200_bond_truncate(cb200_bond* b, const void* phi, int32_t dir, int32_t maxdim, int32_t mindim, double cutoff,
                        int32_t* m, int32_t* charges_mid, double* truncerr, double* spectrum, void* A, void* B);
/* LocalMPO::makeL / makeR: new environment from the bond's L (dir 0, uses A [Dl,d,m] and W1) or R (dir 1, B [m,d,Dr], W2);
 * out is [m, km, m] */
int cb200_bond_env_update(cb200_bond* b, int32_t dir, int32_t m, const int32_t* charges_mid, const void* AorB, void* out);

/* ---- kernel-level entry points (unit tests of the device kernels; all pointers are HOST pointers) ------------- */
/* C[M x N] (ldc) = (accumulate? C : 0) + sum_seg op(A_seg) op(B_seg); segment s has K[s] columns, A_seg at A + a_off[s]. */
int cb200_k_gemm(cb200_ctx* ctx, int32_t dtype, int32_t trans_a, int32_t trans_b, int32_t conj_a, int32_t accumulate,
                 int32_t M, int32_t N, int32_t nseg, const int32_t* K, const void* A, int64_t a_elems, const int64_t* a_off,
                 int64_t lda, const void* B, int64_t b_elems, const int64_t* b_off, int64_t ldb, void* C, int64_t c_elems,
                 int64_t c_off, int64_t ldc, int32_t reps, double* ms_mean);
/* Hermitian eigensolver used inside the truncation: G [npairs][64 x 64] -> Q; returns rotation count */
int cb200_k_jacobi_eig(cb200_ctx* ctx, int32_t dtype, int32_t npairs, const void* G, void* Q, double tol, int64_t* rotations);
/* full one-sided block-Jacobi SVD of X [rows x cols]: X <- X V (columns orthogonal), V [cols x cols], w[cols] = column norms^2 */
int cb200_k_svd(cb200_ctx* ctx, int32_t dtype, int64_t rows, int64_t cols, void* X, void* V, double* w, int32_t* sweeps);
/* Hermitian eigensolver of the truncation (Householder tridiagonalisation, Sturm bisection, inverse iteration, Newton-Schulz
 * orthonormalisation, back-transformation): A [n x n] Hermitian (both triangles) -> w[n] all eigenvalues, descending, and
 * V [n x m] an orthonormal basis of the invariant subspace of the m largest.  Replaces LAPACK dsyev/zheev inside
 * ITensor's diagHermitian(rho). */
int cb200_k_heig(cb200_ctx* ctx, int32_t dtype, int64_t n, const void* A, int64_t m, double* w, void* V, int32_t* ns_iters);
/* the same for nprob independent matrices at once -- the charge blocks of one bond: their tridiagonalisations advance side by side
 * in one cooperative launch per panel step, each matrix on its own slice of the grid (the path cb200_bond_truncate takes for nq > 1) */
int cb200_k_heig_batched(cb200_ctx* ctx, int32_t dtype, int32_t nprob, const int64_t* n, const void* const* A, const int64_t* m,
                         double* const* w, void* const* V);
/* out[j] = <x_j | y>, dst = beta*dst + sum c_j x_j */
int cb200_k_dots(cb200_ctx* ctx, int32_t dtype, int32_t nv, int64_t n, const void* x, const void* y, double* out_re_im);
int cb200_k_lincomb(cb200_ctx* ctx, int32_t dtype, int32_t nv, int64_t n, const void* x, const double* c_re_im, double beta_re,
                    double beta_im, void* dst);
/* device-resident timing of the Davidson level-1 kernels on n-element vectors (HBM roofline entries of bench.py): mean milliseconds of
 * [0] multi_dot with nv vectors ((nv+1) n elements read), [1] lincomb dst = dst + sum_j c_j x_j ((nv+1) n read, n written), [2] device copy */
int cb200_k_level1_timed(cb200_ctx* ctx, int32_t dtype, int32_t nv, int64_t n, int32_t reps, double* ms3);
/* FP64 calibration: achieved TFLOP/s of the DMMA GEMM kernel on a square problem (device-resident, CUDA events) */
int cb200_k_gemm_peak(cb200_ctx* ctx, int32_t dtype, int32_t n, int32_t reps, double* tflops);
/* FP64 pipe issue-rate loops without memory traffic: which 0 = DMMA.8x8x4 (mma.sync.m8n8k4.f64), 1 = DFMA */
int cb200_k_fp64_peak(cb200_ctx* ctx, int32_t which, int32_t iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CHAIN_B200_H */
