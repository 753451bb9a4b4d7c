// Device backend interface used by the host-side DMRG engine (engine.cpp).
//
// The product library implements it in backend_cuda.cu (sm_100a kernels, one CUDA stream per context).
// tests/hostemu/backend_emu.cpp implements the same interface with serial CPU loops so that the engine's
// host logic (block bookkeeping, task-list generation, schedules) can be unit-tested in the GPU-less
// container; that file is never linked into libchain_b200.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include "tasks.h"

namespace cb200 {

constexpr int MAX_LC = 8;       // max vectors in one multi-dot / linear combination
constexpr int JB = 32;          // Jacobi block width (columns); inner eigenproblems are 2*JB x 2*JB

struct Backend;                 // opaque per-context state (stream, scratch)

Backend* backend_create(int device, char* err, size_t errlen);
void backend_destroy(Backend*);
const char* backend_name();
int backend_device(Backend*);

void* dev_alloc(Backend*, size_t bytes);              // returns nullptr on failure
void dev_free(Backend*, void* p);
void dev_memset0(Backend*, void* p, size_t bytes);
void dev_h2d(Backend*, void* dst, const void* src, size_t bytes);
void dev_d2h(Backend*, void* dst, const void* src, size_t bytes);   // synchronises the stream
void dev_d2d(Backend*, void* dst, const void* src, size_t bytes);
void dev_sync(Backend*);
const char* dev_last_error(Backend*);                 // nullptr if no error since last call
uint64_t dev_launch_count(Backend*);                  // kernels launched so far (for bench's gpu_launches)
void dev_event_record(Backend*, int slot);            // timing helpers (CUDA events on the context stream)
float dev_event_elapsed_ms(Backend*, int slot0, int slot1);

// Grouped segment-accumulating GEMM (tasks.h). All tasks of a launch share the transpose flags.
// Tile mode (tasks must be tiled with gemm_tile_n(cplx, mode)): WIDE = 128 columns; NARROW = 64 columns (Jacobi rotations, Gram
// blocks); NARROW_3M / MID_3M = 64 / 96 columns with the 3-multiplication complex product (complex only: 3 DMMA per complex MAC
// instead of 4; the third accumulator set does not fit the register budget at 128 columns).
enum { GEMM_TILE_WIDE = 0, GEMM_TILE_NARROW = 1, GEMM_TILE_NARROW_3M = 2, GEMM_TILE_MID_3M = 3 };
void launch_gemm(Backend*, bool cplx, bool transA, bool transB, int mode, const void* A, const void* B, void* C,
                 const GemmTask* d_tasks, int ntasks, const GemmSeg* d_segs, int total_tiles);
int gemm_tile_m(bool cplx);
int gemm_tile_n(bool cplx, int mode);

// ---- multi-GPU plumbing (one process per GPU) ---------------------------------------------------------
// Symmetric "landing" area: [flags | landing 0 | landing 1], allocated with cudaMalloc so that it can be exported through
// CUDA IPC.  shard_alloc fills a 64-byte handle; shard_connect opens the world's handles (own slot included, not opened).
constexpr int SHARD_HANDLE_BYTES = 64;
int shard_alloc(Backend*, int64_t landing_bytes_each, void* handle_out);
int shard_connect(Backend*, int rank, int world, const void* all_handles);
int64_t shard_landing_bytes(Backend*);                   // capacity of one landing half (0 when not allocated)
void* shard_landing(Backend*, int which);                // this rank's landing half
void shard_peer_bases(Backend*, int which, PeerBases* out);   // every rank's landing half `which`, own included
// stream-ordered barrier across the ranks: publishes this rank's peer stores and waits for everybody else's
void shard_barrier(Backend*);
// the R-stage GEMM (A not transposed, B transposed) with its C tiles stored to every destination in `peers`
void launch_gemm_peer(Backend*, bool cplx, int mode, const void* A, const void* B, const PeerBases& peers, const GemmTask* d_tasks, int ntasks,
                      const GemmSeg* d_segs, int total_tiles);
// broadcast `bytes` of device memory from rank `root` to every rank (in place): peer stores into the landing half `which` + the flag
// barrier (bytes must fit one landing half; the engine chunks), or ncclBroadcast
void shard_bcast_p2p(Backend*, void* buf, int64_t bytes, int root, int which);
void nccl_bcast(Backend*, void* buf, int64_t bytes, int root);
// all-gather in place: rank g holds the slice [g*chunk, min((g+1)*chunk, total)) of buf (chunk a multiple of 16 bytes) and ends up with all of it
void shard_allgather_p2p(Backend*, void* buf, int64_t total_bytes, int64_t chunk_bytes, int which);
// NCCL, loaded at run time (libnccl.so.2): unique id for rank 0, communicator, in-place sum all-reduce on the context stream
int nccl_unique_id(Backend*, void* out128);
int nccl_init(Backend*, const void* uid128, int rank, int world);
void nccl_allreduce_sum(Backend*, void* buf, int64_t n_doubles);

// Sparse MPO "mix" between the GEMM stages.
// max_in = largest in_count of any panel (decides whether the inputs of a row chunk fit in shared memory).
void launch_mix(Backend*, bool cplx, const void* in, void* out, const MixPanel* d_panels, int npanels, const MixIn* d_ins,
                const MixOut* d_outs, const MixEntry* d_entries, int total_blocks, int max_in);
int mix_rows_per_block(bool cplx, int max_in);

// Strided 3-D copies (permutes, gathers, conj-transposes).
void launch_copy(Backend*, bool cplx, const void* src, void* dst, const CopyTask* d_tasks, int ntasks, int total_blocks);
int copy_elems_per_block();

// Level-1: out[j] = sum_i conj(x_j[i]) * y[i], j < nv.  Result (nv complex pairs re,im; im=0 for real) is
// returned to the host (deterministic two-phase reduction, synchronises).
void multi_dot(Backend*, bool cplx, int nv, const void* const* x, const void* y, int64_t n, double* out_re_im);
// dst = beta*dst + sum_j c_j x_j   (c, beta complex given as re,im; imaginary parts ignored for real data)
void lincomb(Backend*, bool cplx, int nv, const void* const* x, const double* c_re_im, double beta_re, double beta_im,
             void* dst, int64_t n);

// Jacobi building blocks. G: npairs Hermitian (2JB x 2JB, column-major, ld 2JB) Gram matrices; on exit Q holds
// the unitary eigenvector matrices.  *d_rot (device int64) is incremented by the number of rotations applied.
// Columns whose squared norm (diagonal of G) is <= floor2 are treated as numerically zero and never rotated.
// Off-diagonal entries with |g|^2 <= abs_g2 are left alone (absolute criterion; 0 = purely relative).
// max_sweeps caps the cyclic sweeps of the inner solver (<= 0: run to convergence); the block SVD uses a small cap because
// its outer sweeps revisit every pair anyway.
void launch_jacobi_eig(Backend*, bool cplx, const void* G, void* Q, int npairs, double tol, double floor2, double abs_g2,
                       int max_sweeps, long long* d_rot);
// colnorm2[j] = sum_i |X[i + j*ld]|^2
void launch_colnorm2(Backend*, bool cplx, const void* X, int64_t ld, int64_t rows, int64_t cols, double* d_out);
// X (n x n, ld) = identity
void launch_set_identity(Backend*, bool cplx, void* X, int64_t ld, int64_t n);

// ---- Hermitian eigensolver of the bond truncation: rho -> tridiagonal -> eigenvalues -> kept eigenvectors ---------------
// (1) Householder tridiagonalisation of the Hermitian n x n matrix A (column-major, ld = n, BOTH triangles valid), in place:
//     d[n], e[n-1] (real) is the tridiagonal T = Q^H A Q; reflector j is H_j = I - tau[j] v v^H with v = [1; A[j+2.., j] * scl[j]]
//     acting on rows j+1.. (LAPACK zhetd2 'L' convention; the unscaled column stays in A).  tau/scl have A's element type.
//     p_work: n elements of A's type.  Returns false when the backend has no such kernel (the engine then uses block Jacobi).
bool launch_tridiag(Backend*, bool cplx, void* A, int64_t n, double* d, double* e, void* tau, void* scl, void* p_work);
//     Batched form for the charge blocks of one bond (independent matrices of different sizes): every panel step is ONE
//     cooperative launch in which each matrix owns a slice of the grid and its own barrier, so the per-step latency of the blocks
//     (two barriers + dependent L2 round trips, the bound below n ~ 4000) overlaps instead of adding up.  Same results, bit for bit,
//     as nprob calls of launch_tridiag as long as the per-matrix CTA counts agree -- they need not: every cross-CTA sum is
//     fixed-order for a GIVEN CTA count, and replicated ranks use the same count.
struct TriProblem {
  void* A;
  int64_t n;
  double* d;
  double* e;
  void* tau;
  void* scl;
  void* p_work;
};
bool launch_tridiag_batched(Backend*, bool cplx, const TriProblem* probs, int nprob);
// (2) all n eigenvalues of T by Sturm bisection, ascending
void launch_bisect(Backend*, const double* d, const double* e, int64_t n, double* w);
// (3) inverse iteration for m given eigenvalues (one thread per eigenvalue); Zt[i*m + j] = component i of eigenvector j, unit
//     2-norm; work: 6*n*m doubles
void launch_invit(Backend*, const double* d, const double* e, int64_t n, const double* lam, int64_t m, double* Zt, double* work);
// (4) Newton-Schulz orthonormalisation step: C = alpha*(1.5 I - 0.5 alpha^2 G) for the m x m Gram matrix G;
//     stats[0] = max |G - I|, stats[1] = max row sum of |G|  (device doubles, overwritten)
void launch_ns_coeff(Backend*, const double* G, int64_t m, double alpha, double* Cout, double* stats);
// (5) V[:, c] = H_0 H_1 ... H_{n-2} z_c for the m columns z_c = Zt[. * m + c]; V is n x m column-major of A's type
void launch_backtransform(Backend*, bool cplx, const void* A, int64_t n, const void* tau, const void* scl, const double* Zt, int64_t m, void* V);

// (5') blocked back-transformation on the GEMM kernel (compact WY, LAPACK dlarft/dlarfb 'Forward','Columnwise'):
//     zt_to_v: V[i + c*n] = Zt[i*m + c];  form_y: Y (len x nbk, ld = len, len = n-j0-1) = the reflectors j0..j0+nbk-1 as explicit
//     columns over rows j0+1..n-1;  wy_t: Tneg = -T with H_j0 ... H_{j0+nbk-1} = I - Y T Y^H, from S = Y^H Y (nbk x nbk) and tau.
void launch_zt_to_v(Backend*, bool cplx, const double* Zt, int64_t n, int64_t m, void* V);
void launch_form_y(Backend*, bool cplx, const void* A, int64_t n, const void* tau, const void* scl, int64_t j0, int nbk, void* Y);
void launch_wy_t(Backend*, bool cplx, const void* S, const void* tau_j0, int nbk, void* Tneg);
//     all blocks of one matrix in one launch: S / Tneg hold npanel slots of 64*64 elements (block k: nb x nb, ld nb, nb = 64 except
//     nb_last for the last block), tau is the whole tau array (block k starts at tau[64 k])
void launch_wy_t_batched(Backend*, bool cplx, const void* S, const void* tau, int npanel, int nb_last, void* Tneg);

// ---- two-stage reduction for large matrices: dense -> band (semi-bandwidth SB_B) with QR panels whose two-sided block updates run on
// the DMMA GEMM kernel, then band -> tridiagonal by bulge chasing in ONE persistent kernel; both sets of reflectors are kept for the
// back-transformation of the kept eigenvectors.  (The one-stage reduction above reads the whole trailing matrix once per reflector:
// n^3/3 elements of HBM traffic, memory-bound by construction.)
constexpr int SB_B = 64;            // semi-bandwidth of the intermediate band matrix = width of the QR panels
constexpr int SB_LD = 2 * SB_B;     // leading dimension of the band storage: AB[(i - j) + j*SB_LD], 0 <= i - j < 2*SB_B (room for the bulge)
constexpr int SB_VLD = 2 * SB_B;    // leading dimension of a block of bulge-chasing reflectors (window of SB_B + SB_B - 1 rows)
// (1) Householder QR of the panel P = A[r0 : n, j0 : j0 + SB_B], r0 = j0 + SB_B, len = n - r0 >= 2 (A column-major, ld = n), in place: the
//     panel is overwritten with R (rows <= columns) and zeros; Y (len x SB_B, ld = len) receives the kp = min(SB_B, len - 1) reflectors as
//     explicit columns (unit diagonal, zeros above it, columns >= kp zero); Tn = -T and Tnh = -T/2 (SB_B x SB_B, ld SB_B, upper
//     triangular, zero-padded) with H_0 ... H_{kp-1} = I - Y T Y^H.  A is then updated two-sidedly by the caller (GEMMs).
void launch_panel_qr(Backend*, bool cplx, void* A, int64_t n, int64_t j0, void* Y, void* Tn, void* Tnh);
// (2) AB = lower band of A (distances 0..SB_B), zero beyond
void launch_band_extract(Backend*, bool cplx, const void* A, int64_t n, void* AB);
// (3) band -> real symmetric tridiagonal (d[n], e[n-1]) by bulge chasing, for nprob independent matrices in one launch.  Sweep s
//     (column s) consists of tasks k = 0, 1, ... acting on rows c0 = s + 1 + k*SB_B .. c0 + 2*SB_B - 1; task (s, k) waits for task
//     (s-1, k+1) (progress counters prog[s], zero-initialised by the caller, n ints).  Reflector (s, k) (LAPACK larfg convention, explicit
//     leading 1) is stored for the back-transformation in block blk = g*K0 - g(g-1)/2 + k  (g = s / SB_B, K0 = ceil((n-1)/SB_B)) of V2
//     (SB_VLD x SB_B elements per block, zero-initialised by the caller), column s % SB_B, rows s % SB_B ...; tau2[blk*SB_B + s % SB_B].
struct ChaseProblem {
  void* AB;
  int64_t n;
  double* d;
  double* e;
  void* V2;
  void* tau2;
  int* prog;
};
int64_t chase_blocks(int64_t n);                       // number of reflector blocks of an n x n problem
void launch_bulge_chase(Backend*, bool cplx, const ChaseProblem* probs, int nprob);
// (4) Z <- Q2 Z for the n x m matrix Z (column-major, ld = n) with Q2 the product of the bulge-chasing reflectors (applied blockwise:
//     sweep groups descending, k ascending inside a group, sweeps descending inside a block -- the order that respects every
//     non-commuting pair)
void launch_q2_apply(Backend*, bool cplx, const void* V2, const void* tau2, int64_t n, void* Z, int64_t m);

// ITensor's truncate() evaluated on the device: pooled[n] (device, any order) is rank-sorted descending into sorted[n], then ONE
// thread applies the rule of SURVEY App. A.5 verbatim (negative tail zeroed, maxdim, relative cutoff on the discarded weight,
// degeneracy guard on docut).  out3 (device): [0] = m, [1] = truncerr, [2] = docut.
void launch_truncate_select(Backend*, const double* pooled, int64_t n, int64_t maxdim, int64_t mindim, double cutoff, double* sorted,
                            double* out3);

// FP64 issue-rate calibration (0 = DMMA.8x8x4, 1 = DFMA); TFLOP/s.  Device-only, used by bench.py for the roofline denominator.
double fp64_pipe_peak(Backend*, int which, int iters);

}  // namespace cb200
