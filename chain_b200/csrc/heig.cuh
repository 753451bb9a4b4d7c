// Hermitian eigensolver kernels of the bond truncation (MPS::svdBond -> denmatDecomp -> diagHermitian in ITensor, reached from
// /root/reference/src/dmrg.h:544,563,596):  rho --Householder--> T (real tridiagonal) --Sturm bisection--> eigenvalues
// --inverse iteration--> eigenvectors of T --(Newton-Schulz, GEMM)--> orthonormal --back-transformation--> eigenvectors of rho.
// Conventions are those of backend.h / LAPACK zhetd2('L').  Everything here is HBM/L2-bandwidth or latency bound; the O(n^3)
// GEMM parts (rho = X^H X, Newton-Schulz, X V) run on the DMMA kernel of gemm.cuh.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>

namespace cb200 {
namespace cg = cooperative_groups;

// ---- scalar helpers: T = double or double2 (complex) ----------------------------------------------------------------
__device__ __forceinline__ double t_zero(double) { return 0.0; }
__device__ __forceinline__ double2 t_zero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double t_one(double) { return 1.0; }
__device__ __forceinline__ double2 t_one(double2) { return make_double2(1.0, 0.0); }
__device__ __forceinline__ double t_conj(double a) { return a; }
__device__ __forceinline__ double2 t_conj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double t_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double2 t_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double t_add(double a, double b) { return a + b; }
__device__ __forceinline__ double2 t_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double t_sub(double a, double b) { return a - b; }
__device__ __forceinline__ double2 t_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double t_scale(double a, double s) { return a * s; }
__device__ __forceinline__ double2 t_scale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ double t_norm2(double a) { return a * a; }
__device__ __forceinline__ double t_norm2(double2 a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ double t_re(double a) { return a; }
__device__ __forceinline__ double t_re(double2 a) { return a.x; }
__device__ __forceinline__ double t_im(double) { return 0.0; }
__device__ __forceinline__ double t_im(double2 a) { return a.y; }
__device__ __forceinline__ bool t_is_zero(double a) { return a == 0.0; }
__device__ __forceinline__ bool t_is_zero(double2 a) { return a.x == 0.0 && a.y == 0.0; }
__device__ __forceinline__ double t_make(double r, double, double) { return r; }
__device__ __forceinline__ double2 t_make(double r, double i, double2) { return make_double2(r, i); }
__device__ __forceinline__ double t_inv(double a) { return 1.0 / a; }
__device__ __forceinline__ double2 t_inv(double2 a) { const double q = 1.0 / (a.x * a.x + a.y * a.y); return make_double2(a.x * q, -a.y * q); }
__device__ __forceinline__ double t_shfl_down(double v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double2 t_shfl_down(double2 v, int o) {
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, o), __shfl_down_sync(0xffffffffu, v.y, o));
}
__device__ __forceinline__ double t_shfl(double v, int l) { return __shfl_sync(0xffffffffu, v, l); }
__device__ __forceinline__ double2 t_shfl(double2 v, int l) { return make_double2(__shfl_sync(0xffffffffu, v.x, l), __shfl_sync(0xffffffffu, v.y, l)); }
template <class T>
__device__ __forceinline__ T t_warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = t_add(v, t_shfl_down(v, o));
  return t_shfl(v, 0);
}

constexpr int TRI_THREADS = 512;
constexpr int TRI_WARPS = TRI_THREADS / 32;

// block-wide sum, identical summation order in every CTA (all CTAs compute the Householder scalars redundantly and must agree bitwise)
template <class T>
__device__ __forceinline__ T tri_block_sum(T v, T* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = t_warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  T s = t_zero(v);
#pragma unroll
  for (int k = 0; k < TRI_WARPS; ++k) s = t_add(s, red[k]);
  return s;
}

// One persistent cooperative kernel for the whole reduction: two grid barriers per Householder step.
//   phase B  p = tau * A22 * v      one warp per column (A22 is Hermitian and stored in full, so row c = conj(column c): coalesced)
//   phase C  w = p - (tau/2)(p^H v) v ;  A22 -= v w^H + w v^H     the same warp updates the same column
template <bool CPLX>
__global__ void __launch_bounds__(TRI_THREADS, 2) tridiag_kernel(void* __restrict__ A_, long long n, double* __restrict__ d, double* __restrict__ e,
                                                              void* __restrict__ tau_, void* __restrict__ scl_, void* __restrict__ p_) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  cg::grid_group grid = cg::this_grid();
  T* A = reinterpret_cast<T*>(A_);
  T* tau = reinterpret_cast<T*>(tau_);
  T* scl = reinterpret_cast<T*>(scl_);
  T* p = reinterpret_cast<T*>(p_);
  __shared__ T red[TRI_WARPS];
  __shared__ double redd[TRI_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * TRI_WARPS + warp, GW = (long long)gridDim.x * TRI_WARPS;
  for (long long j = 0; j + 1 < n; ++j) {
    const long long m = n - j - 1;
    const T* x = A + (j + 1) + j * n;
    // ---- phase A: Householder scalars (zlarfg), redundantly per CTA
    double part = 0.0;
    for (long long i = 1 + threadIdx.x; i < m; i += TRI_THREADS) part += t_norm2(x[i]);
    const double sigma = tri_block_sum(part, redd);
    const T alpha = x[0];
    const double ar = t_re(alpha), ai = t_im(alpha);
    if (blockIdx.x == 0 && threadIdx.x == 0) d[j] = t_re(A[j + j * n]);
    if (sigma == 0.0 && ai == 0.0) {
      if (blockIdx.x == 0 && threadIdx.x == 0) { tau[j] = t_zero(alpha); scl[j] = t_zero(alpha); e[j] = ar; }
      continue;   // H = I; every CTA takes this branch together (same inputs, same arithmetic)
    }
    const double beta = -copysign(sqrt(ar * ar + ai * ai + sigma), ar);
    const T tj = t_make((beta - ar) / beta, -ai / beta, alpha);
    const T sc = t_inv(t_sub(alpha, t_make(beta, 0.0, alpha)));
    if (blockIdx.x == 0 && threadIdx.x == 0) { tau[j] = tj; scl[j] = sc; e[j] = beta; }
    T* A22 = A + (j + 1) + (j + 1) * n;
    // ---- phase B (4 independent loads per lane in flight: the column walk is latency-bound otherwise)
    for (long long c = gw; c < m; c += GW) {
      const T* col = A22 + c * n;
      T acc0 = t_zero(alpha), acc1 = acc0, acc2 = acc0, acc3 = acc0;
      long long r = lane;
      for (; r + 96 < m; r += 128) {
        const T a0 = col[r], a1 = col[r + 32], a2_ = col[r + 64], a3 = col[r + 96];
        const T x0 = x[r], x1 = x[r + 32], x2 = x[r + 64], x3 = x[r + 96];
        acc0 = t_add(acc0, t_mul(t_conj(a0), r == 0 ? t_one(alpha) : t_mul(x0, sc)));
        acc1 = t_add(acc1, t_mul(t_conj(a1), t_mul(x1, sc)));
        acc2 = t_add(acc2, t_mul(t_conj(a2_), t_mul(x2, sc)));
        acc3 = t_add(acc3, t_mul(t_conj(a3), t_mul(x3, sc)));
      }
      for (; r < m; r += 32) acc0 = t_add(acc0, t_mul(t_conj(col[r]), r == 0 ? t_one(alpha) : t_mul(x[r], sc)));
      const T acc = t_warp_sum(t_add(t_add(acc0, acc1), t_add(acc2, acc3)));
      if (lane == 0) p[c] = t_mul(tj, acc);
    }
    grid.sync();
    // ---- phase C
    T sp = t_zero(alpha);
    for (long long i = threadIdx.x; i < m; i += TRI_THREADS) {
      const T vi = i == 0 ? t_one(alpha) : t_mul(x[i], sc);
      sp = t_add(sp, t_mul(t_conj(p[i]), vi));
    }
    const T s = tri_block_sum(sp, red);
    const T a2 = t_scale(t_mul(tj, s), -0.5);
    for (long long c = gw; c < m; c += GW) {
      T* col = A22 + c * n;
      const T vc = c == 0 ? t_one(alpha) : t_mul(x[c], sc);
      const T wc = t_add(p[c], t_mul(a2, vc));
      const T cvc = t_conj(vc), cwc = t_conj(wc);
      long long r = lane;
      for (; r + 96 < m; r += 128) {
        T a[4], xv[4], pv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = col[r + 32 * u]; xv[u] = x[r + 32 * u]; pv[u] = p[r + 32 * u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const T vr = (r + 32 * u) == 0 ? t_one(alpha) : t_mul(xv[u], sc);
          const T wr = t_add(pv[u], t_mul(a2, vr));
          col[r + 32 * u] = t_sub(a[u], t_add(t_mul(vr, cwc), t_mul(wr, cvc)));
        }
      }
      for (; r < m; r += 32) {
        const T vr = r == 0 ? t_one(alpha) : t_mul(x[r], sc);
        const T wr = t_add(p[r], t_mul(a2, vr));
        col[r] = t_sub(col[r], t_add(t_mul(vr, cwc), t_mul(wr, cvc)));
      }
    }
    grid.sync();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) d[n - 1] = t_re(A[(n - 1) + (n - 1) * n]);
}

// Grid-wide barrier policies of the panel kernel: the whole cooperative grid (one matrix per launch), or a slice of it (batched
// launch: the charge blocks of one bond side by side, each on its own CTAs with its own arrival counter; all CTAs of the launch
// are co-resident because the launch is cooperative, the counter only grows -- barrier k completes at k * nblocks arrivals).
struct TriGridSync {
  cg::grid_group grid;
  __device__ __forceinline__ void sync() { grid.sync(); }
};
struct TriSliceSync {
  unsigned int* ctr;
  unsigned int nblocks, target;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += nblocks;
      __threadfence();                                   // publish this CTA's writes before arriving
      atomicAdd(ctr, 1u);
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
      } while (seen < target);
      __threadfence();
    }
    __syncthreads();
  }
};

// Blocked variant (LAPACK zlatrd 'L' restated for one persistent kernel per panel of <= 64 reflectors): the trailing matrix is
// NOT updated inside a panel -- the matrix-vector product is corrected with the panel's V and W columns instead -- so a step
// makes ONE pass over the trailing block (the unblocked kernel makes three) and the rank-2nb update after the panel is a GEMM
// on the DMMA kernel.  Still two grid barriers per step:
//   S2+S3  Householder scalars (redundant per CTA, fixed-order sums); one warp per column: p = A22^H-column dot v for the trailing
//          columns, g1 = W^H v, g2 = V^H v for the panel columns
//   S4+S1' one thread per row: w = tau (p - V g1 - W g2) + alpha v (column i of W), then the deferred update of the NEXT column
//          of A with all panel reflectors so far, and the partial norms the next step's Householder vector needs
template <bool CPLX, class SYNC>
__device__ __forceinline__ void
tridiag_panel_body(void* __restrict__ A_, long long n, long long j0, int nbk, double* __restrict__ d, double* __restrict__ e,
                   void* __restrict__ tau_, void* __restrict__ scl_, void* __restrict__ p_, void* __restrict__ g_,
                   double* __restrict__ partial, void* __restrict__ Vp_, void* __restrict__ W_, void* __restrict__ Vn_, void* __restrict__ Wn_,
                   void* __restrict__ Vt_, void* __restrict__ Wt_, int fsplit_long, const int bid, const int nblk, SYNC& gsync) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  T* A = reinterpret_cast<T*>(A_);
  T* tau = reinterpret_cast<T*>(tau_);
  T* scl = reinterpret_cast<T*>(scl_);
  T* p = reinterpret_cast<T*>(p_);
  T* g = reinterpret_cast<T*>(g_);
  T* Vt = reinterpret_cast<T*>(Vt_);   // transposed copies of the panel, [k + 64*r]
  T* Wt = reinterpret_cast<T*>(Wt_);
  T* Vp = reinterpret_cast<T*>(Vp_);
  T* W = reinterpret_cast<T*>(W_);
  T* Vn = reinterpret_cast<T*>(Vn_);
  T* Wn = reinterpret_cast<T*>(Wn_);
  __shared__ T red[TRI_WARPS];
  __shared__ double redd[TRI_WARPS];
  __shared__ T sg1[64], sg2[64], swp[64], svp[64];
  __shared__ T sa2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)bid * TRI_WARPS + warp, GW = (long long)nblk * TRI_WARPS;
  const long long gtid = (long long)bid * TRI_THREADS + threadIdx.x, GT = (long long)nblk * TRI_THREADS;
  const int G = nblk;
  const T zero = t_zero(T()), one = t_one(T());
  {   // prologue: the first column of the panel is up to date (rank-2nb GEMM of the previous panel); its partial norms are not
    double part = 0.0;
    for (long long r = j0 + 2 + gtid; r < n; r += GT) part += t_norm2(A[r + j0 * n]);
    const double sblk = tri_block_sum(part, redd);
    if (threadIdx.x == 0) partial[bid] = sblk;
  }
  gsync.sync();
  for (int i = 0; i < nbk; ++i) {
    const long long j = j0 + i, m = n - j - 1;
    const T* x = A + (j + 1) + j * n;
    // ---- S2
    double ps = 0.0;
    for (int b = threadIdx.x; b < G; b += TRI_THREADS) ps += partial[b];
    const double sigma = tri_block_sum(ps, redd);
    const T alpha = x[0];
    const double ar = t_re(alpha), ai = t_im(alpha);
    const bool off = (sigma == 0.0 && ai == 0.0);
    double beta = ar;
    T tj = zero, sc = zero;
    if (!off) {
      beta = -copysign(sqrt(ar * ar + ai * ai + sigma), ar);
      tj = t_make((beta - ar) / beta, -ai / beta, alpha);
      sc = t_inv(t_sub(alpha, t_make(beta, 0.0, alpha)));
    }
    if (bid == 0 && threadIdx.x == 0) { d[j] = t_re(A[j + j * n]); e[j] = beta; tau[j] = tj; scl[j] = sc; }
    // ---- S3: f warps of one CTA share a column when there are fewer columns than warps (more bytes in flight per column,
    // shorter dependent chain per warp); their partial sums are combined in a fixed order through shared memory
    if (!off) {
      const long long ncol = m + 2 * i;
      int f = 1;
      // (only for columns long enough to be bandwidth-bound: the two extra CTA barriers cost more than they save on short ones)
      if (m * (long long)sizeof(T) >= 16384) {
        while (f < TRI_WARPS && ncol * (2 * f) <= GW) f *= 2;
        if (f < fsplit_long) f = fsplit_long;     // long columns: few wide streams per CTA instead of one narrow stream per warp
      }
      if (f == 1) {   // the common (latency-bound) case: one warp per column, no CTA barriers
        for (long long cc = gw; cc < ncol; cc += GW) {
          const T* col = cc < m ? A + (j + 1) + (j + 1 + cc) * n : (cc < m + i ? W + (j + 1) + (cc - m) * n : Vp + (j + 1) + (cc - m - i) * n);
          T acc0 = zero, acc1 = zero, acc2 = zero, acc3 = zero;
          long long r = lane;
          for (; r + 96 < m; r += 128) {
            const T a0 = col[r], a1 = col[r + 32], a2_ = col[r + 64], a3 = col[r + 96];
            const T x0 = x[r], x1 = x[r + 32], x2 = x[r + 64], x3 = x[r + 96];
            acc0 = t_add(acc0, t_mul(t_conj(a0), r == 0 ? one : t_mul(x0, sc)));
            acc1 = t_add(acc1, t_mul(t_conj(a1), t_mul(x1, sc)));
            acc2 = t_add(acc2, t_mul(t_conj(a2_), t_mul(x2, sc)));
            acc3 = t_add(acc3, t_mul(t_conj(a3), t_mul(x3, sc)));
          }
          for (; r < m; r += 32) acc0 = t_add(acc0, t_mul(t_conj(col[r]), r == 0 ? one : t_mul(x[r], sc)));
          const T acc = t_warp_sum(t_add(t_add(acc0, acc1), t_add(acc2, acc3)));
          if (lane == 0) {
            if (cc < m) p[cc] = acc;
            else if (cc < m + i) g[cc - m] = acc;
            else g[64 + (cc - m - i)] = acc;
          }
        }
      } else {
      const long long cstride = GW / f;
      const int part = warp & (f - 1);
      const long long seg = ((m + f - 1) / f + 31) / 32 * 32;          // rows per part, multiple of 32
      const long long rbeg = (long long)part * seg, rend = rbeg + seg < m ? rbeg + seg : m;
      const long long trips = (ncol + cstride - 1) / cstride;
      for (long long tt = 0; tt < trips; ++tt) {
        const long long cc = gw / f + tt * cstride;
        T acc = zero;
        if (cc < ncol) {
          const T* col = cc < m ? A + (j + 1) + (j + 1 + cc) * n : (cc < m + i ? W + (j + 1) + (cc - m) * n : Vp + (j + 1) + (cc - m - i) * n);
          T acc0 = zero, acc1 = zero, acc2 = zero, acc3 = zero;
          long long r = rbeg + lane;
          for (; r + 96 < rend; r += 128) {
            const T a0 = col[r], a1 = col[r + 32], a2_ = col[r + 64], a3 = col[r + 96];
            const T x0 = x[r], x1 = x[r + 32], x2 = x[r + 64], x3 = x[r + 96];
            acc0 = t_add(acc0, t_mul(t_conj(a0), r == 0 ? one : t_mul(x0, sc)));
            acc1 = t_add(acc1, t_mul(t_conj(a1), t_mul(x1, sc)));
            acc2 = t_add(acc2, t_mul(t_conj(a2_), t_mul(x2, sc)));
            acc3 = t_add(acc3, t_mul(t_conj(a3), t_mul(x3, sc)));
          }
          for (; r < rend; r += 32) acc0 = t_add(acc0, t_mul(t_conj(col[r]), r == 0 ? one : t_mul(x[r], sc)));
          acc = t_warp_sum(t_add(t_add(acc0, acc1), t_add(acc2, acc3)));
        }
        if (f > 1) {
          __syncthreads();
          if (lane == 0) red[warp] = acc;
          __syncthreads();
          if (part == 0) {
            acc = red[warp];
            for (int q = 1; q < f; ++q) acc = t_add(acc, red[warp + q]);
          }
        }
        if (lane == 0 && part == 0 && cc < ncol) {
          if (cc < m) p[cc] = acc;
          else if (cc < m + i) g[cc - m] = acc;              // g1 = W^H v
          else g[64 + (cc - m - i)] = acc;                    // g2 = V^H v
        }
      }
      }
    }
    gsync.sync();
    // ---- S4 + S1'
    T sp = zero;
    if (!off)
      for (long long q = threadIdx.x; q < m; q += TRI_THREADS) sp = t_add(sp, t_mul(t_conj(p[q]), q == 0 ? one : t_mul(x[q], sc)));
    const T s0 = tri_block_sum(sp, red);
    if (threadIdx.x < i) {   // row j+1 of the panel (transposed copies: contiguous) and the g vectors
      sg1[threadIdx.x] = off ? zero : g[threadIdx.x];
      sg2[threadIdx.x] = off ? zero : g[64 + threadIdx.x];
      swp[threadIdx.x] = t_conj(Wt[threadIdx.x + (j + 1) * 64]);
      svp[threadIdx.x] = t_conj(Vt[threadIdx.x + (j + 1) * 64]);
    }
    __syncthreads();
    if (warp == 0) {   // alpha' and row j+1 of the new W column (every CTA needs both), lanes over the panel index
      T cross = zero, ptc = zero;
      for (int k = lane; k < i; k += 32) {
        cross = t_add(cross, t_add(t_mul(t_conj(sg1[k]), sg2[k]), t_mul(t_conj(sg2[k]), sg1[k])));
        ptc = t_add(ptc, t_add(t_mul(t_conj(svp[k]), sg1[k]), t_mul(t_conj(swp[k]), sg2[k])));
      }
      cross = t_warp_sum(cross);
      ptc = t_warp_sum(ptc);
      if (lane == 0) {
        const T a2w = t_scale(t_sub(s0, cross), -0.5 * t_norm2(tj));
        const T wf0 = off ? zero : t_add(t_mul(tj, t_sub(p[0], ptc)), a2w);
        sa2 = a2w;
        swp[i] = t_conj(wf0);
        svp[i] = one;
      }
    }
    __syncthreads();
    const T a2 = sa2;
    const bool next = (j + 1 <= n - 1) && (i + 1 < nbk || j + 1 == n - 1);
    double part = 0.0;
    // one warp per row: the lanes split the panel index (rows of the transposed panel copies are contiguous -> one round trip)
    for (long long r = (j + 1) + gw; r < n; r += GW) {
      const long long rr = r - (j + 1);
      T pt = zero, an = zero;
      for (int k = lane; k < i; k += 32) {
        const T vk = Vt[k + r * 64], wk = Wt[k + r * 64];
        pt = t_add(pt, t_add(t_mul(vk, sg1[k]), t_mul(wk, sg2[k])));
        an = t_add(an, t_add(t_mul(vk, swp[k]), t_mul(wk, svp[k])));
      }
      T p_r = zero, x_r = zero, a_r = zero;
      if (lane == 0) {
        if (!off) p_r = p[rr];
        if (rr != 0) x_r = x[rr];
        if (next) a_r = A[r + (j + 1) * n];
      }
      pt = t_warp_sum(pt);
      an = t_warp_sum(an);
      if (lane == 0) {
        const T vr = rr == 0 ? one : t_mul(x_r, sc);
        const T wf = off ? zero : t_add(t_mul(tj, t_sub(p_r, pt)), t_mul(a2, vr));
        W[r + (long long)i * n] = wf; Wn[r + (long long)i * n] = t_scale(wf, -1.0);
        Vp[r + (long long)i * n] = vr; Vn[r + (long long)i * n] = t_scale(vr, -1.0);
        Wt[i + r * 64] = wf; Vt[i + r * 64] = vr;
        if (next) {
          const T anew = t_sub(t_sub(a_r, an), t_add(t_mul(vr, swp[i]), t_mul(wf, svp[i])));
          A[r + (j + 1) * n] = anew;
          if (r >= j + 3) part += t_norm2(anew);
        }
      }
    }
    const double sblk = tri_block_sum(part, redd);
    if (threadIdx.x == 0) partial[bid] = sblk;
    gsync.sync();
  }
  if (j0 + nbk == n - 1 && bid == 0 && threadIdx.x == 0) d[n - 1] = t_re(A[(n - 1) + (n - 1) * n]);
}

template <bool CPLX>
__global__ void __launch_bounds__(TRI_THREADS, 2)
tridiag_panel_kernel(void* __restrict__ A_, long long n, long long j0, int nbk, double* __restrict__ d, double* __restrict__ e,
                     void* __restrict__ tau_, void* __restrict__ scl_, void* __restrict__ p_, void* __restrict__ g_,
                     double* __restrict__ partial, void* __restrict__ Vp_, void* __restrict__ W_, void* __restrict__ Vn_, void* __restrict__ Wn_,
                     void* __restrict__ Vt_, void* __restrict__ Wt_, int fsplit_long) {
  TriGridSync gsync{cg::this_grid()};
  tridiag_panel_body<CPLX>(A_, n, j0, nbk, d, e, tau_, scl_, p_, g_, partial, Vp_, W_, Vn_, Wn_, Vt_, Wt_, fsplit_long, (int)blockIdx.x,
                           (int)gridDim.x, gsync);
}

// Batched form: descs[k] is one matrix (one charge block of the bond) advancing through the SAME panel step on CTAs
// [blk_begin, blk_begin + blk_count) of a cooperative launch.  Below n ~ 4000 a Householder step is bound by its two grid barriers
// and its dependent L2 round trips, not by bandwidth: running the blocks side by side overlaps exactly that latency.
struct TriDesc {
  void* A;
  long long n, j0;
  double *d, *e;
  void *tau, *scl, *p, *g;
  double* partial;
  void *Vp, *W, *Vn, *Wn, *Vt, *Wt;
  unsigned int* ctr;
  int nbk, blk_begin, blk_count, fsplit;
};
template <bool CPLX>
__global__ void __launch_bounds__(TRI_THREADS, 2) tridiag_panel_batched_kernel(const TriDesc* __restrict__ descs, int ndesc) {
  int k = 0;
  while (k + 1 < ndesc && (int)blockIdx.x >= descs[k + 1].blk_begin) ++k;
  const TriDesc D = descs[k];
  TriSliceSync gsync{D.ctr, (unsigned int)D.blk_count, 0u};
  tridiag_panel_body<CPLX>(D.A, D.n, D.j0, D.nbk, D.d, D.e, D.tau, D.scl, D.p, D.g, D.partial, D.Vp, D.W, D.Vn, D.Wn, D.Vt, D.Wt, D.fsplit,
                           (int)blockIdx.x - D.blk_begin, D.blk_count, gsync);
}

// bounds[0..3] = Gershgorin lower / upper bound (widened), pivmin, ||T||_1
__global__ void __launch_bounds__(256) tri_bounds_kernel(const double* __restrict__ d, const double* __restrict__ e, long long n, double* __restrict__ bounds) {
  double gl = 1e300, gu = -1e300, em = 0.0, tn = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i + 1 < n ? fabs(e[i]) : 0.0);
    gl = fmin(gl, d[i] - r); gu = fmax(gu, d[i] + r);
    tn = fmax(tn, fabs(d[i]) + r);
    if (i + 1 < n) em = fmax(em, e[i] * e[i]);
  }
  __shared__ double s[4][256];
  s[0][threadIdx.x] = gl; s[1][threadIdx.x] = gu; s[2][threadIdx.x] = em; s[3][threadIdx.x] = tn;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 256; ++k) { gl = fmin(gl, s[0][k]); gu = fmax(gu, s[1][k]); em = fmax(em, s[2][k]); tn = fmax(tn, s[3][k]); }
    const double pivmin = 2.2250738585072014e-308 * fmax(1.0, em);
    const double span = fmax(gu - gl, 1e-300);
    const double pad = 2.0 * span * 2.3e-16 * (double)n + 2.0 * pivmin;
    bounds[0] = gl - pad; bounds[1] = gu + pad; bounds[2] = pivmin; bounds[3] = tn;
  }
}

// warp k -> k-th smallest eigenvalue of T by 32-way multisection on the Sturm count: every round the lanes evaluate the count
// at 32 interior points of [lo, hi] and the interval shrinks 33x, so ~11 sequential passes over (d, e) reach eps/4 * ||T||
// (plain bisection needs ~55; the pass is a division-latency-bound recurrence, so passes are what cost).
__global__ void __launch_bounds__(128) bisect_kernel(const double* __restrict__ d, const double* __restrict__ e, long long n,
                                                     const double* __restrict__ bounds, double* __restrict__ w) {
  const long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= n) return;
  double lo = bounds[0], hi = bounds[1];
  const double pivmin = bounds[2];
  const double atol = 5.6e-17 * bounds[3];     // eps/4 * ||T||: the absolute accuracy a backward-stable solver delivers on rho
  for (int it = 0; it < 64; ++it) {
    if (hi - lo <= atol) break;
    const double step = (hi - lo) * (1.0 / 33.0);
    const double x = lo + step * (double)(lane + 1);
    if (!(step > 0.0) || !(lo + step > lo)) break;
    long long cnt = 0;
    double q = __ldg(d) - x;
    if (fabs(q) < pivmin) q = -pivmin;
    if (q < 0.0) ++cnt;
    for (long long i = 1; i < n; ++i) {
      const double ee = __ldg(e + i - 1);
      q = __ldg(d + i) - x - ee * ee / q;
      if (fabs(q) < pivmin) q = -pivmin;
      if (q < 0.0) ++cnt;
    }
    // counts are monotone in x: lanes 0..t-1 still have at most k eigenvalues below their point
    const int t = __popc(__ballot_sync(0xffffffffu, cnt <= k));
    const double nlo = t == 0 ? lo : lo + step * (double)t;
    const double nhi = t == 32 ? hi : lo + step * (double)(t + 1);
    lo = nlo; hi = nhi;
  }
  if (lane == 0) w[k] = 0.5 * (lo + hi);
}

__device__ __forceinline__ double hash_unit(unsigned long long i, unsigned long long j) {
  unsigned long long z = (i + 1ull) * 0x9E3779B97F4A7C15ull ^ (j + 1ull) * 0xC2B2AE3D27D4EB4Full;
  z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 29; z *= 0x94D049BB133111EBull; z ^= z >> 32;
  return ((double)(z >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}

// Reference form of invit_kernel below (one memory round trip per element and loop; CB200_INVIT=ref selects it):
// thread j -> eigenvector of T for the shift lam[j]: LU with partial pivoting of T - lam (two super-diagonals), three inverse
// iterations from a hashed start vector.  All per-vector arrays are interleaved [i*m + j] so that a warp's accesses coalesce.
__global__ void __launch_bounds__(128) invit_ref_kernel(const double* __restrict__ d, const double* __restrict__ e, long long n,
                                                    const double* __restrict__ lam, long long m, const double* __restrict__ bounds,
                                                    double* __restrict__ Zt, double* __restrict__ work) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const double tol = fmax(2.3e-16 * bounds[3], 1e-300);
  double* u0 = work + j; double* u1 = u0 + n * m; double* u2 = u1 + n * m; double* lm = u2 + n * m; double* sw = lm + n * m;
  double* z = Zt + j;
  const double l = lam[j];
  double p0 = __ldg(d) - l, p1 = n > 1 ? __ldg(e) : 0.0;
  for (long long i = 0; i + 1 < n; ++i) {
    const double r0 = __ldg(e + i), r1 = __ldg(d + i + 1) - l, r2 = (i + 2 < n) ? __ldg(e + i + 1) : 0.0;
    if (fabs(p0) >= fabs(r0)) {
      if (fabs(p0) < tol) p0 = p0 < 0.0 ? -tol : tol;
      const double mult = r0 / p0;
      sw[i * m] = 0.0; lm[i * m] = mult; u0[i * m] = p0; u1[i * m] = p1; u2[i * m] = 0.0;
      p0 = r1 - mult * p1; p1 = r2;
    } else {
      const double mult = p0 / r0;
      sw[i * m] = 1.0; lm[i * m] = mult; u0[i * m] = r0; u1[i * m] = r1; u2[i * m] = r2;
      p0 = p1 - mult * r1; p1 = -mult * r2;
    }
  }
  if (fabs(p0) < tol) p0 = p0 < 0.0 ? -tol : tol;
  u0[(n - 1) * m] = p0; u1[(n - 1) * m] = 0.0; u2[(n - 1) * m] = 0.0;
  for (long long i = 0; i < n; ++i) z[i * m] = hash_unit((unsigned long long)i, (unsigned long long)j);
  for (int it = 0; it < 3; ++it) {
    double cur = z[0];
    for (long long i = 0; i + 1 < n; ++i) {
      double nxt = z[(i + 1) * m];
      if (sw[i * m] != 0.0) { const double t = cur; cur = nxt; nxt = t; }
      z[i * m] = cur;
      cur = nxt - lm[i * m] * cur;
    }
    z[(n - 1) * m] = cur;
    double mx = 0.0, x1 = 0.0, x2 = 0.0;
    for (long long i = n - 1; i >= 0; --i) {
      double y = z[i * m] - u1[i * m] * x1 - u2[i * m] * x2;
      y /= u0[i * m];
      z[i * m] = y;
      x2 = x1; x1 = y;
      mx = fmax(mx, fabs(y));
    }
    const double sc = mx > 0.0 ? 1.0 / mx : 1.0;
    for (long long i = 0; i < n; ++i) z[i * m] *= sc;
  }
  double nn = 0.0;
  for (long long i = 0; i < n; ++i) nn = fma(z[i * m], z[i * m], nn);
  nn = 1.0 / sqrt(nn);
  for (long long i = 0; i < n; ++i) z[i * m] *= nn;
}

// The same arithmetic, operation for operation, with every loop processed in register chunks of INVIT_CH elements: the loads of a
// chunk are independent of the recurrence and are all issued before it runs (the per-vector arrays are strided by the runtime value
// m, so the compiler cannot prove that a store to element i does not alias the load of element i+1 and serialises them: one L2
// round trip per element, 12.7 ms at n = 3200).  Results are bit-identical to invit_ref_kernel.
constexpr int INVIT_CH = 8;
__global__ void __launch_bounds__(128) invit_kernel(const double* __restrict__ d, const double* __restrict__ e, long long n,
                                                    const double* __restrict__ lam, long long m, const double* __restrict__ bounds,
                                                    double* __restrict__ Zt, double* __restrict__ work) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const double tol = fmax(2.3e-16 * bounds[3], 1e-300);
  double* u0 = work + j; double* u1 = u0 + n * m; double* u2 = u1 + n * m; double* lm = u2 + n * m; double* sw = lm + n * m;
  double* z = Zt + j;
  const double l = lam[j];
  double p0 = __ldg(d) - l, p1 = n > 1 ? __ldg(e) : 0.0;
  for (long long i0 = 0; i0 + 1 < n; i0 += INVIT_CH) {
    const int cnt = (int)((n - 1 - i0) < INVIT_CH ? (n - 1 - i0) : INVIT_CH);
    double ee[INVIT_CH + 1], dd[INVIT_CH];
#pragma unroll
    for (int k = 0; k < INVIT_CH; ++k)
      if (k < cnt) { ee[k] = __ldg(e + i0 + k); dd[k] = __ldg(d + i0 + k + 1); }
#pragma unroll
    for (int k = 1; k <= INVIT_CH; ++k)
      if (k == cnt) ee[k] = (i0 + k + 1 < n) ? __ldg(e + i0 + k) : 0.0;
#pragma unroll
    for (int k = 0; k < INVIT_CH; ++k)
      if (k < cnt) {
        const long long i = i0 + k;
        const double r0 = ee[k], r1 = dd[k] - l, r2 = (i + 2 < n) ? ee[k + 1] : 0.0;
        if (fabs(p0) >= fabs(r0)) {
          if (fabs(p0) < tol) p0 = p0 < 0.0 ? -tol : tol;
          const double mult = r0 / p0;
          sw[i * m] = 0.0; lm[i * m] = mult; u0[i * m] = p0; u1[i * m] = p1; u2[i * m] = 0.0;
          p0 = r1 - mult * p1; p1 = r2;
        } else {
          const double mult = p0 / r0;
          sw[i * m] = 1.0; lm[i * m] = mult; u0[i * m] = r0; u1[i * m] = r1; u2[i * m] = r2;
          p0 = p1 - mult * r1; p1 = -mult * r2;
        }
      }
  }
  if (fabs(p0) < tol) p0 = p0 < 0.0 ? -tol : tol;
  u0[(n - 1) * m] = p0; u1[(n - 1) * m] = 0.0; u2[(n - 1) * m] = 0.0;
  for (long long i = 0; i < n; ++i) z[i * m] = hash_unit((unsigned long long)i, (unsigned long long)j);
  for (int it = 0; it < 3; ++it) {
    double cur = z[0];
    for (long long i0 = 0; i0 + 1 < n; i0 += INVIT_CH) {      // forward substitution: reads z[i+1] (old), writes z[i]
      const int cnt = (int)((n - 1 - i0) < INVIT_CH ? (n - 1 - i0) : INVIT_CH);
      double zz[INVIT_CH], ss[INVIT_CH], ll[INVIT_CH];
#pragma unroll
      for (int k = 0; k < INVIT_CH; ++k)
        if (k < cnt) { zz[k] = z[(i0 + k + 1) * m]; ss[k] = sw[(i0 + k) * m]; ll[k] = lm[(i0 + k) * m]; }
#pragma unroll
      for (int k = 0; k < INVIT_CH; ++k)
        if (k < cnt) {
          double nxt = zz[k];
          if (ss[k] != 0.0) { const double t = cur; cur = nxt; nxt = t; }
          z[(i0 + k) * m] = cur;
          cur = nxt - ll[k] * cur;
        }
    }
    z[(n - 1) * m] = cur;
    double mx = 0.0, x1 = 0.0, x2 = 0.0;
    for (long long ih = n - 1; ih >= 0; ih -= INVIT_CH) {       // back substitution, descending
      const int cnt = (int)((ih + 1) < INVIT_CH ? (ih + 1) : INVIT_CH);
      double zz[INVIT_CH], a0[INVIT_CH], a1[INVIT_CH], a2[INVIT_CH];
#pragma unroll
      for (int k = 0; k < INVIT_CH; ++k)
        if (k < cnt) { const long long i = ih - k; zz[k] = z[i * m]; a0[k] = u0[i * m]; a1[k] = u1[i * m]; a2[k] = u2[i * m]; }
#pragma unroll
      for (int k = 0; k < INVIT_CH; ++k)
        if (k < cnt) {
          double y = zz[k] - a1[k] * x1 - a2[k] * x2;
          y /= a0[k];
          z[(ih - k) * m] = y;
          x2 = x1; x1 = y;
          mx = fmax(mx, fabs(y));
        }
    }
    const double sc = mx > 0.0 ? 1.0 / mx : 1.0;
    for (long long i0 = 0; i0 < n; i0 += INVIT_CH) {
      double zz[INVIT_CH];
#pragma unroll
      for (int k = 0; k < INVIT_CH; ++k)
        if (i0 + k < n) zz[k] = z[(i0 + k) * m];
#pragma unroll
      for (int k = 0; k < INVIT_CH; ++k)
        if (i0 + k < n) z[(i0 + k) * m] = zz[k] * sc;
    }
  }
  double nn = 0.0;
  for (long long i0 = 0; i0 < n; i0 += INVIT_CH) {
    double zz[INVIT_CH];
#pragma unroll
    for (int k = 0; k < INVIT_CH; ++k)
      if (i0 + k < n) zz[k] = z[(i0 + k) * m];
#pragma unroll
    for (int k = 0; k < INVIT_CH; ++k)
      if (i0 + k < n) nn = fma(zz[k], zz[k], nn);
  }
  nn = 1.0 / sqrt(nn);
  for (long long i0 = 0; i0 < n; i0 += INVIT_CH) {
    double zz[INVIT_CH];
#pragma unroll
    for (int k = 0; k < INVIT_CH; ++k)
      if (i0 + k < n) zz[k] = z[(i0 + k) * m];
#pragma unroll
    for (int k = 0; k < INVIT_CH; ++k)
      if (i0 + k < n) z[(i0 + k) * m] = zz[k] * nn;
  }
}

// Newton-Schulz coefficient matrix C = alpha (1.5 I - 0.5 alpha^2 G), one block per column; stats (as ordered uint64 bit patterns
// of non-negative doubles): [0] max |G - I|, [1] max column sum of |G|
__global__ void __launch_bounds__(256) ns_coeff_kernel(const double* __restrict__ G, long long m, double alpha, double* __restrict__ Cout,
                                                       unsigned long long* __restrict__ stats) {
  const long long c = blockIdx.x;
  double dev = 0.0, cs = 0.0;
  for (long long r = threadIdx.x; r < m; r += 256) {
    const double g = G[r + c * m];
    dev = fmax(dev, fabs(g - (r == c ? 1.0 : 0.0)));
    cs += fabs(g);
    Cout[r + c * m] = alpha * ((r == c ? 1.5 : 0.0) - 0.5 * alpha * alpha * g);
  }
  __shared__ double sd[256], ss[256];
  sd[threadIdx.x] = dev; ss[threadIdx.x] = cs;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sd[threadIdx.x] = fmax(sd[threadIdx.x], sd[threadIdx.x + o]); ss[threadIdx.x] += ss[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (!(sd[0] == sd[0])) sd[0] = 1e300;   // NaN -> reported as a huge deviation
    atomicMax(stats, (unsigned long long)__double_as_longlong(sd[0]));
    atomicMax(stats + 1, (unsigned long long)__double_as_longlong(ss[0]));
  }
}

// V[:, c] = H_0 H_1 ... H_{n-2} z_c: one warp per column, the column lives in shared memory (or in V itself when it does not fit)
template <bool CPLX>
__global__ void backtransform_kernel(const void* __restrict__ A_, long long n, const void* __restrict__ tau_, const void* __restrict__ scl_,
                                     const double* __restrict__ Zt, long long m, void* __restrict__ V_, int use_smem) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  extern __shared__ __align__(16) char bsm[];
  const T* A = reinterpret_cast<const T*>(A_);
  const T* tau = reinterpret_cast<const T*>(tau_);
  const T* scl = reinterpret_cast<const T*>(scl_);
  T* V = reinterpret_cast<T*>(V_);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long c = (long long)blockIdx.x * nw + warp;
  if (c >= m) return;
  T* z = use_smem ? reinterpret_cast<T*>(bsm) + (size_t)warp * n : V + c * n;
  const T one = t_one(T());
  for (long long i = lane; i < n; i += 32) z[i] = t_scale(one, Zt[i * m + c]);
  __syncwarp();
  for (long long j = n - 2; j >= 0; --j) {
    const T tj = tau[j];
    if (t_is_zero(tj)) continue;
    const T sc = scl[j];
    const T* x = A + (j + 1) + j * n;
    const long long len = n - j - 1;
    T* zz = z + j + 1;
    T t = t_zero(one);
    for (long long i = lane; i < len; i += 32) {
      const T vi = i == 0 ? one : t_mul(x[i], sc);
      t = t_add(t, t_mul(t_conj(vi), zz[i]));
    }
    t = t_mul(tj, t_warp_sum(t));
    for (long long i = lane; i < len; i += 32) {
      const T vi = i == 0 ? one : t_mul(x[i], sc);
      zz[i] = t_sub(zz[i], t_mul(t, vi));
    }
    __syncwarp();
  }
  if (use_smem)
    for (long long i = lane; i < n; i += 32) V[c * n + i] = z[i];
}

// ---- blocked back-transformation helpers (the O(n^2 m) work itself runs on the DMMA GEMM) ----------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256) zt_to_v_kernel(const double* __restrict__ Zt, long long n, long long m, void* __restrict__ V_) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  T* V = reinterpret_cast<T*>(V_);
  // 32 x 32 tiles through shared memory so that both the read ([i*m + c], c fastest) and the write ([i + c*n], i fastest) coalesce
  __shared__ double tile[32][33];
  const long long i0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long long i = i0 + r, c = c0 + tx;
    tile[r][tx] = (i < n && c < m) ? Zt[i * m + c] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const long long c = c0 + r, i = i0 + tx;
    if (i < n && c < m) V[i + c * n] = t_scale(t_one(T()), tile[tx][r]);
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) form_y_kernel(const void* __restrict__ A_, long long n, const void* __restrict__ tau_,
                                                     const void* __restrict__ scl_, long long j0, int nbk, void* __restrict__ Y_) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const T* A = reinterpret_cast<const T*>(A_);
  const T* tau = reinterpret_cast<const T*>(tau_);
  const T* scl = reinterpret_cast<const T*>(scl_);
  T* Y = reinterpret_cast<T*>(Y_);
  const long long r0 = j0 + 1, len = n - r0;
  const int i = blockIdx.y;                 // reflector j0 + i
  const long long j = j0 + i;
  const T sc = scl[j];
  const bool off = t_is_zero(tau[j]);
  for (long long rr = (long long)blockIdx.x * 256 + threadIdx.x; rr < len; rr += (long long)gridDim.x * 256) {
    const long long r = r0 + rr;
    T y = t_zero(T());
    if (!off) {
      if (r == j + 1) y = t_one(T());
      else if (r > j + 1) y = t_mul(A[r + j * n], sc);
    }
    Y[rr + (long long)i * len] = y;
  }
}

template <bool CPLX>
__device__ __forceinline__ void wy_t_body(const void* __restrict__ S_, const void* __restrict__ tau_, int nb, void* __restrict__ Tn_) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const T* S = reinterpret_cast<const T*>(S_);
  const T* tau = reinterpret_cast<const T*>(tau_);
  T* Tn = reinterpret_cast<T*>(Tn_);
  extern __shared__ __align__(16) char wsm[];
  T* Ts = reinterpret_cast<T*>(wsm);
  const int r = threadIdx.x;
  for (int k = r; k < nb * nb; k += 64) Ts[k] = t_zero(T());
  __syncthreads();
  for (int i = 0; i < nb; ++i) {
    const T ti = tau[i];
    if (r < i) {
      T tmp = t_zero(T());
      for (int c = r; c < i; ++c) tmp = t_add(tmp, t_mul(Ts[r + c * nb], S[c + i * nb]));
      Ts[r + i * nb] = t_scale(t_mul(ti, tmp), -1.0);
    } else if (r == i) {
      Ts[i + i * nb] = ti;
    }
    __syncthreads();
  }
  for (int k = r; k < nb * nb; k += 64) Tn[k] = t_scale(Ts[k], -1.0);
}
template <bool CPLX>
__global__ void __launch_bounds__(64) wy_t_kernel(const void* __restrict__ S_, const void* __restrict__ tau_, int nb, void* __restrict__ Tn_) {
  wy_t_body<CPLX>(S_, tau_, nb, Tn_);
}
// all reflector blocks of one matrix at once (they are independent): CTA k builds the T factor of block k from S_k = Y_k^H Y_k stored at
// S + k*64*64 (nb x nb, ld nb) and tau[k*64 ..]; every block has 64 reflectors except the last one (nb_last)
template <bool CPLX>
__global__ void __launch_bounds__(64) wy_t_batched_kernel(const void* __restrict__ S_, const void* __restrict__ tau_, int npanel, int nb_last,
                                                          void* __restrict__ Tn_) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const int k = blockIdx.x;
  wy_t_body<CPLX>(reinterpret_cast<const T*>(S_) + (size_t)k * 4096, reinterpret_cast<const T*>(tau_) + (size_t)k * 64,
                  k == npanel - 1 ? nb_last : 64, reinterpret_cast<T*>(Tn_) + (size_t)k * 4096);
}

// ---- cutoff / maxdim selection on the device (ITensor truncate()) ------------------------------------------
// rank sort: position of pooled[i] in descending order (ties broken by index, so the result is a permutation)
__global__ void __launch_bounds__(256) rank_sort_desc_kernel(const double* __restrict__ in, long long n, double* __restrict__ out) {
  __shared__ double tile[256];
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const double mine = i < n ? in[i] : 0.0;
  long long rank = 0;
  for (long long j0 = 0; j0 < n; j0 += 256) {
    __syncthreads();
    tile[threadIdx.x] = (j0 + threadIdx.x < n) ? in[j0 + threadIdx.x] : -1e308;
    __syncthreads();
    const int lim = (int)((n - j0) < 256 ? (n - j0) : 256);
    for (int t = 0; t < lim; ++t) {
      const double v = tile[t];
      rank += (v > mine) || (v == mine && (j0 + t) < i);
    }
  }
  if (i < n) out[rank] = mine;
}

__global__ void truncate_rule_kernel(double* __restrict__ P, long long origm, long long maxdim, long long mindim, double cutoff,
                                     double* __restrict__ out3) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  long long n = origm - 1;
  for (long long zn = n; zn >= 0 && P[zn] < 0.0; --zn) P[zn] = 0.0;
  if (origm == 1) { out3[0] = 1.0; out3[1] = 0.0; out3[2] = P[0] / 2.0; return; }
  double truncerr = 0.0;
  while (n >= maxdim) { truncerr += P[n]; --n; }
  double scale = 0.0;
  for (long long k = 0; k < origm; ++k) scale += P[k];
  if (scale == 0.0) scale = 1.0;
  while (n >= mindim && n >= 0 && truncerr + P[n] < cutoff * scale) { truncerr += P[n]; --n; }
  truncerr /= scale;
  if (n < 0) n = 0;
  const long long m = n + 1;
  double docut = 0.0;
  if (m < origm) {
    docut = (P[m - 1] + P[m]) / 2.0;
    if (fabs(P[m - 1] - P[m]) < 1e-3 * P[m - 1]) docut += 1e-3 * P[m - 1];
  }
  out3[0] = (double)m; out3[1] = truncerr; out3[2] = docut;
}

}  // namespace cb200
