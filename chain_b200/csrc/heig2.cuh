// Two-stage reduction kernels of the bond truncation's Hermitian eigensolver (ITensor: MPS::svdBond -> denmatDecomp -> diagHermitian,
// reached from /root/reference/src/dmrg.h:544,563,596):
//   stage 1  dense -> band (semi-bandwidth SB_B): panel_qr_kernel factors the block below the band; the two-sided block update of the
//            trailing matrix is five launches of the DMMA GEMM kernel (engine.cpp, sb_stage1) -- the O(n^3) work is tensor-core work;
//   stage 2  band -> tridiagonal: bulge_chase_kernel, ONE persistent launch, a CTA per sweep, sweeps pipelined through progress counters;
//   back     q2_apply_kernel applies the bulge-chasing reflectors to the kept eigenvectors blockwise with the vectors in registers;
//            the QR panels are applied in compact-WY form on the GEMM kernel (engine.cpp, heig_vectors).
// Conventions: backend.h.  Reflectors follow LAPACK larfg: H = I - tau v v^H, v[0] = 1, H^H x = beta e_1, beta real.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include "backend.h"
#include "heig.cuh"

namespace cb200 {

template <class T>
__device__ __forceinline__ T t_ldcg(const T* p);
template <>
__device__ __forceinline__ double t_ldcg<double>(const double* p) { return __ldcg(p); }
template <>
__device__ __forceinline__ double2 t_ldcg<double2>(const double2* p) { return __ldcg(p); }

__device__ __forceinline__ double t_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double2 t_shfl_xor(double2 v, int m) { return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m)); }

// larfg scalars from alpha = x[0] and sigma = |x[1:]|^2: returns tau, sets beta and scale (v[i] = x[i] * scale, i >= 1); tau = 0 <=> H = I
template <class T>
__device__ __forceinline__ T larfg_scalars(T alpha, double sigma, double* beta, T* scale) {
  const double ar = t_re(alpha), ai = t_im(alpha);
  if (sigma == 0.0 && ai == 0.0) { *beta = ar; *scale = t_zero(alpha); return t_zero(alpha); }
  const double b = -copysign(sqrt(ar * ar + ai * ai + sigma), ar);
  *beta = b;
  *scale = t_inv(t_sub(alpha, t_make(b, 0.0, alpha)));
  return t_make((b - ar) / b, -ai / b, alpha);
}

// ===================================================================================================== stage 1: panel QR
constexpr int PQ_THREADS = 256;
constexpr int PQ_MAXROWS_REAL = 192;     // rows of the panel a CTA keeps in shared memory
constexpr int PQ_MAXROWS_CPLX = 128;
__host__ __device__ constexpr int pq_ldp(int R) { return R | 1; }
template <bool CPLX>
constexpr size_t pq_smem_bytes(int R) {
  return (size_t)(CPLX ? 16 : 8) * ((size_t)pq_ldp(R) * SB_B + (size_t)SB_B * SB_B + 4 * SB_B + 3 * SB_B);
}

// Cooperative launch; CTA b owns panel rows [b*R, (b+1)*R) and keeps them in shared memory for the whole panel.  Per column ONE grid
// barrier: every CTA publishes its partial sums  sum_{r > c} conj(P[r,c]) P[r,j]  (j >= c: Gram row of the current column; j < c: the
// products with the earlier reflectors that the T factor needs), the owner of row c publishes that row, and after the barrier every CTA
// reduces the partials in the same fixed order, so all CTAs (and all replicated ranks) form bit-identical Householder scalars.
template <bool CPLX>
__global__ void __launch_bounds__(PQ_THREADS, 1)
panel_qr_kernel(void* __restrict__ A_, long long n, long long j0, int R, void* __restrict__ Y_, void* __restrict__ Tn_, void* __restrict__ Tnh_,
                void* __restrict__ part_, void* __restrict__ rowc_) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  constexpr int B = SB_B;
  T* A = reinterpret_cast<T*>(A_);
  T* Y = reinterpret_cast<T*>(Y_);
  T* part = reinterpret_cast<T*>(part_);     // [2][gridDim.x][B]
  T* rowc = reinterpret_cast<T*>(rowc_);     // [2][B]
  const long long r0 = j0 + B, len = n - r0;
  const int kp = (int)(len - 1 < B ? len - 1 : B);
  const int LDP = pq_ldp(R);
  extern __shared__ __align__(16) char smem_raw[];
  T* Ps = reinterpret_cast<T*>(smem_raw);          // [B][LDP]
  T* Ts = Ps + (size_t)LDP * B;                    // [B][B]  T factor (column-major, ld B)
  T* red = Ts + B * B;                             // [4][B]
  T* dots = red + 4 * B;                           // [B]
  T* rc = dots + B;                                // [B]
  T* zv = rc + B;                                  // [B]
  const int tid = threadIdx.x;
  const int j = tid & (B - 1), q = tid >> 6;       // column j, row quarter q
  const long long row0 = (long long)blockIdx.x * R;
  const int nrows = (int)(len - row0 < R ? (len - row0 > 0 ? len - row0 : 0) : R);
  const T zero = t_zero(T());
  // load the CTA's rows of the panel
  for (int idx = tid; idx < R * B; idx += PQ_THREADS) {
    const int r = idx % R, c = idx / R;
    Ps[r + c * LDP] = r < nrows ? A[(r0 + row0 + r) + (j0 + c) * n] : zero;
  }
  for (int idx = tid; idx < B * B; idx += PQ_THREADS) Ts[idx] = zero;
  __syncthreads();
  const int rq0 = (nrows * q) / 4, rq1 = (nrows * (q + 1)) / 4;    // this thread's row slice
  for (int c = 0; c < kp; ++c) {
    const int buf = c & 1;
    // ---- partial sums over local rows with global row index > c
    {
      T acc = zero;
      int rb = rq0;
      const long long first = (long long)c + 1 - row0;              // local index of the first row > c
      if (first > rb) rb = first > rq1 ? rq1 : (int)first;
      const T* pc = Ps + c * LDP;
      const T* pj = Ps + j * LDP;
      if (j >= c) { for (int r = rb; r < rq1; ++r) acc = t_add(acc, t_mul(t_conj(pc[r]), pj[r])); }
      else { for (int r = rb; r < rq1; ++r) acc = t_add(acc, t_mul(t_conj(pj[r]), pc[r])); }
      red[q * B + j] = acc;
    }
    __syncthreads();
    if (tid < B) {
      const T s = t_add(t_add(red[tid], red[B + tid]), t_add(red[2 * B + tid], red[3 * B + tid]));
      part[((size_t)buf * gridDim.x + blockIdx.x) * B + tid] = s;
      const long long lc = (long long)c - row0;
      if (lc >= 0 && lc < nrows) rowc[buf * B + tid] = Ps[lc + tid * LDP];
    }
    grid.sync();
    // ---- fixed-order reduction of every CTA's partials (redundantly in every CTA)
    {
      T acc = zero;
      const T* pp = part + (size_t)buf * gridDim.x * B + j;
      for (int b = q; b < (int)gridDim.x; b += 4) acc = t_add(acc, t_ldcg(pp + (size_t)b * B));
      red[q * B + j] = acc;
    }
    __syncthreads();
    if (tid < B) {
      dots[tid] = t_add(t_add(red[tid], red[B + tid]), t_add(red[2 * B + tid], red[3 * B + tid]));
      rc[tid] = t_ldcg(rowc + buf * B + tid);
    }
    __syncthreads();
    const T alpha = rc[c];
    double beta;
    T scale;
    const T tau = larfg_scalars(alpha, t_re(dots[c]), &beta, &scale);
    const T ctau = t_conj(tau);
    // ---- T factor column c (larft): z_k = conj(Y[c,k]) + scale * sum_{r>c} conj(Y[r,k]) P[r,c],  T(0:c,c) = -tau T(0:c,0:c) z
    if (tid < c) zv[tid] = t_add(t_conj(rc[tid]), t_mul(scale, dots[tid]));
    // ---- apply H_c^H to the columns to the right: P_j -= conj(tau) (v^H P_j) v,  v^H P_j = P[c,j] + conj(scale) dots_j
    if (j > c && !t_is_zero(tau)) {
      const T f = t_mul(ctau, t_add(rc[j], t_mul(t_conj(scale), dots[j])));
      const T fs = t_mul(f, scale);
      const T* pc = Ps + c * LDP;
      T* pj = Ps + j * LDP;
      for (int r = rq0; r < rq1; ++r) {
        const long long gr = row0 + r;
        if (gr > c) pj[r] = t_sub(pj[r], t_mul(fs, pc[r]));
        else if (gr == c) pj[r] = t_sub(pj[r], f);
      }
    }
    __syncthreads();
    if (tid < c) {
      T acc = zero;
      for (int k = tid; k < c; ++k) acc = t_add(acc, t_mul(Ts[tid + k * B], zv[k]));
      Ts[tid + c * B] = t_sub(zero, t_mul(tau, acc));
    }
    if (tid == c) Ts[c + c * B] = tau;
    // ---- column c itself: beta on the diagonal, v below it
    if (j == c) {
      T* pc = Ps + c * LDP;
      for (int r = rq0; r < rq1; ++r) {
        const long long gr = row0 + r;
        if (gr > c) pc[r] = t_mul(pc[r], scale);
        else if (gr == c) pc[r] = t_make(beta, 0.0, T());
      }
    }
    __syncthreads();
  }
  // ---- outputs: Y (explicit reflectors), R over the panel, T factors
  for (int idx = tid; idx < R * B; idx += PQ_THREADS) {
    const int r = idx % R, c = idx / R;
    if (r >= nrows) continue;
    const long long gr = row0 + r;
    const T v = Ps[r + c * LDP];
    T y = zero, a = v;
    if (c < kp) {
      if (gr > c) { y = v; a = zero; }
      else if (gr == c) y = t_one(T());
    }
    Y[gr + (long long)c * len] = y;
    A[(r0 + gr) + (j0 + c) * n] = a;
  }
  if (blockIdx.x == 0) {
    T* Tn = reinterpret_cast<T*>(Tn_);
    T* Tnh = reinterpret_cast<T*>(Tnh_);
    for (int idx = tid; idx < B * B; idx += PQ_THREADS) {
      const T t = Ts[idx];
      Tn[idx] = t_sub(zero, t);
      Tnh[idx] = t_scale(t, -0.5);
    }
  }
}

template <bool CPLX>
__global__ void band_extract_kernel(const void* __restrict__ A_, long long n, void* __restrict__ AB_) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const T* A = reinterpret_cast<const T*>(A_);
  T* AB = reinterpret_cast<T*>(AB_);
  const long long total = n * SB_LD;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long jj = idx / SB_LD;
    const int dd = (int)(idx % SB_LD);
    const long long i = jj + dd;
    AB[idx] = (dd <= SB_B && i < n) ? A[i + jj * n] : t_zero(T());
  }
}

// ===================================================================================================== stage 2: bulge chasing
constexpr int BC_THREADS = 256;
constexpr int BC_LD = SB_B + 1;
template <bool CPLX>
constexpr size_t bc_smem_bytes() { return (size_t)(CPLX ? 16 : 8) * (2 * (size_t)SB_B * BC_LD + 4 * SB_B + 4 * SB_B) + 64; }

struct ChaseDesc {          // device copy of ChaseProblem + derived counts
  void* AB;
  long long n;
  double* d;
  double* e;
  void* V2;
  void* tau2;
  int* prog;
};

// One CTA per sweep (sweeps handed out in order through an atomic ticket, so every sweep a CTA can wait for is owned by a resident CTA).
// Task (s, k): D = A[c0:c0+B, c0:c0+B] <- H^H D H;  Bk = A[c0+B:c0+2B, c0:c0+B] <- Bk H;  next reflector H' from the first column of Bk;
// Bk <- H'^H Bk.  It may start once sweep s-1 has finished task k+2.
template <bool CPLX>
__global__ void __launch_bounds__(BC_THREADS, 1)
bulge_chase_kernel(const ChaseDesc* __restrict__ probs, int nprob, long long max_sweeps, unsigned long long* __restrict__ ticket) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int B = SB_B, LD = BC_LD;
  extern __shared__ __align__(16) char smem_raw[];
  T* Ds = reinterpret_cast<T*>(smem_raw);       // [B][LD]
  T* Bs = Ds + B * LD;                          // [B][LD]
  T* pr = Bs + B * LD;                          // [4][B] partial sums
  T* v = pr + 4 * B;                            // [B]
  T* vn = v + B;                                // [B]
  T* w = vn + B;                                // [B]
  T* y = w + B;                                 // [B]
  __shared__ unsigned long long s_ticket;
  __shared__ double s_dbl[4];
  __shared__ T s_t[4];
  const int tid = threadIdx.x;
  const int r = tid & (B - 1), q = tid >> 6;
  const int lane = tid & 31, warp = tid >> 5;
  const T zero = t_zero(T());
  for (;;) {
    if (tid == 0) s_ticket = atomicAdd(ticket, 1ull);
    __syncthreads();
    const unsigned long long tk = s_ticket;
    __syncthreads();
    if (tk >= (unsigned long long)max_sweeps * nprob) break;
    const ChaseDesc P = probs[tk % nprob];
    const long long s = (long long)(tk / nprob);
    const long long n = P.n;
    if (s + 1 >= n) continue;
    T* AB = reinterpret_cast<T*>(P.AB);
    const long long K0 = (n - 1 + B - 1) / B;
    const long long g = s / B;
    const int sp = (int)(s % B);
    const long long nk_prev = s > 0 ? (n - 1 - (s - 1) + B - 1) / B : 0;
    T tau = zero;
    for (long long k = 0;; ++k) {
      const long long c0 = s + 1 + k * B;
      if (c0 > n - 1) break;
      const int nd = (int)(n - c0 < B ? n - c0 : B);
      const long long rem = n - c0 - B;
      const int nb2 = (int)(rem <= 0 ? 0 : (rem < B ? rem : B));
      // ---- wait for sweep s-1 to be three tasks ahead (or finished)
      if (s > 0) {
        if (tid == 0) {
          const int need = (int)(k + 3 < nk_prev ? k + 3 : nk_prev);
          int seen;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(P.prog + (s - 1)) : "memory");
          } while (seen < need);
          __threadfence();
        }
        __syncthreads();
      }
      // ---- load D (Hermitian, lower triangle stored) and Bk
      for (int jj = q; jj < B; jj += 4) {
        T a = zero;
        if (r >= jj && r < nd && jj < nd) a = t_ldcg(AB + (r - jj) + (c0 + jj) * SB_LD);
        if (r >= jj) { Ds[r + jj * LD] = a; if (r > jj) Ds[jj + r * LD] = t_conj(a); }
        T b = zero;
        if (r < nb2 && jj < nd) b = t_ldcg(AB + (B + r - jj) + (c0 + jj) * SB_LD);
        Bs[r + jj * LD] = b;
      }
      if (k == 0) {
        // reflector of the sweep from column s: x = A[c0 : c0+nd, s]
        T x = zero;
        if (tid < B) {
          if (tid < nd) x = t_ldcg(AB + (c0 + tid - s) + s * SB_LD);
          double part = tid >= 1 ? t_norm2(x) : 0.0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
          if (lane == 0) s_dbl[warp] = part;
          v[tid] = x;
        }
        __syncthreads();
        double beta;
        T scale;
        tau = larfg_scalars(v[0], s_dbl[0] + s_dbl[1], &beta, &scale);     // redundantly in every thread
        __syncthreads();
        if (tid < B) {
          T vi = tid == 0 ? t_one(T()) : t_mul(x, scale);
          if (tid >= nd) vi = zero;
          v[tid] = vi;
          if (tid < nd) AB[(c0 + tid - s) + s * SB_LD] = tid == 0 ? t_make(beta, 0.0, T()) : zero;
          if (tid == 0) {
            P.d[s] = t_re(t_ldcg(AB + s * SB_LD));
            P.e[s] = beta;
          }
        }
      }
      __syncthreads();
      // ---- store the reflector for the back-transformation
      {
        const long long blk = g * K0 - g * (g - 1) / 2 + k;
        if (tid < nd) reinterpret_cast<T*>(P.V2)[blk * (long long)SB_VLD * B + (long long)sp * SB_VLD + sp + tid] = v[tid];
        if (tid == 0) reinterpret_cast<T*>(P.tau2)[blk * B + sp] = tau;
      }
      // ---- D <- H^H D H:  p = tau D v,  w = p - 1/2 conj(tau) (v^H p) v,  D -= v w^H + w v^H
      {
        T acc = zero;
#pragma unroll 4
        for (int jj = q * 16; jj < q * 16 + 16; ++jj) acc = t_add(acc, t_mul(Ds[r + jj * LD], v[jj]));
        pr[q * B + r] = acc;
      }
      __syncthreads();
      if (tid < B) {
        const T p = t_mul(tau, t_add(t_add(pr[tid], pr[B + tid]), t_add(pr[2 * B + tid], pr[3 * B + tid])));
        w[tid] = p;
        T vp = t_mul(t_conj(v[tid]), p);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vp = t_add(vp, t_shfl_down(vp, o));
        if (lane == 0) s_t[1 + warp] = vp;
      }
      __syncthreads();
      if (tid < B) {
        const T a2 = t_scale(t_mul(t_conj(tau), t_add(s_t[1], s_t[2])), -0.5);
        w[tid] = t_add(w[tid], t_mul(a2, v[tid]));
      }
      __syncthreads();
      {
        const T vr = v[r], wr = w[r];
        for (int jj = q; jj < B; jj += 4) {
          if (r >= jj && r < nd) {
            const T a = t_sub(Ds[r + jj * LD], t_add(t_mul(vr, t_conj(w[jj])), t_mul(wr, t_conj(v[jj]))));
            AB[(r - jj) + (c0 + jj) * SB_LD] = a;
            if (r == jj && c0 + r == n - 1) P.d[n - 1] = t_re(a);
          }
        }
      }
      if (nb2 <= 0) {
        // sweep finished: publish
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          const int done = (int)(k + 1);
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(P.prog + s), "r"(done) : "memory");
        }
        break;
      }
      // ---- Bk <- Bk H:  z = tau Bk v,  Bk -= z v^H
      {
        T acc = zero;
#pragma unroll 4
        for (int jj = q * 16; jj < q * 16 + 16; ++jj) acc = t_add(acc, t_mul(Bs[r + jj * LD], v[jj]));
        pr[q * B + r] = acc;
      }
      __syncthreads();
      if (tid < B) w[tid] = t_mul(tau, t_add(t_add(pr[tid], pr[B + tid]), t_add(pr[2 * B + tid], pr[3 * B + tid])));
      __syncthreads();
      {
        const T zr = w[r];
        for (int jj = q; jj < B; jj += 4) Bs[r + jj * LD] = t_sub(Bs[r + jj * LD], t_mul(zr, t_conj(v[jj])));
      }
      __syncthreads();
      // ---- next reflector from the first column of Bk
      if (tid < B) {
        const T x = Bs[tid];
        double part = (tid >= 1 && tid < nb2) ? t_norm2(x) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
        if (lane == 0) s_dbl[warp] = part;
      }
      __syncthreads();
      T taun;
      {
        double beta;
        T scale;
        taun = larfg_scalars(Bs[0], s_dbl[0] + s_dbl[1], &beta, &scale);
        if (tid < B) {
          T vi = tid == 0 ? t_one(T()) : t_mul(Bs[tid], scale);
          if (tid >= nb2) vi = zero;
          vn[tid] = vi;
        }
        __syncthreads();
        // first column becomes (beta, 0, ...)
        if (tid < B) Bs[tid] = tid == 0 ? t_make(beta, 0.0, T()) : zero;
      }
      // ---- Bk <- H'^H Bk (columns 1..):  y_j = conj(tau') vn^H Bk[:,j],  Bk[:,j] -= vn y_j      (thread (j = r, q): rows 16q..16q+15)
      {
        T acc = zero;
        if (r >= 1) {
#pragma unroll 4
          for (int i = q * 16; i < q * 16 + 16; ++i) acc = t_add(acc, t_mul(t_conj(vn[i]), Bs[i + r * LD]));
        }
        pr[q * B + r] = acc;
      }
      __syncthreads();
      if (tid < B) y[tid] = tid >= 1 ? t_mul(t_conj(taun), t_add(t_add(pr[tid], pr[B + tid]), t_add(pr[2 * B + tid], pr[3 * B + tid]))) : zero;
      __syncthreads();
      {
        const T vr = vn[r];
        for (int jj = q; jj < B; jj += 4) {
          if (r < nb2 && jj < nd) {
            const T b = jj == 0 ? Bs[r] : t_sub(Bs[r + jj * LD], t_mul(vr, y[jj]));
            AB[(B + r - jj) + (c0 + jj) * SB_LD] = b;
          }
        }
      }
      __syncthreads();
      if (tid < B) v[tid] = vn[tid];
      tau = taun;
      if (tid == 0) {
        __threadfence();
        const int done = (int)(k + 1);
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(P.prog + s), "r"(done) : "memory");
      }
      __syncthreads();
    }
  }
}

// ===================================================================================================== back-transformation with Q2
constexpr int Q2_THREADS = 256;
constexpr int Q2_NC = 32;                 // columns of Z per CTA: 8 warps x 4 columns, 8 lanes per column
constexpr int Q2_VROW = SB_B + 16 + 1;    // compact reflector row: 8 zeros | 64 entries | 8 zeros (+1 pad)
template <bool CPLX>
constexpr size_t q2_smem_bytes() { return (size_t)(CPLX ? 16 : 8) * (2 * (size_t)SB_B * Q2_VROW + 2 * SB_B); }

template <class T>
__device__ __forceinline__ void q2_load_block(T* Vs, T* taus, const T* __restrict__ V2, const T* __restrict__ tau2, long long blk, int tid) {
  // compact copy of the band of the reflector block: Vs[sp][8 + i] = V2blk[sp*SB_VLD + sp + i]
  const T* src = V2 + blk * (long long)SB_VLD * SB_B;
  for (int idx = tid; idx < SB_B * SB_B; idx += Q2_THREADS) {
    const int sp = idx >> 6, i = idx & 63;
    Vs[sp * Q2_VROW + 8 + i] = src[sp * SB_VLD + sp + i];
  }
  if (tid < SB_B) taus[tid] = tau2[blk * SB_B + tid];
}

// lane layout inside a warp: column cl = lane >> 3 (4 columns), part = lane & 7; the thread owns window rows a = part + 8 k, k = 0..15, of
// its column in registers.  Reflector sp of a block acts on window rows sp .. sp+63: with sp = 8 kb + sm its rows in this thread are
// k = kb .. kb+8 with v index i = part + 8 (k - kb) - sm in [-7, 71]; the zero padding of the compact rows makes the two partial ends exact.
template <bool CPLX>
__global__ void __launch_bounds__(Q2_THREADS, 1)
q2_apply_kernel(const void* __restrict__ V2_, const void* __restrict__ tau2_, long long n, void* __restrict__ Z_, long long m) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int B = SB_B;
  const T* V2 = reinterpret_cast<const T*>(V2_);
  const T* tau2 = reinterpret_cast<const T*>(tau2_);
  T* Z = reinterpret_cast<T*>(Z_);
  extern __shared__ __align__(16) char smem_raw[];
  T* Vbuf = reinterpret_cast<T*>(smem_raw);            // [2][B][Q2_VROW]
  T* tbuf = Vbuf + 2 * B * Q2_VROW;                    // [2][B]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = lane & 7;
  const long long col = (long long)blockIdx.x * Q2_NC + warp * 4 + (lane >> 3);
  const bool live = col < m;
  T* zc = Z + (live ? col : 0) * n;
  const T zero = t_zero(T());
  for (int idx = tid; idx < 2 * B * Q2_VROW; idx += Q2_THREADS) Vbuf[idx] = zero;
  __syncthreads();
  const long long K0 = (n - 1 + B - 1) / B;
  T z[16];
  for (long long g = K0 - 1; g >= 0; --g) {
    const long long nk = K0 - g;
    const long long blk0 = g * K0 - g * (g - 1) / 2;
    long long base = g * B + 1;                         // first row of the window of block (g, 0)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const long long row = base + part + 8 * k;
      z[k] = (live && row < n) ? zc[row] : zero;
    }
    q2_load_block(Vbuf, tbuf, V2, tau2, blk0, tid);
    __syncthreads();
    for (long long k = 0; k < nk; ++k) {
      const int cur = (int)(k & 1);
      // prefetch the next block of reflectors and the next 64 rows of Z while this block is applied
      T znext[8];
      if (k + 1 < nk) {
        q2_load_block(Vbuf + (cur ^ 1) * B * Q2_VROW, tbuf + (cur ^ 1) * B, V2, tau2, blk0 + k + 1, tid);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const long long row = base + 128 + part + 8 * kk;
          znext[kk] = (live && row < n) ? zc[row] : zero;
        }
      }
      const T* Vs = Vbuf + cur * B * Q2_VROW;
      const T* ts = tbuf + cur * B;
#pragma unroll
      for (int sp = B - 1; sp >= 0; --sp) {
        const T tau = ts[sp];
        if (t_is_zero(tau)) continue;
        const int kb = sp >> 3, sm = sp & 7;
        const T* vrow = Vs + sp * Q2_VROW + 8 + part - sm;
        T vv[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) vv[t] = vrow[8 * t];
        T acc = zero;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc = t_add(acc, t_mul(t_conj(vv[t]), z[kb + t]));
        acc = t_add(acc, t_shfl_xor(acc, 1));
        acc = t_add(acc, t_shfl_xor(acc, 2));
        acc = t_add(acc, t_shfl_xor(acc, 4));
        acc = t_mul(tau, acc);
#pragma unroll
        for (int t = 0; t < 9; ++t) z[kb + t] = t_sub(z[kb + t], t_mul(acc, vv[t]));
      }
      // rows 0..63 of the window are final for this sweep group: store them, slide the window down by 64 rows
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const long long row = base + part + 8 * kk;
        if (live && row < n) zc[row] = z[kk];
      }
      if (k + 1 < nk) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { z[kk] = z[kk + 8]; z[kk + 8] = znext[kk]; }
        base += B;
      } else {
#pragma unroll
        for (int kk = 8; kk < 16; ++kk) {
          const long long row = base + part + 8 * kk;
          if (live && row < n) zc[row] = z[kk];
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace cb200
