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
// Bk <- H'^H Bk.  It may start once sweep s-1 has finished task k+1.
__device__ __forceinline__ void chase_wait(const int* flag, int need) {
  int seen;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
  } while (seen < need);
}

template <bool CPLX>
__global__ void __launch_bounds__(BC_THREADS, 1)
bulge_chase_kernel(const ChaseDesc* __restrict__ probs, int nprob, long long max_sweeps, unsigned long long* __restrict__ ticket) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int B = SB_B, LD = BC_LD;
  extern __shared__ __align__(16) char smem_raw[];
  T* Ds = reinterpret_cast<T*>(smem_raw);       // [B][LD]
  T* Bs = Ds + B * LD;                          // [B][LD]
  T* pr = Bs + B * LD;                          // [4][B] partial sums (D v, then vn^H B)
  T* v = pr + 4 * B;                            // [B] current reflector
  T* vn = v + B;                                // [B] next reflector
  T* w = vn + B;                                // [B]
  T* y = w + B;                                 // [B]
  __shared__ T zr[4 * SB_B];                    // partial sums of Bk v
  __shared__ T zz[SB_B];                        // z = tau Bk v
  __shared__ T xs[SB_B];                        // first column of Bk H
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
      // ---- (1) storage columns c0 .. c0+B-2 were last written by task (s-1, k): already awaited (k >= 1: by the previous task of this
      // sweep; k = 0: here).  Fetch them -- and column s of the first task -- BEFORE waiting for task (s-1, k+1), which only owns the
      // last column: the L2 round trip of the block load hides behind the wait.
      if (s > 0 && k == 0) {
        if (tid == 0) { chase_wait(P.prog + (s - 1), (int)(1 < nk_prev ? 1 : nk_prev)); __threadfence(); }
        __syncthreads();
      }
      T da[16], db[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int jj = q + 4 * t;
        da[t] = zero; db[t] = zero;
        if (jj < B - 1) {
          if (r >= jj && r < nd && jj < nd) da[t] = t_ldcg(AB + (r - jj) + (c0 + jj) * SB_LD);
          if (r < nb2 && jj < nd) db[t] = t_ldcg(AB + (B + r - jj) + (c0 + jj) * SB_LD);
        }
      }
      T x = zero;
      if (k == 0 && tid < nd) x = t_ldcg(AB + (c0 + tid - s) + s * SB_LD);
      // ---- (2) wait for task (s-1, k+1) (or the end of sweep s-1): it shares index c0+B-1 with this task; task (s-1, k+2) touches storage
      // columns >= c0+2B-1 only, so a lag of two tasks suffices
      if (s > 0) {
        if (tid == 0) { chase_wait(P.prog + (s - 1), (int)(k + 2 < nk_prev ? k + 2 : nk_prev)); __threadfence(); }
        __syncthreads();
      }
      if (q == 3) {           // the last column (jj = 63 = 3 + 4*15)
        constexpr int jj = B - 1;
        if (r >= jj && r < nd && jj < nd) da[15] = t_ldcg(AB + (r - jj) + (c0 + jj) * SB_LD);
        if (r < nb2 && jj < nd) db[15] = t_ldcg(AB + (B + r - jj) + (c0 + jj) * SB_LD);
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int jj = q + 4 * t;
        if (r >= jj) { Ds[r + jj * LD] = da[t]; if (r > jj) Ds[jj + r * LD] = t_conj(da[t]); }
        Bs[r + jj * LD] = db[t];
      }
      if (k == 0) {
        // reflector of the sweep from column s: x = A[c0 : c0+nd, s]
        if (tid < B) {
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
      // ---- (3) p = D v and z = Bk v (partial sums over 16 columns per thread)
      {
        T ap = zero, az = zero;
#pragma unroll 4
        for (int jj = q * 16; jj < q * 16 + 16; ++jj) {
          const T vj = v[jj];
          ap = t_add(ap, t_mul(Ds[r + jj * LD], vj));
          az = t_add(az, t_mul(Bs[r + jj * LD], vj));
        }
        pr[q * B + r] = ap;
        zr[q * B + r] = az;
      }
      __syncthreads();
      if (tid < B) {
        const T p = t_mul(tau, t_add(t_add(pr[tid], pr[B + tid]), t_add(pr[2 * B + tid], pr[3 * B + tid])));
        const T z = t_mul(tau, t_add(t_add(zr[tid], zr[B + tid]), t_add(zr[2 * B + tid], zr[3 * B + tid])));
        w[tid] = p;
        zz[tid] = z;
        const T xr = t_sub(Bs[tid], z);                    // first column of Bk H  (conj(v[0]) = 1)
        xs[tid] = xr;
        T vp = t_mul(t_conj(v[tid]), p);
        double part = (tid >= 1 && tid < nb2) ? t_norm2(xr) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { vp = t_add(vp, t_shfl_down(vp, o)); part += __shfl_down_sync(0xffffffffu, part, o); }
        if (lane == 0) { s_t[warp] = vp; s_dbl[warp] = part; }
      }
      __syncthreads();
      // ---- (4) scalars, redundantly in every thread: w = p - 1/2 conj(tau) (v^H p) v;  next reflector from the first column of Bk H
      const T a2 = t_scale(t_mul(t_conj(tau), t_add(s_t[0], s_t[1])), -0.5);
      double betan = 0.0;
      T scalen = zero, taun = zero;
      if (nb2 > 0) taun = larfg_scalars(xs[0], s_dbl[0] + s_dbl[1], &betan, &scalen);
      __syncthreads();
      if (tid < B) {
        w[tid] = t_add(w[tid], t_mul(a2, v[tid]));
        T vi = tid == 0 ? t_one(T()) : t_mul(xs[tid], scalen);
        if (tid >= nb2) vi = zero;
        vn[tid] = vi;
      }
      __syncthreads();
      // ---- (5) D <- D - v w^H - w v^H straight to global;  partial sums of vn^H (Bk H) per column (thread (column r, rows 16q..16q+15))
      {
        const T vr = v[r], wr = w[r];
#pragma unroll 4
        for (int jj = q; jj < B; jj += 4) {
          if (r >= jj && r < nd) {
            const T a = t_sub(Ds[r + jj * LD], t_add(t_mul(vr, t_conj(w[jj])), t_mul(wr, t_conj(v[jj]))));
            AB[(r - jj) + (c0 + jj) * SB_LD] = a;
            if (r == jj && c0 + r == n - 1) P.d[n - 1] = t_re(a);
          }
        }
      }
      if (nb2 <= 0) {
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          const int done = (int)(k + 1);
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(P.prog + s), "r"(done) : "memory");
        }
        break;
      }
      {
        T acc = zero;
        const T cvr = t_conj(v[r]);
#pragma unroll 4
        for (int i = q * 16; i < q * 16 + 16; ++i) acc = t_add(acc, t_mul(t_conj(vn[i]), t_sub(Bs[i + r * LD], t_mul(zz[i], cvr))));
        pr[q * B + r] = acc;
      }
      __syncthreads();
      if (tid < B) y[tid] = tid >= 1 ? t_mul(t_conj(taun), t_add(t_add(pr[tid], pr[B + tid]), t_add(pr[2 * B + tid], pr[3 * B + tid]))) : zero;
      __syncthreads();
      // ---- (6) Bk <- H'^H (Bk H) straight to global: first column (beta', 0, ...), the others Bk - z v^H - vn y
      {
        const T zrr = zz[r], vnr = vn[r];
#pragma unroll 4
        for (int jj = q; jj < B; jj += 4) {
          if (r < nb2 && jj < nd) {
            T b;
            if (jj == 0) b = r == 0 ? t_make(betan, 0.0, T()) : zero;
            else b = t_sub(t_sub(Bs[r + jj * LD], t_mul(zrr, t_conj(v[jj]))), t_mul(vnr, y[jj]));
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
__host__ __device__ constexpr int q2_group() { return CPLX ? 4 : 8; }      // reflectors applied together as one compact-WY factor (complex: register budget)
template <bool CPLX>
constexpr size_t q2_smem_bytes() { return (size_t)(CPLX ? 16 : 8) * (2 * (size_t)SB_B * Q2_VROW + 2 * SB_B * q2_group<CPLX>() + 8); }

// T factors of the groups of Q2_G consecutive reflectors of every block:  H_lo H_lo+1 ... H_hi = I - V T V^H  (larft, forward/columnwise).
// One warp per (block, group); T8[(blk*(64/G) + grp)*G*G + j*G + i] = T[i][j] (upper triangular, stored transposed for conflict-free reads).
template <bool CPLX>
__global__ void __launch_bounds__(256) q2_tprep_kernel(const void* __restrict__ V2_, const void* __restrict__ tau2_, long long nblk, void* __restrict__ T8_) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int B = SB_B, G = q2_group<CPLX>();
  const T* V2 = reinterpret_cast<const T*>(V2_);
  const T* tau2 = reinterpret_cast<const T*>(tau2_);
  T* T8 = reinterpret_cast<T*>(T8_);
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= nblk * (B / G)) return;
  const long long blk = wid / (B / G);
  const int grp = (int)(wid % (B / G));
  const T* Vb = V2 + blk * (long long)SB_VLD * B;
  const T zero = t_zero(T());
  __shared__ T Gs[8][G * G];
  __shared__ T Ts[8][G * G];
  T* Gm = Gs[threadIdx.x >> 5];
  T* Tm = Ts[threadIdx.x >> 5];
  // Gram entries G[j][i] = v_j^H v_i for j < i (reflector lo+i starts i-j rows below reflector lo+j)
  for (int i = 1; i < G; ++i)
    for (int j = 0; j < i; ++j) {
      const int spi = grp * G + i, spj = grp * G + j, sh = i - j;
      const T* vi = Vb + spi * SB_VLD + spi;
      const T* vj = Vb + spj * SB_VLD + spj;
      T acc = zero;
      for (int r = lane; r < B - sh; r += 32) acc = t_add(acc, t_mul(t_conj(vj[sh + r]), vi[r]));
      acc = t_warp_sum(acc);
      if (lane == 0) Gm[j * G + i] = acc;
    }
  __syncwarp();
  if (lane == 0) {
    for (int k = 0; k < G * G; ++k) Tm[k] = zero;
    for (int i = 0; i < G; ++i) {
      const T tau = tau2[blk * B + grp * G + i];
      for (int r = 0; r < i; ++r) {
        T acc = zero;
        for (int k = r; k < i; ++k) acc = t_add(acc, t_mul(Tm[r * G + k], Gm[k * G + i]));
        Tm[r * G + i] = t_sub(zero, t_mul(tau, acc));
      }
      Tm[i * G + i] = tau;
    }
  }
  __syncwarp();
  for (int k = lane; k < G * G; k += 32) {
    const int i = k / G, j = k % G;            // T[i][j] -> slot j*G + i
    T8[(blk * (B / G) + grp) * (G * G) + j * G + i] = Tm[i * G + j];
  }
}

template <int G, class T>
__device__ __forceinline__ void q2_load_block(T* Vs, T* Tg, const T* __restrict__ V2, const T* __restrict__ T8, long long blk, int tid, int nthr) {
  // compact copy of the band of the reflector block: Vs[sp][8 + i] = V2blk[sp*SB_VLD + sp + i]; and the block's T factors.  Asynchronous
  // copies (no register staging): the next block streams in while the current one is applied
  const T* src = V2 + blk * (long long)SB_VLD * SB_B;
  for (int idx = tid; idx < SB_B * SB_B; idx += nthr) {
    const int sp = idx >> 6, i = idx & 63;
    if constexpr (sizeof(T) == 16) cp_async16(Vs + sp * Q2_VROW + 8 + i, src + sp * SB_VLD + sp + i, 16);
    else cp_async8(Vs + sp * Q2_VROW + 8 + i, src + sp * SB_VLD + sp + i, 8);
  }
  for (int idx = tid; idx < SB_B * G; idx += nthr) {
    if constexpr (sizeof(T) == 16) cp_async16(Tg + idx, T8 + blk * (SB_B * G) + idx, 16);
    else cp_async8(Tg + idx, T8 + blk * (SB_B * G) + idx, 8);
  }
  cp_async_commit();
}

// Lane layout inside a warp: column cl = lane >> 3 (4 columns), part = lane & 7; the thread owns window rows a = part + 8 k, k = 0..15, of
// its column in registers.  The 8 reflectors sp = 8 kb .. 8 kb + 7 of a group act on window rows 8 kb .. 8 kb + 70, i.e. on this thread's
// registers k = kb .. kb+8; reflector 8 kb + i needs v index part + 8 (k - kb) - i in [-7, 71]: the zero padding of the compact rows makes
// the two partial ends exact.  A group is applied as z <- z - V (T (V^H z)): the 8 dot products are independent instruction chains (the
// reflector-by-reflector form is one 64-long dependent chain per block and was latency-bound at 13 us per block).
template <bool CPLX>
__global__ void __launch_bounds__(Q2_THREADS, 1)
q2_apply_kernel(const void* __restrict__ V2_, const void* __restrict__ T8_, long long n, void* __restrict__ Z_, long long m) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int B = SB_B, G = q2_group<CPLX>();
  const T* V2 = reinterpret_cast<const T*>(V2_);
  const T* T8 = reinterpret_cast<const T*>(T8_);
  T* Z = reinterpret_cast<T*>(Z_);
  extern __shared__ __align__(16) char smem_raw[];
  T* Vbuf = reinterpret_cast<T*>(smem_raw);            // [2][B][Q2_VROW]
  T* tbuf = Vbuf + 2 * B * Q2_VROW;                    // [2][64/G groups][G*G] (+ slack: lanes part >= G read past their group's rows)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int part = lane & 7, lbase = lane & ~7;
  // blockDim.x / 32 warps of 4 columns each: the host picks the warp count so that the column chunks cover all SMs (a CTA walks ALL blocks
  // sequentially; only the columns are parallel)
  const long long col = (long long)blockIdx.x * (nthr >> 3) + warp * 4 + (lane >> 3);
  const bool live = col < m;
  T* zc = Z + (live ? col : 0) * n;
  const T zero = t_zero(T());
  for (int idx = tid; idx < 2 * B * Q2_VROW; idx += nthr) Vbuf[idx] = zero;
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
    q2_load_block<G>(Vbuf, tbuf, V2, T8, blk0, tid, nthr);
    cp_async_wait<0>();
    __syncthreads();
    for (long long k = 0; k < nk; ++k) {
      const int cur = (int)(k & 1);
      // prefetch the next block of reflectors and the next 64 rows of Z while this block is applied
      T znext[8];
      if (k + 1 < nk) {
        q2_load_block<G>(Vbuf + (cur ^ 1) * B * Q2_VROW, tbuf + (cur ^ 1) * B * G, V2, T8, blk0 + k + 1, tid, nthr);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const long long row = base + 128 + part + 8 * kk;
          znext[kk] = (live && row < n) ? zc[row] : zero;
        }
      }
      const T* Vs = Vbuf + cur * B * Q2_VROW;
      const T* Tb = tbuf + cur * B * G;
#pragma unroll
      for (int gq = B / G - 1; gq >= 0; --gq) {
        // reflectors sp = gq*G + i, i < G:  kb = sp >> 3 is the same for the whole group (G divides 8), sm = (gq*G & 7) + i
        const int kb = (gq * G) >> 3, sm0 = (gq * G) & 7;
        const T* vbase = Vs + (gq * G) * Q2_VROW + 8 + part - sm0;
        // d_i = v_i^H z  (G independent chains)
        T d[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const T* vrow = vbase + i * Q2_VROW - i;
          T acc = zero;
#pragma unroll
          for (int t = 0; t < 9; ++t) acc = t_add(acc, t_mul(t_conj(vrow[8 * t]), z[kb + t]));
          d[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
          d[i] = t_add(d[i], t_shfl_xor(d[i], 1));
          d[i] = t_add(d[i], t_shfl_xor(d[i], 2));
          d[i] = t_add(d[i], t_shfl_xor(d[i], 4));
        }
        // t = T d: lane `part` (< G) forms t_part, then the G values are exchanged inside the 8-lane group
        T tp = zero;
        const T* Tg = Tb + gq * (G * G);
#pragma unroll
        for (int j = 0; j < G; ++j) tp = t_add(tp, t_mul(Tg[j * G + part], d[j]));      // T[part][j] (zero below the diagonal)
        T tv[G];
#pragma unroll
        for (int i = 0; i < G; ++i) tv[i] = t_shfl(tp, lbase + i);
        // z -= V t
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const T* vrow = vbase + i * Q2_VROW - i;
#pragma unroll
          for (int t = 0; t < 9; ++t) z[kb + t] = t_sub(z[kb + t], t_mul(vrow[8 * t], tv[i]));
        }
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
      cp_async_wait<0>();
      __syncthreads();
    }
  }
}

}  // namespace cb200
