// TEST INFRASTRUCTURE ONLY -- never linked into libchain_b200.so.
// Serial CPU implementation of chain_b200/csrc/backend.h so that the engine's HOST logic (block bookkeeping,
// task-list generation, Davidson/truncation sequencing, sweep driver) can be exercised by `pytest -m "not gpu"`
// in the GPU-less container.  Each function executes the same task lists the CUDA kernels execute, with the
// simplest possible loops; it makes no attempt to mirror the kernels' arithmetic order.
#include <cmath>
#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include "../../chain_b200/csrc/backend.h"

namespace cb200 {

using cd = std::complex<double>;

constexpr int64_t SYM_FLAG_BYTES = 256;    // arrival flags of the barrier (MAX_PEERS x u64), in front of the two landing halves

struct Backend {
  uint64_t launches = 0;
  double ev[8] = {0};
  // stand-in for the CUDA-IPC landing areas: one POSIX shared-memory segment per rank, [flags | landing 0 | landing 1], so that the
  // engine's exchange logic (fused all-gather, broadcasts, slice-wise end-to-end call) runs between real processes on the CPU
  char* sym = nullptr;
  size_t sym_bytes = 0;
  int64_t landing_each = 0;
  char shm_name[SHARD_HANDLE_BYTES] = {0};
  char* peer_sym[MAX_PEERS] = {nullptr};
  size_t peer_bytes[MAX_PEERS] = {0};
  int rank = 0, world = 1;
  unsigned long long barrier_seq = 0;
  std::string err;
  bool has_err = false;
};

Backend* backend_create(int, char*, size_t) { return new Backend(); }
void backend_destroy(Backend* b) {
  for (int r = 0; r < MAX_PEERS; ++r)
    if (b->peer_sym[r] && b->peer_sym[r] != b->sym) munmap(b->peer_sym[r], b->peer_bytes[r]);
  if (b->sym) { munmap(b->sym, b->sym_bytes); shm_unlink(b->shm_name); }
  delete b;
}
const char* backend_name() { return "host-emulation (tests only)"; }
int backend_device(Backend*) { return -1; }
// CB200_EMU_POISON=1: every allocation is filled with 0xFF bytes (NaN as a double, -1 as an integer) -- cudaMallocAsync hands out recycled,
// non-zero memory, while a fresh malloc() of a large block is zero pages: anything the engine reads before writing it shows up as NaN here.
void* dev_alloc(Backend*, size_t bytes) {
  static const bool poison = [] { const char* e = getenv("CB200_EMU_POISON"); return e && atoi(e) != 0; }();
  void* p = std::malloc(bytes ? bytes : 16);
  if (p && poison) std::memset(p, 0xFF, bytes ? bytes : 16);
  return p;
}
void dev_free(Backend*, void* p) { std::free(p); }
void dev_memset0(Backend*, void* p, size_t bytes) { std::memset(p, 0, bytes); }
void dev_h2d(Backend*, void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
void dev_d2h(Backend*, void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
void dev_d2d(Backend*, void* d, const void* s, size_t n) { std::memmove(d, s, n); }
void dev_sync(Backend*) {}
const char* dev_last_error(Backend* b) { return b->has_err ? b->err.c_str() : nullptr; }
uint64_t dev_launch_count(Backend* b) { return b->launches; }
void dev_event_record(Backend*, int) {}
float dev_event_elapsed_ms(Backend*, int, int) { return 0.f; }
int gemm_tile_m(bool cplx) { return cplx ? 64 : 128; }
int gemm_tile_n(bool, int mode) { return mode == GEMM_TILE_WIDE ? 128 : (mode == GEMM_TILE_MID_3M ? 96 : 64); }
int mix_rows_per_block(bool, int) { return 128; }
int copy_elems_per_block() { return 2048; }

// Rounding audit (tools/rounding_audit.py): with CB200_EMU_ROUNDING=<seed> every GEMM result is perturbed before it is stored by what a
// DIFFERENT summation order would have changed: u * ulps * 2^-53 * (sum_k |a_k b_k|) / sqrt(K), u uniform in [-1, 1) per real component
// (a random walk of K roundings of terms of the average magnitude; for complex numbers both components get the perturbation of the whole
// |a||b| sum -- normwise, which is what the 3M complex kernels of the product library do to a small component beside a large one).
// CB200_EMU_ROUNDING_ULPS (default 1) scales it, to measure the margin of the tolerances in the shared test bodies against arithmetic that
// differs from this serial loop the way the GPU's tile order / k-splitting / 3M products do.
static uint64_t emu_rounding_seed() {
  const char* e = getenv("CB200_EMU_ROUNDING");
  return e && *e ? strtoull(e, nullptr, 10) : 0;
}
static double emu_rounding_ulps() {
  const char* e = getenv("CB200_EMU_ROUNDING_ULPS");
  return e && *e ? atof(e) : 1.0;
}
static inline double emu_u(uint64_t seed, uint64_t& ctr) {   // splitmix64 -> uniform in [-1, 1)
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ++ctr;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(int64_t)z * (1.0 / 9223372036854775808.0);
}
// (the stream is keyed by the element's offset in C, so the same call repeated gives the same bits -- the suite asserts bit-identity between
// some call forms, as the GPU kernels guarantee it)
static inline double emu_perturb(double v, double mag, uint64_t seed, uint64_t key) {
  uint64_t ctr = key * 2;
  return v + mag * emu_u(seed, ctr);
}
static inline cd emu_perturb(cd v, double mag, uint64_t seed, uint64_t key) {
  uint64_t ctr = key * 2;
  const double u0 = emu_u(seed, ctr), u1 = emu_u(seed, ctr);
  return cd(v.real() + mag * u0, v.imag() + mag * u1);
}

template <class T>
static void gemm_t(bool ta, bool tb, const T* A, const T* B, T* C, const GemmTask* tasks, int nt, const GemmSeg* segs) {
  const uint64_t rseed = emu_rounding_seed();
  const double rulps = rseed ? emu_rounding_ulps() : 0.0;
  for (int ti = 0; ti < nt; ++ti) {
    const GemmTask& t = tasks[ti];
    const bool conjA = t.flags & GEMM_CONJ_A, acc = t.flags & GEMM_ACCUM, conjOut = t.flags & GEMM_CONJ_OUT;
    if (bool(t.flags & GEMM_TRANS_A) != ta && t.seg_count) { std::fprintf(stderr, "emu: transA flag mismatch\n"); std::abort(); }
    std::vector<T> tmp((size_t)t.M * t.N, T(0));
    std::vector<double> mag(rseed ? (size_t)t.M * t.N : 0, 0.0);   // sum_k |a_k b_k| per output element (rounding audit only)
    int64_t ktot = 0;
    for (int s = 0; s < t.seg_count; ++s) ktot += segs[t.seg_begin + s].K;
    for (int s = 0; s < t.seg_count; ++s) {
      const GemmSeg& g = segs[t.seg_begin + s];
      for (int n = 0; n < t.N; ++n)
        for (int k = 0; k < g.K; ++k) {
          const T b = tb ? B[g.b_off + n + (int64_t)k * g.ldb] : B[g.b_off + k + (int64_t)n * g.ldb];
          if (b == T(0)) continue;
          for (int m = 0; m < t.M; ++m) {
            T a = ta ? A[g.a_off + k + (int64_t)m * g.lda] : A[g.a_off + m + (int64_t)k * g.lda];
            if constexpr (std::is_same<T, cd>::value) { if (conjA) a = std::conj(a); }
            tmp[(size_t)m + (size_t)n * t.M] += a * b;
            if (rseed) mag[(size_t)m + (size_t)n * t.M] += std::abs(a) * std::abs(b);
          }
        }
    }
    for (int n = 0; n < t.N; ++n)
      for (int m = 0; m < t.M; ++m) {
        T v = tmp[(size_t)m + (size_t)n * t.M];
        if constexpr (std::is_same<T, cd>::value) { if (conjOut) v = std::conj(v); }
        const int64_t mc = (int64_t)(m % t.chunk) + (int64_t)(m / t.chunk) * t.chunk_stride;
        const int64_t nc = (int64_t)(n % t.chunk) + (int64_t)(n / t.chunk) * t.chunk_stride;
        T* p = (t.flags & GEMM_TRANS_C) ? &C[t.c_off + nc + (int64_t)m * t.ldc]
               : (n < t.n_split) ? &C[t.c_off + mc + (int64_t)n * t.ldc] : &C[t.c_off2 + mc + (int64_t)(n - t.n_split) * t.ldc];
        if (rseed) v = emu_perturb(v, rulps * 0x1p-53 * mag[(size_t)m + (size_t)n * t.M] / std::sqrt((double)std::max<int64_t>(ktot, 1)), rseed, (uint64_t)(p - C));
        *p = acc ? *p + v : v;
      }
  }
}

void launch_gemm(Backend* b, bool cplx, bool ta, bool tb, int, const void* A, const void* B, void* C, const GemmTask* t, int nt,
                 const GemmSeg* s, int tiles) {
  if (tiles <= 0 || nt <= 0) return;
  b->launches++;
  if (cplx) gemm_t<cd>(ta, tb, (const cd*)A, (const cd*)B, (cd*)C, t, nt, s);
  else gemm_t<double>(ta, tb, (const double*)A, (const double*)B, (double*)C, t, nt, s);
}

template <class T>
static void mix_t(const T* in, T* out, const MixPanel* panels, int np, const MixIn* ins, const MixOut* outs, const MixEntry* ents) {
  for (int pi = 0; pi < np; ++pi) {
    const MixPanel& p = panels[pi];
    for (int64_t r = 0; r < p.nr; ++r)
      for (int o = p.out_begin; o < p.out_end; ++o) {
        const MixOut& mo = outs[o];
        for (int row = 0; row < p.rows; ++row) {
          T acc(0);
          for (int e = mo.e_begin; e < mo.e_end; ++e) {
            const MixEntry& me = ents[e];
            T c;
            if constexpr (std::is_same<T, cd>::value) c = cd(me.cre, me.cim); else c = me.cre;
            const MixIn& mi = ins[p.in_begin + me.slot];
            acc += c * in[mi.off + r * mi.rstride + row];
          }
          out[mo.off + r * mo.rstride + row] = acc;
        }
      }
  }
}

void launch_mix(Backend* b, bool cplx, const void* in, void* out, const MixPanel* p, int np, const MixIn* mi, const MixOut* o, const MixEntry* e, int blocks, int) {
  if (blocks <= 0) return;
  b->launches++;
  if (cplx) mix_t<cd>((const cd*)in, (cd*)out, p, np, mi, o, e); else mix_t<double>((const double*)in, (double*)out, p, np, mi, o, e);
}

template <class T>
static void copy_t(const T* src, T* dst, const CopyTask* tasks, int nt) {
  for (int ti = 0; ti < nt; ++ti) {
    const CopyTask& t = tasks[ti];
    for (int64_t k = 0; k < t.n2; ++k)
      for (int64_t j = 0; j < t.n1; ++j)
        for (int64_t i = 0; i < t.n0; ++i) {
          T v = src[t.src_off + i * t.ss0 + j * t.ss1 + k * t.ss2];
          if constexpr (std::is_same<T, cd>::value) { if (t.conj) v = std::conj(v); }
          dst[t.dst_off + i * t.ds0 + j * t.ds1 + k * t.ds2] = v;
        }
  }
}

void launch_copy(Backend* b, bool cplx, const void* src, void* dst, const CopyTask* t, int nt, int blocks) {
  if (blocks <= 0) return;
  b->launches++;
  if (cplx) copy_t<cd>((const cd*)src, (cd*)dst, t, nt); else copy_t<double>((const double*)src, (double*)dst, t, nt);
}

void multi_dot(Backend* b, bool cplx, int nv, const void* const* x, const void* y, int64_t n, double* out) {
  b->launches += 2;
  for (int j = 0; j < nv; ++j) {
    if (cplx) {
      cd s(0);
      for (int64_t i = 0; i < n; ++i) s += std::conj(((const cd*)x[j])[i]) * ((const cd*)y)[i];
      out[2 * j] = s.real(); out[2 * j + 1] = s.imag();
    } else {
      double s = 0;
      for (int64_t i = 0; i < n; ++i) s += ((const double*)x[j])[i] * ((const double*)y)[i];
      out[2 * j] = s; out[2 * j + 1] = 0;
    }
  }
}

void lincomb(Backend* b, bool cplx, int nv, const void* const* x, const double* c, double br, double bi, void* dst, int64_t n) {
  b->launches++;
  for (int64_t i = 0; i < n; ++i) {
    if (cplx) {
      cd acc = (br != 0 || bi != 0) ? cd(br, bi) * ((cd*)dst)[i] : cd(0);
      for (int j = 0; j < nv; ++j) acc += cd(c[2 * j], c[2 * j + 1]) * ((const cd*)x[j])[i];
      ((cd*)dst)[i] = acc;
    } else {
      double acc = (br != 0) ? br * ((double*)dst)[i] : 0.0;
      for (int j = 0; j < nv; ++j) acc += c[2 * j] * ((const double*)x[j])[i];
      ((double*)dst)[i] = acc;
    }
  }
}

// serial cyclic Jacobi on one NJ x NJ Hermitian matrix
static long long jacobi_one(int n, std::vector<cd>& G, std::vector<cd>& Q, double tol, double floor2, double abs_g2, int max_sweeps) {
  long long total = 0;
  if (max_sweeps <= 0 || max_sweeps > 60) max_sweeps = 60;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    long long nrot = 0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const cd g = G[p + (size_t)q * n];
        const double g2 = std::norm(g), gpp = G[p + (size_t)p * n].real(), gqq = G[q + (size_t)q * n].real();
        if (!(g2 > tol * tol * std::fabs(gpp * gqq)) || !(g2 > abs_g2) || !(gpp > floor2) || !(gqq > floor2)) continue;
        ++nrot;
        const double ag = std::sqrt(g2);
        const cd e = g / ag;
        const double tau = (gqq - gpp) / (2.0 * ag);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int r = 0; r < n; ++r) {
          cd a = G[r + (size_t)p * n], b = G[r + (size_t)q * n];
          G[r + (size_t)p * n] = c * a - s * std::conj(e) * b;
          G[r + (size_t)q * n] = s * e * a + c * b;
          a = Q[r + (size_t)p * n]; b = Q[r + (size_t)q * n];
          Q[r + (size_t)p * n] = c * a - s * std::conj(e) * b;
          Q[r + (size_t)q * n] = s * e * a + c * b;
        }
        for (int cc = 0; cc < n; ++cc) {
          const cd a = G[p + (size_t)cc * n], b = G[q + (size_t)cc * n];
          G[p + (size_t)cc * n] = c * a - s * e * b;
          G[q + (size_t)cc * n] = s * std::conj(e) * a + c * b;
        }
        G[p + (size_t)q * n] = 0; G[q + (size_t)p * n] = 0;
      }
    total += nrot;
    if (!nrot) break;
  }
  return total;
}

void launch_jacobi_eig(Backend* b, bool cplx, const void* G_, void* Q_, int npairs, double tol, double floor2, double abs_g2,
                       int max_sweeps, long long* d_rot) {
  b->launches++;
  const int n = 2 * JB;
  for (int pi = 0; pi < npairs; ++pi) {
    std::vector<cd> G((size_t)n * n), Q((size_t)n * n, cd(0));
    for (int c = 0; c < n; ++c)
      for (int r = 0; r < n; ++r) {
        const int rr = r <= c ? r : c, cc = r <= c ? c : r;
        cd v = cplx ? ((const cd*)G_)[(size_t)pi * n * n + rr + (size_t)cc * n] : cd(((const double*)G_)[(size_t)pi * n * n + rr + (size_t)cc * n], 0);
        if (r > c) v = std::conj(v);
        if (r == c) v = cd(v.real(), 0);
        G[r + (size_t)c * n] = v;
      }
    for (int i = 0; i < n; ++i) Q[i + (size_t)i * n] = 1;
    *d_rot += jacobi_one(n, G, Q, tol, floor2, abs_g2, max_sweeps);
    for (size_t i = 0; i < (size_t)n * n; ++i) {
      if (cplx) ((cd*)Q_)[(size_t)pi * n * n + i] = Q[i]; else ((double*)Q_)[(size_t)pi * n * n + i] = Q[i].real();
    }
  }
}

void launch_colnorm2(Backend* b, bool cplx, const void* X, int64_t ld, int64_t rows, int64_t cols, double* out) {
  b->launches++;
  for (int64_t c = 0; c < cols; ++c) {
    double s = 0;
    for (int64_t i = 0; i < rows; ++i) s += cplx ? std::norm(((const cd*)X)[i + c * ld]) : ((const double*)X)[i + c * ld] * ((const double*)X)[i + c * ld];
    out[c] = s;
  }
}

void launch_set_identity(Backend* b, bool cplx, void* X, int64_t ld, int64_t n) {
  b->launches++;
  for (int64_t c = 0; c < n; ++c)
    for (int64_t r = 0; r < n; ++r) {
      if (cplx) ((cd*)X)[r + c * ld] = cd(r == c ? 1.0 : 0.0, 0); else ((double*)X)[r + c * ld] = (r == c ? 1.0 : 0.0);
    }
}

double fp64_pipe_peak(Backend*, int, int) { return 0.0; }

// ---- Hermitian eigensolver building blocks: serial restatement of the device kernels (same conventions, see backend.h) ----
template <class T> static inline T cj(const T& x) { return x; }
template <> inline cd cj<cd>(const cd& x) { return std::conj(x); }
template <class T> static inline double re_(const T& x) { return x; }
template <> inline double re_<cd>(const cd& x) { return x.real(); }
template <class T> static inline double im_(const T&) { return 0.0; }
template <> inline double im_<cd>(const cd& x) { return x.imag(); }
template <class T> static inline T mk(double r, double i) { return r; }
template <> inline cd mk<cd>(double r, double i) { return cd(r, i); }

template <class T>
static void tridiag_t(T* A, int64_t n, double* d, double* e, T* tau, T* scl, T* p) {
  for (int64_t j = 0; j + 1 < n; ++j) {
    const int64_t m = n - j - 1;
    T* x = A + (j + 1) + j * n;
    const T alpha = x[0];
    double sigma = 0;
    for (int64_t i = 1; i < m; ++i) sigma += std::norm(cd(re_(x[i]), im_(x[i])));
    const double ar = re_(alpha), ai = im_(alpha);
    d[j] = re_(A[j + j * n]);
    if (sigma == 0.0 && ai == 0.0) { tau[j] = T(0); scl[j] = T(0); e[j] = ar; continue; }
    const double beta = -std::copysign(std::sqrt(ar * ar + ai * ai + sigma), ar);
    const T tj = mk<T>((beta - ar) / beta, -ai / beta);
    const T sc = T(1) / (alpha - T(beta));
    auto v = [&](int64_t i) { return i == 0 ? T(1) : x[i] * sc; };
    T* A22 = A + (j + 1) + (j + 1) * n;
    for (int64_t c = 0; c < m; ++c) {
      T acc(0);
      for (int64_t r = 0; r < m; ++r) acc += cj(A22[r + c * n]) * v(r);
      p[c] = tj * acc;
    }
    T s(0);
    for (int64_t i = 0; i < m; ++i) s += cj(p[i]) * v(i);
    const T a2 = T(-0.5) * tj * s;
    for (int64_t i = 0; i < m; ++i) p[i] += a2 * v(i);     // w
    for (int64_t c = 0; c < m; ++c)
      for (int64_t r = 0; r < m; ++r) A22[r + c * n] -= v(r) * cj(p[c]) + p[r] * cj(v(c));
    e[j] = beta; tau[j] = tj; scl[j] = sc;
  }
  d[n - 1] = re_(A[(n - 1) + (n - 1) * n]);
}

bool launch_tridiag(Backend* b, bool cplx, void* A, int64_t n, double* d, double* e, void* tau, void* scl, void* p) {
  b->launches++;
  if (cplx) tridiag_t<cd>((cd*)A, n, d, e, (cd*)tau, (cd*)scl, (cd*)p); else tridiag_t<double>((double*)A, n, d, e, (double*)tau, (double*)scl, (double*)p);
  return true;
}

bool launch_tridiag_batched(Backend* b, bool cplx, const TriProblem* probs, int nprob) {
  for (int k = 0; k < nprob; ++k)
    if (!launch_tridiag(b, cplx, probs[k].A, probs[k].n, probs[k].d, probs[k].e, probs[k].tau, probs[k].scl, probs[k].p_work)) return false;
  return true;
}

void launch_bisect(Backend* b, const double* d, const double* e, int64_t n, double* w) {
  b->launches++;
  double gl = d[0], gu = d[0], emax2 = 0, tn = 0;
  for (int64_t i = 0; i < n; ++i) {
    const double r = (i > 0 ? std::fabs(e[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(e[i]) : 0.0);
    tn = std::max(tn, std::fabs(d[i]) + r);
    gl = std::min(gl, d[i] - r); gu = std::max(gu, d[i] + r);
    if (i + 1 < n) emax2 = std::max(emax2, e[i] * e[i]);
  }
  const double pivmin = 2.2250738585072014e-308 * std::max(1.0, emax2);
  const double span = std::max(gu - gl, 1e-300);
  gl -= 2.0 * span * 2.3e-16 * n + 2 * pivmin; gu += 2.0 * span * 2.3e-16 * n + 2 * pivmin;
  for (int64_t k = 0; k < n; ++k) {
    double lo = gl, hi = gu;
    for (int it = 0; it < 200; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (mid <= lo || mid >= hi || hi - lo <= 5.6e-17 * tn) break;
      int64_t cnt = 0;
      double q = d[0] - mid;
      if (std::fabs(q) < pivmin) q = -pivmin;
      if (q < 0) ++cnt;
      for (int64_t i = 1; i < n; ++i) {
        q = d[i] - mid - e[i - 1] * e[i - 1] / q;
        if (std::fabs(q) < pivmin) q = -pivmin;
        if (q < 0) ++cnt;
      }
      if (cnt > k) hi = mid; else lo = mid;
    }
    w[k] = 0.5 * (lo + hi);
  }
}

static inline double hash_unit(uint64_t i, uint64_t j) {   // deterministic start vector in (-1, 1)
  uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull ^ (j + 1) * 0xC2B2AE3D27D4EB4Full;
  z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 29; z *= 0x94D049BB133111EBull; z ^= z >> 32;
  return ((double)(z >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}

void launch_invit(Backend* b, const double* d, const double* e, int64_t n, const double* lam, int64_t m, double* Zt, double* work) {
  b->launches++;
  double tnorm = 0;
  for (int64_t i = 0; i < n; ++i) tnorm = std::max(tnorm, std::fabs(d[i]) + (i > 0 ? std::fabs(e[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(e[i]) : 0.0));
  const double tol = std::max(2.3e-16 * tnorm, 1e-300);
  double* u0 = work; double* u1 = work + n * m; double* u2 = work + 2 * n * m; double* lm = work + 3 * n * m; double* sw = work + 4 * n * m;
  for (int64_t j = 0; j < m; ++j) {
    auto at = [&](double* a, int64_t i) -> double& { return a[i * m + j]; };
    const double l = lam[j];
    double p0 = d[0] - l, p1 = n > 1 ? e[0] : 0.0;
    for (int64_t i = 0; i + 1 < n; ++i) {
      const double r0 = e[i], r1 = d[i + 1] - l, r2 = (i + 2 < n) ? e[i + 1] : 0.0;
      if (std::fabs(p0) >= std::fabs(r0)) {
        if (std::fabs(p0) < tol) p0 = p0 < 0 ? -tol : tol;
        const double mult = r0 / p0;
        at(sw, i) = 0; at(lm, i) = mult; at(u0, i) = p0; at(u1, i) = p1; at(u2, i) = 0;
        p0 = r1 - mult * p1; p1 = r2;
      } else {
        const double mult = p0 / r0;
        at(sw, i) = 1; at(lm, i) = mult; at(u0, i) = r0; at(u1, i) = r1; at(u2, i) = r2;
        p0 = p1 - mult * r1; p1 = -mult * r2;
      }
    }
    if (std::fabs(p0) < tol) p0 = p0 < 0 ? -tol : tol;
    at(u0, n - 1) = p0; at(u1, n - 1) = 0; at(u2, n - 1) = 0;
    for (int64_t i = 0; i < n; ++i) at(Zt, i) = hash_unit(i, j);
    for (int it = 0; it < 3; ++it) {
      for (int64_t i = 0; i + 1 < n; ++i) {
        if (at(sw, i) != 0) std::swap(at(Zt, i), at(Zt, i + 1));
        at(Zt, i + 1) -= at(lm, i) * at(Zt, i);
      }
      double mx = 0;
      for (int64_t i = n - 1; i >= 0; --i) {
        double y = at(Zt, i);
        if (i + 1 < n) y -= at(u1, i) * at(Zt, i + 1);
        if (i + 2 < n) y -= at(u2, i) * at(Zt, i + 2);
        y /= at(u0, i);
        at(Zt, i) = y;
        mx = std::max(mx, std::fabs(y));
      }
      const double sc = mx > 0 ? 1.0 / mx : 1.0;
      for (int64_t i = 0; i < n; ++i) at(Zt, i) *= sc;
    }
    double nn = 0;
    for (int64_t i = 0; i < n; ++i) nn += at(Zt, i) * at(Zt, i);
    nn = 1.0 / std::sqrt(nn);
    for (int64_t i = 0; i < n; ++i) at(Zt, i) *= nn;
  }
}

void launch_ns_coeff(Backend* b, const double* G, int64_t m, double alpha, double* Cout, double* stats) {
  b->launches++;
  double dev = 0, rs = 0;
  for (int64_t r = 0; r < m; ++r) {
    double s = 0;
    for (int64_t c = 0; c < m; ++c) {
      const double g = G[r + c * m];
      dev = std::max(dev, std::fabs(g - (r == c ? 1.0 : 0.0)));
      s += std::fabs(g);
      Cout[r + c * m] = alpha * ((r == c ? 1.5 : 0.0) - 0.5 * alpha * alpha * g);
    }
    rs = std::max(rs, s);
  }
  stats[0] = dev; stats[1] = rs;
}

template <class T>
static void backtransform_t(const T* A, int64_t n, const T* tau, const T* scl, const double* Zt, int64_t m, T* V) {
  for (int64_t c = 0; c < m; ++c) {
    T* z = V + c * n;
    for (int64_t i = 0; i < n; ++i) z[i] = T(Zt[i * m + c]);
    for (int64_t j = n - 2; j >= 0; --j) {
      if (tau[j] == T(0)) continue;
      const T* x = A + (j + 1) + j * n;
      const int64_t len = n - j - 1;
      auto v = [&](int64_t i) { return i == 0 ? T(1) : x[i] * scl[j]; };
      T t(0);
      for (int64_t i = 0; i < len; ++i) t += cj(v(i)) * z[j + 1 + i];
      t *= tau[j];
      for (int64_t i = 0; i < len; ++i) z[j + 1 + i] -= t * v(i);
    }
  }
}

template <class T>
static void zt_to_v_t(const double* Zt, int64_t n, int64_t m, T* V) {
  for (int64_t c = 0; c < m; ++c)
    for (int64_t i = 0; i < n; ++i) V[i + c * n] = T(Zt[i * m + c]);
}
void launch_zt_to_v(Backend* b, bool cplx, const double* Zt, int64_t n, int64_t m, void* V) {
  b->launches++;
  if (cplx) zt_to_v_t<cd>(Zt, n, m, (cd*)V); else zt_to_v_t<double>(Zt, n, m, (double*)V);
}
template <class T>
static void form_y_t(const T* A, int64_t n, const T* tau, const T* scl, int64_t j0, int nbk, T* Y) {
  const int64_t r0 = j0 + 1, len = n - r0;
  for (int i = 0; i < nbk; ++i) {
    const int64_t j = j0 + i;
    for (int64_t rr = 0; rr < len; ++rr) {
      const int64_t r = r0 + rr;
      T y(0);
      if (!(tau[j] == T(0))) {
        if (r == j + 1) y = T(1);
        else if (r > j + 1) y = A[r + j * n] * scl[j];
      }
      Y[rr + (int64_t)i * len] = y;
    }
  }
}
void launch_form_y(Backend* b, bool cplx, const void* A, int64_t n, const void* tau, const void* scl, int64_t j0, int nbk, void* Y) {
  b->launches++;
  if (cplx) form_y_t<cd>((const cd*)A, n, (const cd*)tau, (const cd*)scl, j0, nbk, (cd*)Y);
  else form_y_t<double>((const double*)A, n, (const double*)tau, (const double*)scl, j0, nbk, (double*)Y);
}
template <class T>
static void wy_t_t(const T* S, const T* tau, int nb, T* Tn) {
  std::vector<T> Ts((size_t)nb * nb, T(0));
  for (int i = 0; i < nb; ++i) {
    for (int r = 0; r < i; ++r) {
      T tmp(0);
      for (int c = r; c < i; ++c) tmp += Ts[r + (size_t)c * nb] * S[c + (size_t)i * nb];
      Ts[r + (size_t)i * nb] = -tau[i] * tmp;
    }
    Ts[i + (size_t)i * nb] = tau[i];
  }
  for (size_t k = 0; k < Ts.size(); ++k) Tn[k] = -Ts[k];
}
void launch_wy_t(Backend* b, bool cplx, const void* S, const void* tau_j0, int nbk, void* Tneg) {
  b->launches++;
  if (cplx) wy_t_t<cd>((const cd*)S, (const cd*)tau_j0, nbk, (cd*)Tneg); else wy_t_t<double>((const double*)S, (const double*)tau_j0, nbk, (double*)Tneg);
}

void launch_wy_t_batched(Backend* b, bool cplx, const void* S, const void* tau, int npanel, int nb_last, void* Tneg) {
  const size_t es = cplx ? sizeof(cd) : sizeof(double);
  for (int k = 0; k < npanel; ++k)
    launch_wy_t(b, cplx, (const char*)S + es * 4096 * (size_t)k, (const char*)tau + es * 64 * (size_t)k, k == npanel - 1 ? nb_last : 64,
                (char*)Tneg + es * 4096 * (size_t)k);
}

// ---- two-stage reduction (serial restatement of the device algorithm in heig2.cuh: same formulas, same storage conventions)
template <class T>
static void larfg_t(int64_t len, T* x, T* tau_out, double* beta_out) {      // x[0] = alpha; on exit x = v with v[0] = 1
  const T alpha = x[0];
  double sigma = 0;
  for (int64_t i = 1; i < len; ++i) sigma += re_(x[i]) * re_(x[i]) + im_(x[i]) * im_(x[i]);
  const double ar = re_(alpha), ai = im_(alpha);
  if (sigma == 0.0 && ai == 0.0) { *tau_out = T(0); *beta_out = ar; x[0] = T(1); for (int64_t i = 1; i < len; ++i) x[i] = T(0); return; }
  const double beta = -std::copysign(std::sqrt(ar * ar + ai * ai + sigma), ar);
  *tau_out = mk<T>((beta - ar) / beta, -ai / beta);
  const T sc = T(1) / (alpha - T(beta));
  for (int64_t i = 1; i < len; ++i) x[i] *= sc;
  x[0] = T(1);
  *beta_out = beta;
}

template <class T>
static void panel_qr_t(T* A, int64_t n, int64_t j0, T* Y, T* Tn, T* Tnh) {
  const int B = SB_B;
  const int64_t r0 = j0 + B, len = n - r0;
  const int kp = (int)std::min<int64_t>(B, len - 1);
  std::vector<T> P((size_t)len * B), Tm((size_t)B * B, T(0)), tau(B, T(0));
  for (int c = 0; c < B; ++c) for (int64_t r = 0; r < len; ++r) P[r + (size_t)c * len] = A[(r0 + r) + (j0 + c) * n];
  for (int c = 0; c < kp; ++c) {
    T* x = &P[c + (size_t)c * len];
    double beta;
    larfg_t<T>(len - c, x, &tau[c], &beta);
    // apply H_c^H = I - conj(tau) v v^H to the columns to the right
    for (int j = c + 1; j < B; ++j) {
      T* pj = &P[c + (size_t)j * len];
      T sdot(0);
      for (int64_t r = 0; r < len - c; ++r) sdot += cj(x[r]) * pj[r];
      sdot *= cj(tau[c]);
      for (int64_t r = 0; r < len - c; ++r) pj[r] -= sdot * x[r];
    }
    // T(0:c, c) = -tau_c T(0:c,0:c) (Y(:,0:c)^H v_c),  T(c,c) = tau_c     (LAPACK larft, forward / columnwise)
    std::vector<T> z(c, T(0));
    for (int k = 0; k < c; ++k) {
      const T* yk = &P[(size_t)k * len];     // column k holds v_k below its diagonal (explicit 1 at row k)
      T acc(0);
      for (int64_t r = c; r < len; ++r) acc += cj(yk[r]) * P[r + (size_t)c * len];
      z[k] = acc;
    }
    for (int r = 0; r < c; ++r) {
      T acc(0);
      for (int k = r; k < c; ++k) acc += Tm[r + (size_t)k * B] * z[k];
      Tm[r + (size_t)c * B] = -tau[c] * acc;
    }
    Tm[c + (size_t)c * B] = tau[c];
    // keep beta aside: the diagonal slot currently holds the explicit 1 of v_c
    A[(r0 + c) + (j0 + c) * n] = T(beta);
  }
  // outputs: Y explicit, R into A (rows <= columns; diagonal of reflector columns already written), zeros below
  for (int c = 0; c < B; ++c)
    for (int64_t r = 0; r < len; ++r) {
      T y(0);
      if (c < kp && r >= c) y = P[r + (size_t)c * len];
      Y[r + (size_t)c * len] = y;
      T a;
      if (r < c) a = P[r + (size_t)c * len];
      else if (r == c) a = (c < kp) ? A[(r0 + r) + (j0 + c) * n] : P[r + (size_t)c * len];
      else a = (c < kp) ? T(0) : P[r + (size_t)c * len];
      A[(r0 + r) + (j0 + c) * n] = a;
    }
  for (size_t k = 0; k < Tm.size(); ++k) { Tn[k] = -Tm[k]; Tnh[k] = T(-0.5) * Tm[k]; }
}
void launch_panel_qr(Backend* b, bool cplx, void* A, int64_t n, int64_t j0, void* Y, void* Tn, void* Tnh) {
  b->launches++;
  if (cplx) panel_qr_t<cd>((cd*)A, n, j0, (cd*)Y, (cd*)Tn, (cd*)Tnh); else panel_qr_t<double>((double*)A, n, j0, (double*)Y, (double*)Tn, (double*)Tnh);
}

template <class T>
static void band_extract_t(const T* A, int64_t n, T* AB) {
  for (int64_t j = 0; j < n; ++j)
    for (int dd = 0; dd < SB_LD; ++dd) {
      const int64_t i = j + dd;
      AB[dd + j * SB_LD] = (dd <= SB_B && i < n) ? A[i + j * n] : T(0);
    }
}
void launch_band_extract(Backend* b, bool cplx, const void* A, int64_t n, void* AB) {
  b->launches++;
  if (cplx) band_extract_t<cd>((const cd*)A, n, (cd*)AB); else band_extract_t<double>((const double*)A, n, (double*)AB);
}

int64_t chase_blocks(int64_t n) {
  const int64_t K0 = (n - 1 + SB_B - 1) / SB_B;
  return std::max<int64_t>(K0 * (K0 + 1) / 2, 1);
}

template <class T>
static void bulge_chase_t(T* AB, int64_t n, double* d, double* e, T* V2, T* tau2) {
  const int B = SB_B;
  const int64_t K0 = (n - 1 + B - 1) / B;
  auto el = [&](int64_t i, int64_t j) -> T& { return AB[(i - j) + j * SB_LD]; };     // i >= j, i - j < SB_LD
  std::vector<T> D((size_t)B * B), Bk((size_t)B * B), v(B), vn(B), w(B), z(B), y(B);
  for (int64_t s = 0; s + 1 < n; ++s) {
    const int64_t g = s / B, sp = s % B;
    T tau(0);
    for (int64_t k = 0;; ++k) {
      const int64_t c0 = s + 1 + k * B;
      if (c0 > n - 1) break;
      const int nd = (int)std::min<int64_t>(B, n - c0);
      const int nb2 = (int)std::max<int64_t>(0, std::min<int64_t>(B, n - c0 - B));
      if (k == 0) {
        d[s] = re_(el(s, s));
        for (int i = 0; i < nd; ++i) v[i] = el(c0 + i, s);
        double beta;
        larfg_t<T>(nd, v.data(), &tau, &beta);
        e[s] = beta;
        el(c0, s) = T(beta);
        for (int i = 1; i < nd; ++i) el(c0 + i, s) = T(0);
      }
      // store the reflector for the back-transformation
      {
        const int64_t blk = g * K0 - g * (g - 1) / 2 + k;
        T* dst = V2 + blk * (int64_t)SB_VLD * B + sp * (int64_t)SB_VLD + sp;
        for (int i = 0; i < nd; ++i) dst[i] = v[i];
        tau2[blk * B + sp] = tau;
      }
      // D <- H^H D H  (Hermitian nd x nd, lower triangle stored)
      for (int j = 0; j < nd; ++j) for (int i = 0; i < nd; ++i) D[i + (size_t)j * B] = (i >= j) ? el(c0 + i, c0 + j) : cj(el(c0 + j, c0 + i));
      for (int i = 0; i < nd; ++i) { T acc(0); for (int j = 0; j < nd; ++j) acc += D[i + (size_t)j * B] * v[j]; w[i] = tau * acc; }
      T vp(0);
      for (int i = 0; i < nd; ++i) vp += cj(v[i]) * w[i];
      const T a2 = T(-0.5) * cj(tau) * vp;
      for (int i = 0; i < nd; ++i) w[i] += a2 * v[i];
      for (int j = 0; j < nd; ++j) for (int i = j; i < nd; ++i) el(c0 + i, c0 + j) = D[i + (size_t)j * B] - v[i] * cj(w[j]) - w[i] * cj(v[j]);
      if (nb2 <= 0) {
        if (c0 + nd == n && nd == 1) { /* 1 x 1 trailing block: unchanged by a unitary 1 x 1 similarity */ }
        break;
      }
      // B <- B H (right), then the next reflector from its first column, then B <- H'^H B (left)
      for (int j = 0; j < nd; ++j) for (int i = 0; i < nb2; ++i) Bk[i + (size_t)j * B] = el(c0 + B + i, c0 + j);
      for (int i = 0; i < nb2; ++i) { T acc(0); for (int j = 0; j < nd; ++j) acc += Bk[i + (size_t)j * B] * v[j]; z[i] = tau * acc; }
      for (int j = 0; j < nd; ++j) for (int i = 0; i < nb2; ++i) Bk[i + (size_t)j * B] -= z[i] * cj(v[j]);
      for (int i = 0; i < nb2; ++i) vn[i] = Bk[i];
      T taun;
      double betan;
      larfg_t<T>(nb2, vn.data(), &taun, &betan);
      for (int j = 1; j < nd; ++j) { T acc(0); for (int i = 0; i < nb2; ++i) acc += cj(vn[i]) * Bk[i + (size_t)j * B]; y[j] = cj(taun) * acc; }
      for (int j = 1; j < nd; ++j) for (int i = 0; i < nb2; ++i) Bk[i + (size_t)j * B] -= vn[i] * y[j];
      Bk[0] = T(betan);
      for (int i = 1; i < nb2; ++i) Bk[i] = T(0);
      for (int j = 0; j < nd; ++j) for (int i = 0; i < nb2; ++i) el(c0 + B + i, c0 + j) = Bk[i + (size_t)j * B];
      for (int i = 0; i < nb2; ++i) v[i] = vn[i];
      for (int i = nb2; i < B; ++i) v[i] = T(0);
      tau = taun;
    }
  }
  d[n - 1] = re_(el(n - 1, n - 1));
}
void launch_bulge_chase(Backend* b, bool cplx, const ChaseProblem* probs, int nprob) {
  b->launches++;
  for (int p = 0; p < nprob; ++p) {
    const ChaseProblem& q = probs[p];
    if (cplx) bulge_chase_t<cd>((cd*)q.AB, q.n, q.d, q.e, (cd*)q.V2, (cd*)q.tau2); else bulge_chase_t<double>((double*)q.AB, q.n, q.d, q.e, (double*)q.V2, (double*)q.tau2);
  }
}

template <class T>
static void q2_apply_t(const T* V2, const T* tau2, int64_t n, T* Z, int64_t m) {
  const int B = SB_B;
  const int64_t K0 = (n - 1 + B - 1) / B;
  for (int64_t g = K0 - 1; g >= 0; --g)
    for (int64_t k = 0; k < K0 - g; ++k) {
      const int64_t blk = g * K0 - g * (g - 1) / 2 + k;
      const int64_t base = g * B + 1 + k * B;
      for (int sp = B - 1; sp >= 0; --sp) {
        const T tau = tau2[blk * B + sp];
        if (tau == T(0)) continue;
        const T* v = V2 + blk * (int64_t)SB_VLD * B + sp * (int64_t)SB_VLD + sp;
        const int64_t r0 = base + sp;
        const int len = (int)std::min<int64_t>(B, n - r0);
        if (len <= 0) continue;
        for (int64_t c = 0; c < m; ++c) {
          T* zc = Z + r0 + c * n;
          T acc(0);
          for (int i = 0; i < len; ++i) acc += cj(v[i]) * zc[i];
          acc *= tau;
          for (int i = 0; i < len; ++i) zc[i] -= acc * v[i];
        }
      }
    }
}
void launch_q2_apply(Backend* b, bool cplx, const void* V2, const void* tau2, int64_t n, void* Z, int64_t m) {
  b->launches++;
  if (cplx) q2_apply_t<cd>((const cd*)V2, (const cd*)tau2, n, (cd*)Z, m); else q2_apply_t<double>((const double*)V2, (const double*)tau2, n, (double*)Z, m);
}

void launch_truncate_select(Backend* b, const double* pooled, int64_t n, int64_t maxdim, int64_t mindim, double cutoff, double* P,
                            double* out3) {
  b->launches++;
  if (n <= 0) return;
  std::vector<double> v(pooled, pooled + n);
  std::stable_sort(v.begin(), v.end(), [](double x, double y) { return x > y; });
  for (int64_t i = 0; i < n; ++i) P[i] = v[i];
  const int64_t origm = n;
  int64_t nn = origm - 1;
  for (int64_t zn = nn; zn >= 0 && P[zn] < 0; --zn) P[zn] = 0;
  if (origm == 1) { out3[0] = 1; out3[1] = 0; out3[2] = P[0] / 2.0; return; }
  double truncerr = 0;
  while (nn >= maxdim) { truncerr += P[nn]; --nn; }
  double scale = 0;
  for (int64_t k = 0; k < origm; ++k) scale += P[k];
  if (scale == 0) scale = 1.0;
  while (nn >= mindim && nn >= 0 && truncerr + P[nn] < cutoff * scale) { truncerr += P[nn]; --nn; }
  truncerr /= scale;
  if (nn < 0) nn = 0;
  const int64_t m = nn + 1;
  double docut = 0;
  if (m < origm) {
    docut = (P[m - 1] + P[m]) / 2.0;
    if (std::fabs(P[m - 1] - P[m]) < 1e-3 * P[m - 1]) docut += 1e-3 * P[m - 1];
  }
  out3[0] = (double)m; out3[1] = truncerr; out3[2] = docut;
}

void launch_backtransform(Backend* b, bool cplx, const void* A, int64_t n, const void* tau, const void* scl, const double* Zt, int64_t m, void* V) {
  b->launches++;
  if (cplx) backtransform_t<cd>((const cd*)A, n, (const cd*)tau, (const cd*)scl, Zt, m, (cd*)V);
  else backtransform_t<double>((const double*)A, n, (const double*)tau, (const double*)scl, Zt, m, (double*)V);
}

// ---- multi-process plumbing of the host emulation: the p2p exchange (mode 2) over POSIX shared memory, same protocol as the CUDA backend
// (per-rank segment [flags | landing 0 | landing 1], peers map each other's segments, arrival-counter barrier, broadcasts / all-gather as
// stores into every landing half followed by the barrier).  NCCL (mode 1) has no stand-in.
static void emu_err(Backend* b, const std::string& m) { if (!b->has_err) { b->err = m; b->has_err = true; } }
int shard_alloc(Backend* b, int64_t landing_each, void* handle_out) {
  if (b->sym) { emu_err(b, "shard_alloc: already allocated"); return -1; }
  static int counter = 0;
  landing_each = (landing_each + 255) / 256 * 256;
  std::snprintf(b->shm_name, sizeof(b->shm_name), "/cb200emu_%d_%d", (int)getpid(), counter++);
  const int fd = shm_open(b->shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) { emu_err(b, "shard_alloc: shm_open failed"); return -1; }
  b->sym_bytes = (size_t)(SYM_FLAG_BYTES + 2 * landing_each);
  if (ftruncate(fd, (off_t)b->sym_bytes) != 0) { close(fd); shm_unlink(b->shm_name); emu_err(b, "shard_alloc: ftruncate failed"); return -1; }
  void* p = mmap(nullptr, b->sym_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { shm_unlink(b->shm_name); emu_err(b, "shard_alloc: mmap failed"); return -1; }
  b->sym = (char*)p;
  b->landing_each = landing_each;
  std::memset(b->sym, 0, SYM_FLAG_BYTES);
  std::memcpy(handle_out, b->shm_name, SHARD_HANDLE_BYTES);
  return 0;
}
int shard_connect(Backend* b, int rank, int world, const void* all_handles) {
  if (!b->sym || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) { emu_err(b, "shard_connect: bad arguments or no landing area"); return -1; }
  b->rank = rank; b->world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { b->peer_sym[r] = b->sym; b->peer_bytes[r] = b->sym_bytes; continue; }
    char name[SHARD_HANDLE_BYTES + 1] = {0};
    std::memcpy(name, (const char*)all_handles + (size_t)r * SHARD_HANDLE_BYTES, SHARD_HANDLE_BYTES);
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) { emu_err(b, "shard_connect: shm_open of a peer segment failed"); return -1; }
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size != b->sym_bytes) { close(fd); emu_err(b, "shard_connect: peer landing area has a different size"); return -1; }
    void* p = mmap(nullptr, b->sym_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { emu_err(b, "shard_connect: mmap of a peer segment failed"); return -1; }
    b->peer_sym[r] = (char*)p; b->peer_bytes[r] = b->sym_bytes;
  }
  return 0;
}
int64_t shard_landing_bytes(Backend* b) { return b->sym ? b->landing_each : 0; }
void* shard_landing(Backend* b, int which) { return b->sym ? b->sym + SYM_FLAG_BYTES + (int64_t)(which & 1) * b->landing_each : nullptr; }
void shard_peer_bases(Backend* b, int which, PeerBases* out) {
  out->n = b->sym ? b->world : 0;
  for (int r = 0; r < MAX_PEERS; ++r)
    out->base[r] = (b->sym && r < b->world) ? b->peer_sym[r] + SYM_FLAG_BYTES + (int64_t)(which & 1) * b->landing_each : nullptr;
}
// rank publishes its arrival number in every peer's flag row (release) and waits for every peer's number in its own row (acquire)
void shard_barrier(Backend* b) {
  if (b->world <= 1 || !b->sym) return;
  const unsigned long long seq = ++b->barrier_seq;
  for (int t = 0; t < b->world; ++t)
    __atomic_store_n(reinterpret_cast<unsigned long long*>(b->peer_sym[t]) + b->rank, seq, __ATOMIC_RELEASE);
  struct timespec t0, now;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < b->world; ++t) {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(b->sym) + t;
    while (__atomic_load_n(src, __ATOMIC_ACQUIRE) < seq) {
      sched_yield();
      clock_gettime(CLOCK_MONOTONIC, &now);
      if (now.tv_sec - t0.tv_sec > 120) { emu_err(b, "shard_barrier: a peer did not arrive within 120 s"); return; }
    }
  }
  b->launches++;
}
void launch_gemm_peer(Backend* b, bool cplx, int, const void* A, const void* B, const PeerBases& peers, const GemmTask* t, int nt,
                      const GemmSeg* s, int tiles) {
  if (tiles <= 0 || nt <= 0) return;
  b->launches++;
  for (int r = 0; r < peers.n; ++r) {   // the R-stage product (A not transposed, B transposed) stored into every rank's landing half
    if (cplx) gemm_t<cd>(false, true, (const cd*)A, (const cd*)B, (cd*)peers.base[r], t, nt, s);
    else gemm_t<double>(false, true, (const double*)A, (const double*)B, (double*)peers.base[r], t, nt, s);
  }
}
int nccl_unique_id(Backend*, void*) { return -1; }
int nccl_init(Backend*, const void*, int, int) { return -1; }
void nccl_allreduce_sum(Backend*, void*, int64_t) {}
void nccl_bcast(Backend* b, void*, int64_t, int) { emu_err(b, "nccl_bcast: the host emulation has no NCCL stand-in"); }
void shard_bcast_p2p(Backend* b, void* buf, int64_t bytes, int root, int which) {
  if (b->world <= 1 || bytes <= 0) return;
  if (bytes > b->landing_each) { emu_err(b, "shard_bcast_p2p: message larger than the landing area"); return; }
  if (b->rank == root) {
    PeerBases pb;
    shard_peer_bases(b, which, &pb);
    for (int r = 0; r < pb.n; ++r) std::memcpy(pb.base[r], buf, (size_t)bytes);
    b->launches++;
  }
  shard_barrier(b);
  if (b->rank != root) std::memcpy(buf, shard_landing(b, which), (size_t)bytes);
}
void shard_allgather_p2p(Backend* b, void* buf, int64_t total_bytes, int64_t chunk_bytes, int which) {
  if (b->world <= 1 || total_bytes <= 0) return;
  if (total_bytes > b->landing_each || (chunk_bytes & 15)) { emu_err(b, "shard_allgather_p2p: message larger than the landing area / chunk not 16-byte granular"); return; }
  const int64_t off = std::min<int64_t>((int64_t)b->rank * chunk_bytes, total_bytes);
  const int64_t mine = std::min<int64_t>(chunk_bytes, total_bytes - off);
  if (mine > 0) {
    PeerBases pb;
    shard_peer_bases(b, which, &pb);
    for (int r = 0; r < pb.n; ++r) std::memcpy(pb.base[r] + off, (const char*)buf + off, (size_t)mine);
    b->launches++;
  }
  shard_barrier(b);
  const char* land = (const char*)shard_landing(b, which);
  if (off > 0) std::memcpy(buf, land, (size_t)off);
  if (off + mine < total_bytes) std::memcpy((char*)buf + off + mine, land + off + mine, (size_t)(total_bytes - off - mine));
}

}  // namespace cb200
