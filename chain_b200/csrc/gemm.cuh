// Grouped, segment-accumulating FP64 GEMM on the sm_100a DMMA pipe.
//
// One launch executes a whole task list (tasks.h): every CTA owns one 128x128 (real) or 64x128 (complex)
// tile of one task's C and runs a single multi-stage cp.async pipeline over the concatenation of all of that
// task's K-segments, so that  C = sum_seg op(A_seg) * op(B_seg)  is formed in registers without ever
// materialising the per-segment products.  This is how the environment x MPO x wavefunction contractions of
// LocalMPO::product / makeL / makeR (ITensor, called from /root/reference/src/dmrg.h:544,563,596) are executed:
// the index permutations ITensor performs as explicit copies are absorbed into the (lda, ldb, transpose)
// addressing of the tile loads.
//
// FP64 on sm_100a has no tcgen05 kind; the tensor instruction is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
// Fragment layout (PTX ISA): A[row=lane>>2][k=lane&3], B[k=lane&3][col=lane>>2], C[row=lane>>2][col=2*(lane&3)+{0,1}].
// Shared-memory tiles are padded so that the 64-bit (half-warp) / 128-bit (quarter-warp) fragment loads are
// bank-conflict free in both operand orientations ("DMMA GEMM").
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include "tasks.h"

namespace cb200 {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- mbarrier helpers (full/empty pipeline: no CTA-wide barrier inside the k loop) -------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_LOOP:\n"
      " mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
      " @p bra WAIT_DONE;\n"
      " bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
// all cp.async issued so far by this thread arrive on `bar` when they have landed (does not change the expected count)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];\n" ::"r"(a) : "memory");
}

// ---- bulk asynchronous copies (the TMA engine, SASS UBLKCP): one instruction moves a whole contiguous tile row global -> shared and
// reports its bytes to an mbarrier (complete_tx).  Addresses and sizes must be multiples of 16 bytes.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared.b64 st, [%0], %1;\n}\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s), "l"(gmem), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// 1 = complex operand tiles are staged with bulk copies by ONE producer warp (set from CB200_GEMM_BULK at context creation; default on)
__device__ int g_gemm_bulk = 1;

// Stage one complex operand tile with bulk copies: a tile row (contiguous in global memory) per copy, issued by the lanes of one warp.
//   MN_CONTIG: rows = k (kvalid of them, mnvalid elements each), smem row stride LD = BMN + PAD;  else rows = mn (mnvalid, kvalid elements), LD = LDK.
// Rows / row tails past K must read as zero (they enter every output element) and are written with plain stores; elements past MN are left
// stale (they only reach output rows / columns that are never stored).
template <int BMN, int BK, int LD, bool MN_CONTIG>
__device__ __forceinline__ void bulk_operand_zero_tail(char* smem, int kvalid, int lane) {
  const double2 z2 = make_double2(0.0, 0.0);
  double2* s2 = reinterpret_cast<double2*>(smem);
  if constexpr (MN_CONTIG) {
    for (int idx = lane; idx < (BK - kvalid) * BMN; idx += 32) s2[(kvalid + idx / BMN) * LD + idx % BMN] = z2;
  } else {
    for (int idx = lane; idx < (BK - kvalid) * BMN; idx += 32) s2[(idx / (BK - kvalid)) * LD + kvalid + idx % (BK - kvalid)] = z2;
  }
}
template <int BMN, int BK, int LD, bool MN_CONTIG>
__device__ __forceinline__ void bulk_operand_tile(char* smem, const char* g, int64_t ld, int mn0, int mnvalid, int k0, int kvalid, uint64_t* bar, int lane) {
  if constexpr (MN_CONTIG) {
    for (int k = lane; k < kvalid; k += 32)
      bulk_g2s(smem + (size_t)k * LD * 16, g + ((int64_t)mn0 + (int64_t)(k0 + k) * ld) * 16, (unsigned)mnvalid * 16u, bar);
  } else {
    for (int mn = lane; mn < mnvalid; mn += 32)
      bulk_g2s(smem + (size_t)mn * LD * 16, g + ((int64_t)k0 + (int64_t)(mn0 + mn) * ld) * 16, (unsigned)kvalid * 16u, bar);
  }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <bool CPLX, int BN_ = 128>
struct GemmCfg {
  // real:    CTA 128x128x16, 8 warps as 2(m) x 4(n), warp tile 64x32  -> 8x4 DMMA tiles
  // complex: CTA  64x128x16, 8 warps as 2(m) x 4(n), warp tile 32x32  -> 4x4 complex tiles (4 DMMA each)
  static constexpr int BM = CPLX ? 64 : 128;
  static constexpr int BN = BN_;                     // 128 (default) or 64 (narrow outputs: Jacobi rotations, Gram blocks)
  // k-tile depth / pipeline stages.  Measured on cfg2 (c128): BK 8 x 4 stages -> 0.84 of the cuBLAS DGEMM rate, BK 16 x 3 stages ->
  // 0.90: the per-k-tile producer code (addressing + cp.async issue, ~300 instructions) is amortised over twice the DMMAs.
  // The real kernel is insensitive (BK 16 x 4 and BK 32 x 3 both 0.87).  Overridable for A/B builds.
#ifndef CB200_BK_CPLX
#define CB200_BK_CPLX 16
#endif
#ifndef CB200_BK_REAL
#define CB200_BK_REAL 16
#endif
#ifndef CB200_STAGES_CPLX
#define CB200_STAGES_CPLX 3
#endif
#ifndef CB200_STAGES_REAL
#define CB200_STAGES_REAL 4
#endif
  static constexpr int BK = CPLX ? CB200_BK_CPLX : CB200_BK_REAL;
  static constexpr int WM = CPLX ? 32 : 64;
  static constexpr int WN = BN_ / 4;
  static constexpr int MT = WM / 8;
  static constexpr int NT = WN / 8;
  static constexpr int ES = CPLX ? 16 : 8;           // element bytes
  static constexpr int PAD_MN = CPLX ? 2 : 4;        // [k][mn] layout: LD = B{M,N} + PAD_MN
  static constexpr int LDK = BK + 4;                 // [mn][k] layout
  static constexpr int A_ELEMS = (BK * (BM + PAD_MN) > BM * LDK) ? BK * (BM + PAD_MN) : BM * LDK;
  static constexpr int B_ELEMS = (BK * (BN + PAD_MN) > BN * LDK) ? BK * (BN + PAD_MN) : BN * LDK;
  static constexpr int STAGE_BYTES = (A_ELEMS + B_ELEMS) * ES;
#ifndef CB200_STAGES_CPLX_NARROW
#define CB200_STAGES_CPLX_NARROW 3
#endif
  // (the 64x128 complex tile fits 3 stages of k-tile 16 in 227 KB; the 64x96 / 64x64 tiles of the 3M kernel would fit 4: measured no gain)
  static constexpr int STAGES = CPLX ? (BN_ <= 96 ? CB200_STAGES_CPLX_NARROW : CB200_STAGES_CPLX) : CB200_STAGES_REAL;
  static constexpr int AHEAD = STAGES - 2;            // k-tiles in flight; one stage of slack decouples the warps
  static constexpr int SMEM_BYTES = STAGE_BYTES * STAGES;
  static constexpr int CONSUMERS = 256;             // 8 DMMA warps
  static constexpr int PRODUCERS = 128;             // 1 warpgroup that only issues the cp.async tile loads (warp specialisation)
  // setmaxnreg budgets.  The CTA's register pool is fixed at launch: 168 registers (what __launch_bounds__(384,1) yields) x 384
  // threads; setmaxnreg.inc BLOCKS until the pool can serve it, so the two budgets must add up to no more than the pool.
  static constexpr int LAUNCH_REGS = 168;
  static constexpr int PRODUCER_REGS = 40;
  static constexpr int CONSUMER_REGS = 232;
  static_assert(CONSUMER_REGS * CONSUMERS + PRODUCER_REGS * PRODUCERS <= LAUNCH_REGS * (CONSUMERS + PRODUCERS), "setmaxnreg budgets exceed the CTA register pool (deadlock)");
  static constexpr int THREADS = CONSUMERS + PRODUCERS;
};

// Load one operand tile (BMN x BK elements) into shared memory.
//   MN_CONTIG = true : global element (mn,k) at g[mn + k*ld]; smem layout [k][mn], LD = BMN+PAD
//   MN_CONTIG = false: global element (mn,k) at g[k + mn*ld]; smem layout [mn][k], LD = BK+4
// Out-of-range elements are zero-filled by cp.async's src-size operand.
template <bool CPLX, int BN_, int BMN, bool MN_CONTIG>
__device__ __forceinline__ void load_operand_tile(char* smem, const char* g, int64_t ld, int mn0, int MN, int k0, int K,
                                                  bool vec_ok, int tid) {
  using Cfg = GemmCfg<CPLX, BN_>;
  constexpr int BK = Cfg::BK;
  constexpr int ES = Cfg::ES;
  constexpr int LD = MN_CONTIG ? (BMN + Cfg::PAD_MN) : Cfg::LDK;
  if constexpr (CPLX) {
    // one 16-byte element per copy
    constexpr int TOTAL = BMN * BK;
#pragma unroll 4
    for (int c = tid; c < TOTAL; c += Cfg::PRODUCERS) {
      int mn, k;
      if constexpr (MN_CONTIG) { k = c / BMN; mn = c % BMN; } else { mn = c / BK; k = c % BK; }
      const bool ok = (mn0 + mn < MN) && (k0 + k < K);
      const int64_t goff = MN_CONTIG ? ((int64_t)(mn0 + mn) + (int64_t)(k0 + k) * ld) : ((int64_t)(k0 + k) + (int64_t)(mn0 + mn) * ld);
      const int soff = MN_CONTIG ? (k * LD + mn) : (mn * LD + k);
      cp_async16(smem + (size_t)soff * ES, ok ? (g + goff * ES) : g, ok ? 16 : 0);
    }
  } else {
    // two doubles per 16-byte copy along the contiguous direction
    constexpr int TOTAL = BMN * BK / 2;
#pragma unroll 4
    for (int c = tid; c < TOTAL; c += Cfg::PRODUCERS) {
      int mn, k, rem;
      int64_t goff;
      int soff;
      if constexpr (MN_CONTIG) {
        k = c / (BMN / 2); mn = (c % (BMN / 2)) * 2;
        rem = (k0 + k < K) ? (MN - (mn0 + mn)) : 0;
        goff = (int64_t)(mn0 + mn) + (int64_t)(k0 + k) * ld;
        soff = k * LD + mn;
      } else {
        mn = c / (BK / 2); k = (c % (BK / 2)) * 2;
        rem = (mn0 + mn < MN) ? (K - (k0 + k)) : 0;
        goff = (int64_t)(k0 + k) + (int64_t)(mn0 + mn) * ld;
        soff = mn * LD + k;
      }
      rem = rem < 0 ? 0 : (rem > 2 ? 2 : rem);
      char* sp = smem + (size_t)soff * 8;
      const char* gp = rem ? (g + goff * 8) : g;
      if (vec_ok) {
        cp_async16(sp, gp, rem * 8);
      } else {
        cp_async8(sp, gp, rem >= 1 ? 8 : 0);
        cp_async8(sp + 8, rem >= 2 ? gp + 8 : g, rem >= 2 ? 8 : 0);
      }
    }
  }
}

template <bool CPLX, bool TA, bool TB, int BN_, bool PEER, bool M3>
__device__ __forceinline__ void gemm_grouped_body(const char* __restrict__ Abase, const char* __restrict__ Bbase, char* __restrict__ Cbase,
                                                  const GemmTask* __restrict__ tasks, int ntasks, const GemmSeg* __restrict__ segs,
                                                  const PeerBases& peers) {
  using Cfg = GemmCfg<CPLX, BN_>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, MT = Cfg::MT, NT = Cfg::NT, ES = Cfg::ES, STAGES = Cfg::STAGES;
  // smem strides (elements) of the fragment reads
  constexpr int A_SM = TA ? Cfg::LDK : 1;                       // stride along m
  constexpr int A_SK = TA ? 1 : (BM + Cfg::PAD_MN);             // stride along k
  constexpr int B_SN = TB ? 1 : Cfg::LDK;                       // stride along n
  constexpr int B_SK = TB ? (BN + Cfg::PAD_MN) : 1;             // stride along k

  extern __shared__ __align__(16) char smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int wm0 = (warp & 1) * Cfg::WM;
  const int wn0 = (warp >> 1) * Cfg::WN;

  // locate the task owning this tile (tile_begin is ascending)
  int lo = 0, hi = ntasks - 1;
  const int tile = blockIdx.x;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (tasks[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
  }
  const GemmTask t = tasks[lo];
  const int lt = tile - t.tile_begin;
  const int tiles_n = (t.N + BN - 1) / BN;
  const bool n_fast = (t.flags & GEMM_TILE_N_FAST) != 0;
  const int m0 = (n_fast ? lt / tiles_n : lt % t.tiles_m) * BM;
  const int n0 = (n_fast ? lt % tiles_n : lt / t.tiles_m) * BN;
  const GemmSeg* sg = segs + t.seg_begin;

  int total_kt = 0;
  for (int s = 0; s < t.seg_count; ++s) total_kt += (sg[s].K + BK - 1) / BK;

  // full[s]: the k-tile in stage s has landed (the producer lanes' cp.async arrive on it);
  // empty[s]: every consumer warp has finished reading stage s.
  // Warp specialisation: warps 8-11 do nothing but address arithmetic + cp.async for the whole CTA, run up to STAGES k-tiles
  // ahead and hand their registers to the DMMA warps (setmaxnreg); warps 0-7 never execute producer code, so the DMMA pipe is
  // fed by whichever of them has a landed tile.
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  // Bulk (TMA-engine) staging needs 16-byte aligned row starts -- always true for complex (16-byte elements), not for the odd leading
  // dimensions of real Z_N sectors -- and pays off only when a tile row is long: with both operands contiguous along M / N (the R-stage
  // and peer kernels, every rank-k update) a k-tile is 16 + 16 copies of >= 1 KB; a K-contiguous operand would be 64..128 copies of 256 B,
  // measured 28 % slower than the per-thread cp.async path on the L-stage of BASELINE configs[1] (profiles/r02_bench_cfg2_n4.json).
  const bool bulk = CPLX && !TA && TB && g_gemm_bulk != 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], bulk ? 1u : (unsigned)Cfg::PRODUCERS); mbar_init(&empty_bar[s], Cfg::CONSUMERS / 32); }
  }
  __syncthreads();

  if (warp >= Cfg::CONSUMERS / 32) {
    // ---------------------------------------------------------------- producer warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(Cfg::PRODUCER_REGS));
    const int ptid = tid - Cfg::CONSUMERS;
    int ps = 0, pk = 0;
    if constexpr (CPLX) {
      if (bulk) {
        // ---- TMA-engine producer: ONE warp issues a bulk copy per tile row and arms the stage's mbarrier with the byte count
        if (ptid >= 32) return;
        const int mvalid = (t.M - m0) < BM ? (t.M - m0) : BM, nvalid = (t.N - n0) < BN ? (t.N - n0) : BN;
        for (int itl = 0; itl < total_kt; ++itl) {
          const int stage = itl % STAGES, use = itl / STAGES;
          if (use > 0) mbar_wait(&empty_bar[stage], (unsigned)((use - 1) & 1));
          while (sg[ps].K <= 0) ++ps;
          const GemmSeg s = sg[ps];
          char* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
          char* sb = sa + (size_t)Cfg::A_ELEMS * ES;
          const int kvalid = (s.K - pk) < BK ? (s.K - pk) : BK;
          if (kvalid < BK) {
            // last k-tile of a segment: the rows / row tails past K enter every output element and must read as zero.  Plain stores, made
            // visible to the consumers by the release of lane 0's arrive below (after the warp-level barrier), and ordered before later
            // async-proxy writes to the same bytes
            bulk_operand_zero_tail<BM, BK, TA ? Cfg::LDK : (BM + Cfg::PAD_MN), !TA>(sa, kvalid, ptid);
            bulk_operand_zero_tail<BN, BK, TB ? (BN + Cfg::PAD_MN) : Cfg::LDK, TB>(sb, kvalid, ptid);
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (ptid == 0) mbar_arrive_expect_tx(&full_bar[stage], (unsigned)(kvalid * (mvalid + nvalid)) * 16u);
          __syncwarp();
          bulk_operand_tile<BM, BK, TA ? Cfg::LDK : (BM + Cfg::PAD_MN), !TA>(sa, Abase + s.a_off * ES, s.lda, m0, mvalid, pk, kvalid, &full_bar[stage], ptid);
          bulk_operand_tile<BN, BK, TB ? (BN + Cfg::PAD_MN) : Cfg::LDK, TB>(sb, Bbase + s.b_off * ES, s.ldb, n0, nvalid, pk, kvalid, &full_bar[stage], ptid);
          pk += BK;
          if (pk >= s.K) { pk = 0; ++ps; }
        }
        return;
      }
    }
    for (int itl = 0; itl < total_kt; ++itl) {
      const int stage = itl % STAGES, use = itl / STAGES;
      if (use > 0) mbar_wait(&empty_bar[stage], (unsigned)((use - 1) & 1));
      while (sg[ps].K <= 0) ++ps;   // empty segments own no k-tile
      const GemmSeg s = sg[ps];
      char* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
      char* sb = sa + (size_t)Cfg::A_ELEMS * ES;
      const char* ga = Abase + s.a_off * ES;
      const char* gb = Bbase + s.b_off * ES;
      bool va = true, vb = true;
      if constexpr (!CPLX) {
        va = ((((uintptr_t)ga) & 15) == 0) && ((s.lda & 1) == 0);
        vb = ((((uintptr_t)gb) & 15) == 0) && ((s.ldb & 1) == 0);
      }
      load_operand_tile<CPLX, BN_, BM, !TA>(sa, ga, s.lda, m0, t.M, pk, s.K, va, ptid);
      load_operand_tile<CPLX, BN_, BN, TB>(sb, gb, s.ldb, n0, t.N, pk, s.K, vb, ptid);
      cp_async_arrive(&full_bar[stage]);
      pk += BK;
      if (pk >= s.K) { pk = 0; ++ps; }
    }
    return;   // the producer has no part in the epilogue (its cp.async have all been waited on by the consumers' full barriers)
  }

  // -------------------------------------------------------------------- consumer warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(Cfg::CONSUMER_REGS));
  // complex, standard: acc_re = Re, acc_im = Im (4 DMMA per complex MAC).  complex, M3 (Karatsuba "3M"): acc_re = P1 = sum ar*br,
  // acc_im = P2 = sum ai*bi, acc_p3 = P3 = sum (ar+ai)(br+bi) -- 3 DMMA per complex MAC, Re = P1 - P2, Im = P3 - P1 - P2 formed
  // once in the epilogue (normwise backward stable; the FP64 adds of the fragment sums run on the DFMA pipe, not the DMMA pipe).
  double acc_re[MT][NT][2];
  double acc_im[CPLX ? MT : 1][CPLX ? NT : 1][2];
  double acc_p3[(CPLX && M3) ? MT : 1][(CPLX && M3) ? NT : 1][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      acc_re[i][j][0] = 0.0; acc_re[i][j][1] = 0.0;
      if constexpr (CPLX) { acc_im[i][j][0] = 0.0; acc_im[i][j][1] = 0.0; }
      if constexpr (CPLX && M3) { acc_p3[i][j][0] = 0.0; acc_p3[i][j][1] = 0.0; }
    }

  const bool conjA = (t.flags & GEMM_CONJ_A) != 0;
  for (int it = 0; it < total_kt; ++it) {
    mbar_wait(&full_bar[it % STAGES], (unsigned)((it / STAGES) & 1));

    const char* sa = smem + (size_t)(it % STAGES) * Cfg::STAGE_BYTES;
    const char* sb = sa + (size_t)Cfg::A_ELEMS * ES;
    if constexpr (!CPLX) {
      const double* As = reinterpret_cast<const double*>(sa) + (wm0 + g) * A_SM + tq * A_SK;
      const double* Bs = reinterpret_cast<const double*>(sb) + (wn0 + g) * B_SN + tq * B_SK;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[MT], b[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) a[i] = As[i * 8 * A_SM + kk * 4 * A_SK];
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = Bs[j * 8 * B_SN + kk * 4 * B_SK];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) dmma884(acc_re[i][j][0], acc_re[i][j][1], a[i], b[j]);
      }
    } else {
      const double2* As = reinterpret_cast<const double2*>(sa) + (wm0 + g) * A_SM + tq * A_SK;
      const double2* Bs = reinterpret_cast<const double2*>(sb) + (wn0 + g) * B_SN + tq * B_SK;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double2 a[MT], b[NT];
        double nai[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          a[i] = As[i * 8 * A_SM + kk * 4 * A_SK];
          if (conjA) a[i].y = -a[i].y;
          nai[i] = M3 ? a[i].x + a[i].y : -a[i].y;          // M3: the fragment sum ar + ai
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = Bs[j * 8 * B_SN + kk * 4 * B_SK];
        if constexpr (M3) {
          double bs[NT];
#pragma unroll
          for (int j = 0; j < NT; ++j) bs[j] = b[j].x + b[j].y;
#pragma unroll
          for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              dmma884(acc_re[i][j][0], acc_re[i][j][1], a[i].x, b[j].x);
              dmma884(acc_im[i][j][0], acc_im[i][j][1], a[i].y, b[j].y);
              dmma884(acc_p3[i][j][0], acc_p3[i][j][1], nai[i], bs[j]);
            }
        } else {
#pragma unroll
          for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              dmma884(acc_re[i][j][0], acc_re[i][j][1], a[i].x, b[j].x);
              dmma884(acc_im[i][j][0], acc_im[i][j][1], a[i].x, b[j].y);
              dmma884(acc_re[i][j][0], acc_re[i][j][1], nai[i], b[j].y);
              dmma884(acc_im[i][j][0], acc_im[i][j][1], a[i].y, b[j].x);
            }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[it % STAGES]);
  }

  if constexpr (CPLX && M3) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double p1 = acc_re[i][j][e], p2 = acc_im[i][j][e], p3 = acc_p3[i][j][e];
          acc_re[i][j][e] = p1 - p2;
          acc_im[i][j][e] = p3 - p1 - p2;
        }
  }
  // epilogue: direct stores from the accumulator fragments
  const bool accum = (t.flags & GEMM_ACCUM) != 0;
  const bool conj_out = (t.flags & GEMM_CONJ_OUT) != 0;
  const bool trans_c = (t.flags & GEMM_TRANS_C) != 0;
  char* Cp = Cbase + t.c_off * ES;
  auto c_index = [&](int m, int n) -> int64_t {
    const int64_t mc = (int64_t)(m % t.chunk) + (int64_t)(m / t.chunk) * t.chunk_stride;
    return trans_c ? ((int64_t)(n % t.chunk) + (int64_t)(n / t.chunk) * t.chunk_stride + (int64_t)m * t.ldc)
           : (n < t.n_split) ? (mc + (int64_t)n * t.ldc)
                             : ((t.c_off2 - t.c_off) + mc + (int64_t)(n - t.n_split) * t.ldc);
  };
  if constexpr (!PEER) {
    if (accum) {
      // C += result.  The old values are fetched IB row groups at a time BEFORE any of their stores is issued: a load placed after a store
      // to the same array cannot be hoisted by the compiler, and one exposed L2/HBM round trip per element (64 per thread) cost ~80 us per
      // tile -- ten times the main loop of a rank-64 update.
      constexpr int IB = 2;
#pragma unroll
      for (int i0 = 0; i0 < MT; i0 += IB) {
        using ET = typename std::conditional<CPLX, double2, double>::type;
        ET old[IB][NT][2];
#pragma unroll
        for (int ii = 0; ii < IB; ++ii) {
          const int m = m0 + wm0 + (i0 + ii) * 8 + g;
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int n = n0 + wn0 + j * 8 + tq * 2 + e;
              if constexpr (CPLX) old[ii][j][e] = make_double2(0.0, 0.0); else old[ii][j][e] = 0.0;
              if (i0 + ii < MT && m < t.M && n < t.N) old[ii][j][e] = __ldcg(reinterpret_cast<const ET*>(Cp) + c_index(m, n));
            }
        }
#pragma unroll
        for (int ii = 0; ii < IB; ++ii) {
          const int i = i0 + ii;
          const int m = m0 + wm0 + i * 8 + g;
          if (i >= MT || m >= t.M) continue;
#pragma unroll
          for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int n = n0 + wn0 + j * 8 + tq * 2 + e;
              if (n >= t.N) continue;
              const int64_t off = c_index(m, n);
              if constexpr (!CPLX) {
                reinterpret_cast<double*>(Cp)[off] = acc_re[i][j][e] + old[ii][j][e];
              } else {
                const double2 o = old[ii][j][e];
                reinterpret_cast<double2*>(Cp)[off] = make_double2(acc_re[i][j][e] + o.x, (conj_out ? -acc_im[i][j][e] : acc_im[i][j][e]) + o.y);
              }
            }
        }
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int m = m0 + wm0 + i * 8 + g;
    if (m >= t.M) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = n0 + wn0 + j * 8 + tq * 2 + e;
        if (n >= t.N) continue;
        const int64_t off = c_index(m, n);
        if constexpr (PEER) {
          // all-gather fused into the epilogue: the tile goes to every rank's landing buffer over NVLink (plain stores;
          // the cross-GPU barrier that follows the launch publishes them)
          for (int pr = 0; pr < peers.n; ++pr) {
            char* Pp = peers.base[pr] + t.c_off * ES;
            if constexpr (!CPLX) reinterpret_cast<double*>(Pp)[off] = acc_re[i][j][e];
            else reinterpret_cast<double2*>(Pp)[off] = make_double2(acc_re[i][j][e], conj_out ? -acc_im[i][j][e] : acc_im[i][j][e]);
          }
          continue;
        }
        if constexpr (!CPLX) {
          reinterpret_cast<double*>(Cp)[off] = acc_re[i][j][e];
        } else {
          reinterpret_cast<double2*>(Cp)[off] = make_double2(acc_re[i][j][e], conj_out ? -acc_im[i][j][e] : acc_im[i][j][e]);
        }
      }
    }
  }
}

template <bool CPLX, bool TA, bool TB, int BN_ = 128, bool M3 = false>
__global__ void __launch_bounds__(GemmCfg<CPLX, BN_>::THREADS, 1)
gemm_grouped_kernel(const char* __restrict__ Abase, const char* __restrict__ Bbase, char* __restrict__ Cbase,
                    const GemmTask* __restrict__ tasks, int ntasks, const GemmSeg* __restrict__ segs) {
  PeerBases none;
  none.n = 0;
  gemm_grouped_body<CPLX, TA, TB, BN_, false, M3>(Abase, Bbase, Cbase, tasks, ntasks, segs, none);
}

// R-stage of the row-sharded matvec: same mainloop, C tiles stored to every rank's landing buffer
template <bool CPLX, int BN_ = 128, bool M3 = false>
__global__ void __launch_bounds__(GemmCfg<CPLX, BN_>::THREADS, 1)
gemm_grouped_peer_kernel(const char* __restrict__ Abase, const char* __restrict__ Bbase, const GemmTask* __restrict__ tasks, int ntasks,
                         const GemmSeg* __restrict__ segs, const __grid_constant__ PeerBases peers) {
  gemm_grouped_body<CPLX, false, true, BN_, true, M3>(Abase, Bbase, nullptr, tasks, ntasks, segs, peers);
}

}  // namespace cb200
