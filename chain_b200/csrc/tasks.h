// Task descriptors shared by the host-side planner (engine.cpp) and the device kernels.
// Everything the hot path does on the GPU is expressed as flat lists of these PODs, uploaded once per
// bond shape and replayed for every Davidson matvec.  All matrices are column-major.
#pragma once
#include <cstdint>

namespace cb200 {

enum : uint32_t {
  GEMM_TRANS_A = 1u,   // A is stored K x M (k contiguous); element (m,k) at A[k + m*lda]
  GEMM_TRANS_B = 2u,   // B is stored N x K (n contiguous); element (k,n) at B[n + k*ldb]
  GEMM_CONJ_A  = 4u,   // complex only: use conj(A)
  GEMM_ACCUM   = 8u,   // C += result (otherwise C = result)
  GEMM_CONJ_OUT = 16u, // complex only: store conj(result)
  GEMM_TRANS_C = 32u,  // store element (m,n) at C[n + m*ldc] (lets the planner pick the operand roles that waste fewer tile slots)
  GEMM_TILE_N_FAST = 64u,  // number the task's tiles N-fastest: concurrently resident CTAs then share the A rows (set by the planner when
                           // A is the large streamed operand, e.g. the 12.8 GB intermediate of the R-stage at D = 1600, and B fits in L2)
};

// One K-segment of a task: C += op(A_seg) * op(B_seg).  Offsets are in elements from the base pointers
// passed at launch (so one task list can be replayed on different buffers, e.g. different Krylov vectors).
struct GemmSeg {
  int64_t a_off, b_off;
  int64_t lda, ldb;
  int32_t K;
  int32_t pad_;
};

// C (M x N, ldc) = sum over segments.  Tiles of one task are numbered tile_begin .. tile_begin+tiles-1,
// M-fastest, so that concurrently resident CTAs share B tiles through L2 (N-fastest with GEMM_TILE_N_FAST).
struct GemmTask {
  int64_t c_off, ldc;
  int64_t c_off2;       // columns n >= n_split are stored at c_off2 + (n - n_split)*ldc (block-Jacobi column shuffles)
  int32_t M, N;
  int32_t seg_begin, seg_count;
  int32_t tile_begin, tiles_m;
  uint32_t flags;
  int32_t n_split;      // set to N (or larger) when unused
  // Chunked addressing of the contiguous C index (m, or n with GEMM_TRANS_C): index i is stored at
  // (i % chunk) + (i / chunk) * chunk_stride.  A row-sharded matvec writes its (l'_local, s) rows straight into the
  // full [l, s, r] layout of the output vector this way (chunk = local rows, chunk_stride = full left dimension).
  // chunk = INT32_MAX disables it.
  int32_t chunk = 0x7fffffff;
  int32_t pad_ = 0;
  int64_t chunk_stride = 0;
};

// Destinations of a GEMM launched with peer stores: the same C tile goes to every base pointer (own landing buffer and the
// peers' buffers mapped through CUDA IPC) -- the all-gather of the row-sharded matvec, fused into the R-stage epilogue.
constexpr int MAX_PEERS = 8;
struct PeerBases {
  char* base[MAX_PEERS];
  int32_t n;
  int32_t pad_;
};

// "Mix": the bandwidth-bound application of the (sparse) MPO tensors between the two GEMM stages.
// A panel is a set of input column slots and output column slots sharing a row count and a column
// count nr; out[row, oslot, r] = sum_e coef[e] * in[row, islot(e), r].
struct MixOut {
  int64_t off, rstride;   // element address = out_base + off + r*rstride + row
  int32_t e_begin, e_end; // range into the entry list
};
struct MixIn {
  int64_t off, rstride;         // element address = in_base + off + r*rstride + row
};
struct MixEntry {                // 32 bytes: fetched as two 16-byte uniform loads
  double cre, cim;
  int32_t slot;                 // index into the panel's MixIn table (in_begin + slot)
  int32_t pad_[3];
};
struct MixPanel {
  int32_t rows, nr;
  int32_t out_begin, out_end;   // range into the MixOut list
  int32_t in_begin, in_count;   // the distinct input slots of this panel (staged in shared memory)
  int32_t blk_begin;            // first CTA of this panel in the flattened grid
  int32_t row_chunks;           // CTAs along rows
};

// Strided gather/scatter (permutes between the phi layout, the Jacobi work matrices and MPS tensors).
// dst[dst_off + i*ds0 + j*ds1 + k*ds2] = (conj?) src[src_off + i*ss0 + j*ss1 + k*ss2], i<n0 (fastest), j<n1, k<n2.
struct CopyTask {
  int64_t src_off, dst_off;
  int64_t ss0, ss1, ss2, ds0, ds1, ds2;
  int32_t n0, n1, n2;
  int32_t conj;
  int32_t blk_begin, pad_;
};

}  // namespace cb200
