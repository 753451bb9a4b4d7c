// CUDA (sm_100a) implementation of backend.h: the kernels of the Chain DMRG hot path.
//   * grouped DMMA GEMM            -> gemm.cuh            (H_eff stages, environment update, rho/Gram, projections)
//   * mix_kernel                   -> sparse W (x) W application between the GEMM stages (HBM/L2-bandwidth bound)
//   * copy_kernel                  -> strided permutes / gathers
//   * dot_partial/dot_final, lincomb_kernel -> Davidson orthogonalisation + penalty projectors, Krylov basis resident in HBM
//   * jacobi_eig_kernel            -> 2JB x 2JB Hermitian eigensolver in shared memory (block one-sided Jacobi SVD)
// There is no CPU fallback here: every entry point launches a kernel or reports a CUDA error.
#include <cuda_runtime.h>
#include <atomic>
#include <dlfcn.h>
#include <nccl.h>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <type_traits>
#include <vector>
#include "backend.h"
#include "gemm.cuh"
#include "heig.cuh"
#include "heig2.cuh"

namespace cb200 {

struct Backend {
  int device = 0;
  cudaStream_t stream = nullptr;
  double* d_partial = nullptr;   // multi_dot partials
  double* d_dot_out = nullptr;
  double* h_dot_out = nullptr;   // pinned
  uint64_t launches = 0;
  std::string err;
  bool has_err = false;
  cudaEvent_t ev[8] = {};
  // Exact-size block cache in front of the stream-ordered pool.  A sweep asks for the same multi-GB temporaries bond after bond;
  // the pool sub-allocates and fragments, so large requests kept falling through to new physical allocations (seconds per
  // bond at D = 3000).  Freed blocks >= 32 MiB are parked here and handed back to the next request of (nearly) the same size; all work is
  // on one stream, so reuse is ordered.  Bounded: the largest parked blocks are released when the cache exceeds its cap.
  std::unordered_map<void*, size_t> live_size;
  std::multimap<size_t, void*> parked;
  size_t parked_bytes = 0, parked_cap = (size_t)48 << 30;
  // Pinned staging ring for small host->device uploads (task lists, scalars): the data is copied into the ring and sent with an
  // asynchronous copy, so the dozens of descriptor uploads per bond no longer cost a stream synchronisation each.  The stream is
  // synchronised only when the ring wraps.
  char* h_ring = nullptr;
  size_t ring_bytes = (size_t)8 << 20, ring_off = 0;
  double* d_bounds = nullptr;     // eigensolver scratch (Gershgorin bounds, pivmin, ||T||)
  // multi-GPU: symmetric area [flags (256 B) | landing 0 | landing 1] + the peers' mappings (CUDA IPC)
  char* sym = nullptr;
  int64_t landing_each = 0;
  int rank = 0, world = 1;
  char* peer_sym[MAX_PEERS] = {};
  bool peer_opened[MAX_PEERS] = {};
  uint64_t barrier_seq = 0;
  // NCCL through dlopen: no link-time dependency, the library still loads on a CPU-only box
  void* nccl_lib = nullptr;
  ncclComm_t comm = nullptr;
  ncclResult_t (*p_ncclGetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*p_ncclCommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*p_ncclAllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*p_ncclBroadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*p_ncclCommDestroy)(ncclComm_t) = nullptr;
  const char* (*p_ncclGetErrorString)(ncclResult_t) = nullptr;
};
static constexpr int64_t SYM_FLAG_BYTES = 256;

static constexpr int DOT_BLOCKS = 592;   // 4 x 148 SMs
static constexpr int DOT_THREADS = 256;

static void set_err(Backend* b, const char* what, cudaError_t e) {
  if (!b->has_err) {
    b->err = std::string(what) + ": " + cudaGetErrorString(e);
    b->has_err = true;
  }
}
#define CB_CHECK(b, call)                                   \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) set_err((b), #call, e_);         \
  } while (0)
#define CB_LAUNCH_CHECK(b, name)                            \
  do {                                                      \
    (b)->launches++;                                        \
    cudaError_t e_ = cudaGetLastError();                    \
    if (e_ != cudaSuccess) set_err((b), name, e_);          \
  } while (0)

const char* backend_name() { return "cuda-sm_100a"; }
int backend_device(Backend* b) { return b->device; }

Backend* backend_create(int device, char* err, size_t errlen) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    snprintf(err, errlen, "chain_b200: no CUDA device available (%s); this library has no CPU fallback",
             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return nullptr;
  }
  if (device < 0 || device >= ndev) {
    snprintf(err, errlen, "chain_b200: device %d out of range (%d devices)", device, ndev);
    return nullptr;
  }
  // One device per process.  Kernel attributes (dynamic shared memory opt-ins), occupancy caches and the current-device setting are
  // process-wide state; the multi-GPU design is one process per GPU, so a second context on a DIFFERENT device is refused
  // instead of silently launching on the wrong one.
  {
    static std::atomic<int> bound_device{-1};
    int expected = -1;
    if (!bound_device.compare_exchange_strong(expected, device) && expected != device) {
      snprintf(err, errlen, "chain_b200: this process already drives device %d; one device per process (launch one process per GPU)", expected);
      return nullptr;
    }
  }
  cudaSetDevice(device);
  Backend* b = new Backend();
  b->device = device;
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&b->d_partial, sizeof(double) * 2 * MAX_LC * DOT_BLOCKS) != cudaSuccess ||
      cudaMalloc(&b->d_dot_out, sizeof(double) * 2 * MAX_LC) != cudaSuccess ||
      cudaMallocHost(&b->h_dot_out, sizeof(double) * 2 * MAX_LC) != cudaSuccess ||
      cudaMallocHost(&b->h_ring, b->ring_bytes) != cudaSuccess) {
    snprintf(err, errlen, "chain_b200: backend init failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete b;
    return nullptr;
  }
  for (auto& ev : b->ev) cudaEventCreate(&ev);
  // keep freed blocks in the stream-ordered pool across synchronisations (the default threshold of 0 hands them back to
  // the driver at every cudaStreamSynchronize, which turns each per-bond temporary into a cudaMalloc/cudaFree pair)
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  // opt in to the large dynamic shared memory the GEMM / Jacobi kernels need
  {
    auto opt = [](const void* f, int bytes) { cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); };
#define CB_OPT(C, TA, TB)                                                                           \
    opt((const void*)gemm_grouped_kernel<C, TA, TB, 128>, GemmCfg<C, 128>::SMEM_BYTES);            \
    opt((const void*)gemm_grouped_kernel<C, TA, TB, 64>, GemmCfg<C, 64>::SMEM_BYTES);
    CB_OPT(false, false, false) CB_OPT(false, false, true) CB_OPT(false, true, false) CB_OPT(false, true, true)
    CB_OPT(true, false, false) CB_OPT(true, false, true) CB_OPT(true, true, false) CB_OPT(true, true, true)
#undef CB_OPT
    opt((const void*)gemm_grouped_peer_kernel<false>, GemmCfg<false, 128>::SMEM_BYTES);
    opt((const void*)gemm_grouped_peer_kernel<true>, GemmCfg<true, 128>::SMEM_BYTES);
    opt((const void*)gemm_grouped_peer_kernel<true, 64, true>, GemmCfg<true, 64>::SMEM_BYTES);
    opt((const void*)gemm_grouped_peer_kernel<true, 96, true>, GemmCfg<true, 96>::SMEM_BYTES);
    opt((const void*)gemm_grouped_kernel<true, false, false, 96, true>, GemmCfg<true, 96>::SMEM_BYTES);
    opt((const void*)gemm_grouped_kernel<true, false, true, 96, true>, GemmCfg<true, 96>::SMEM_BYTES);
    opt((const void*)gemm_grouped_kernel<true, true, false, 96, true>, GemmCfg<true, 96>::SMEM_BYTES);
    opt((const void*)gemm_grouped_kernel<true, true, true, 96, true>, GemmCfg<true, 96>::SMEM_BYTES);
    opt((const void*)gemm_grouped_kernel<true, false, false, 64, true>, GemmCfg<true, 64>::SMEM_BYTES);
    opt((const void*)gemm_grouped_kernel<true, false, true, 64, true>, GemmCfg<true, 64>::SMEM_BYTES);
    opt((const void*)gemm_grouped_kernel<true, true, false, 64, true>, GemmCfg<true, 64>::SMEM_BYTES);
    opt((const void*)gemm_grouped_kernel<true, true, true, 64, true>, GemmCfg<true, 64>::SMEM_BYTES);
  }
  {
    const char* e = getenv("CB200_GEMM_BULK");        // 0: per-thread cp.async tile loads for complex operands too (A/B switch)
    const int v = e ? atoi(e) : 1;
    cudaMemcpyToSymbol(g_gemm_bulk, &v, sizeof(int));
  }
  cudaGetLastError();
  return b;
}

void backend_destroy(Backend* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  cudaStreamSynchronize(b->stream);
  for (auto& ev : b->ev) cudaEventDestroy(ev);
  if (b->comm && b->p_ncclCommDestroy) b->p_ncclCommDestroy(b->comm);
  for (int r = 0; r < MAX_PEERS; ++r)
    if (b->peer_opened[r]) cudaIpcCloseMemHandle(b->peer_sym[r]);
  for (auto& kv : b->parked) cudaFreeAsync(kv.second, b->stream);
  b->parked.clear();
  cudaStreamSynchronize(b->stream);
  if (b->sym) cudaFree(b->sym);
  if (b->d_bounds) cudaFree(b->d_bounds);
  cudaFree(b->d_partial);
  cudaFree(b->d_dot_out);
  cudaFreeHost(b->h_dot_out);
  if (b->h_ring) cudaFreeHost(b->h_ring);
  cudaStreamDestroy(b->stream);
  delete b;
}

static constexpr size_t PARK_MIN = (size_t)32 << 20;   // smaller blocks are the pool's business (it recycles them well)

void* dev_alloc(Backend* b, size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  if (bytes >= PARK_MIN) {
    auto it = b->parked.lower_bound(bytes);           // smallest parked block that fits, if it wastes < 1/8
    if (it != b->parked.end() && it->first <= bytes + bytes / 8) {
      p = it->second;
      const size_t got = it->first;
      b->parked.erase(it);
      b->parked_bytes -= got;
      b->live_size[p] = got;
      return p;
    }
  }
  cudaError_t e = cudaMallocAsync(&p, bytes, b->stream);
  if (e != cudaSuccess && !b->parked.empty()) {       // out of memory with blocks parked: release them and retry once
    cudaGetLastError();
    for (auto& kv : b->parked) cudaFreeAsync(kv.second, b->stream);
    b->parked.clear();
    b->parked_bytes = 0;
    cudaStreamSynchronize(b->stream);
    e = cudaMallocAsync(&p, bytes, b->stream);
  }
  if (e != cudaSuccess) { set_err(b, "cudaMallocAsync", e); return nullptr; }
  if (bytes >= PARK_MIN) b->live_size[p] = bytes;
  return p;
}
void dev_free(Backend* b, void* p) {
  if (!p) return;
  auto it = b->live_size.find(p);
  if (it == b->live_size.end()) { CB_CHECK(b, cudaFreeAsync(p, b->stream)); return; }
  const size_t bytes = it->second;
  b->live_size.erase(it);
  b->parked.emplace(bytes, p);
  b->parked_bytes += bytes;
  while (b->parked_bytes > b->parked_cap && !b->parked.empty()) {   // over the cap: let the largest parked blocks go
    auto big = std::prev(b->parked.end());
    b->parked_bytes -= big->first;
    CB_CHECK(b, cudaFreeAsync(big->second, b->stream));
    b->parked.erase(big);
  }
}
void dev_memset0(Backend* b, void* p, size_t bytes) { if (bytes) CB_CHECK(b, cudaMemsetAsync(p, 0, bytes, b->stream)); }
void dev_h2d(Backend* b, void* dst, const void* src, size_t bytes) {
  if (!bytes) return;
  if (b->h_ring && bytes <= ((size_t)256 << 10)) {
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (b->ring_off + need > b->ring_bytes) {          // wrap: everything staged so far must have left the ring
      CB_CHECK(b, cudaStreamSynchronize(b->stream));
      b->ring_off = 0;
    }
    char* stage = b->h_ring + b->ring_off;
    memcpy(stage, src, bytes);
    b->ring_off += need;
    CB_CHECK(b, cudaMemcpyAsync(dst, stage, bytes, cudaMemcpyHostToDevice, b->stream));
    return;
  }
  CB_CHECK(b, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, b->stream));
  CB_CHECK(b, cudaStreamSynchronize(b->stream));   // source may be pageable / short-lived
}
void dev_d2h(Backend* b, void* dst, const void* src, size_t bytes) {
  if (!bytes) return;
  CB_CHECK(b, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, b->stream));
  CB_CHECK(b, cudaStreamSynchronize(b->stream));
}
void dev_d2d(Backend* b, void* dst, const void* src, size_t bytes) {
  if (bytes) CB_CHECK(b, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, b->stream));
}
void dev_sync(Backend* b) { CB_CHECK(b, cudaStreamSynchronize(b->stream)); }
const char* dev_last_error(Backend* b) { return b->has_err ? b->err.c_str() : nullptr; }
uint64_t dev_launch_count(Backend* b) { return b->launches; }
void dev_event_record(Backend* b, int slot) { CB_CHECK(b, cudaEventRecord(b->ev[slot & 7], b->stream)); }
float dev_event_elapsed_ms(Backend* b, int s0, int s1) {
  float ms = 0.f;
  CB_CHECK(b, cudaEventSynchronize(b->ev[s1 & 7]));
  CB_CHECK(b, cudaEventElapsedTime(&ms, b->ev[s0 & 7], b->ev[s1 & 7]));
  return ms;
}

// ------------------------------------------------------------------------------------------------ GEMM
int gemm_tile_m(bool cplx) { return cplx ? GemmCfg<true>::BM : GemmCfg<false>::BM; }
int gemm_tile_n(bool, int mode) { return mode == GEMM_TILE_WIDE ? 128 : (mode == GEMM_TILE_MID_3M ? 96 : 64); }

template <bool CPLX, int BN_, bool M3 = false>
static void launch_gemm_t(Backend* b, bool ta, bool tb, const void* A, const void* B, void* C, const GemmTask* t, int nt,
                          const GemmSeg* s, int tiles) {
  const int smem = GemmCfg<CPLX, BN_>::SMEM_BYTES;
  const char* a = (const char*)A;
  const char* bb = (const char*)B;
  char* c = (char*)C;
  if (!ta && !tb) gemm_grouped_kernel<CPLX, false, false, BN_, M3><<<tiles, GemmCfg<CPLX, BN_>::THREADS, smem, b->stream>>>(a, bb, c, t, nt, s);
  else if (!ta && tb) gemm_grouped_kernel<CPLX, false, true, BN_, M3><<<tiles, GemmCfg<CPLX, BN_>::THREADS, smem, b->stream>>>(a, bb, c, t, nt, s);
  else if (ta && !tb) gemm_grouped_kernel<CPLX, true, false, BN_, M3><<<tiles, GemmCfg<CPLX, BN_>::THREADS, smem, b->stream>>>(a, bb, c, t, nt, s);
  else gemm_grouped_kernel<CPLX, true, true, BN_, M3><<<tiles, GemmCfg<CPLX, BN_>::THREADS, smem, b->stream>>>(a, bb, c, t, nt, s);
}

void launch_gemm(Backend* b, bool cplx, bool ta, bool tb, int mode, const void* A, const void* B, void* C, const GemmTask* t,
                 int nt, const GemmSeg* s, int tiles) {
  if (tiles <= 0 || nt <= 0) return;
  if (cplx) {
    if (mode == GEMM_TILE_MID_3M) launch_gemm_t<true, 96, true>(b, ta, tb, A, B, C, t, nt, s, tiles);
    else if (mode == GEMM_TILE_NARROW_3M) launch_gemm_t<true, 64, true>(b, ta, tb, A, B, C, t, nt, s, tiles);
    else if (mode == GEMM_TILE_NARROW) launch_gemm_t<true, 64>(b, ta, tb, A, B, C, t, nt, s, tiles);
    else launch_gemm_t<true, 128>(b, ta, tb, A, B, C, t, nt, s, tiles);
  } else {
    if (mode == GEMM_TILE_NARROW) launch_gemm_t<false, 64>(b, ta, tb, A, B, C, t, nt, s, tiles);
    else if (mode == GEMM_TILE_WIDE) launch_gemm_t<false, 128>(b, ta, tb, A, B, C, t, nt, s, tiles);
    else set_err(b, "launch_gemm: the 3M tile modes are complex-only", cudaErrorInvalidValue);
  }
  CB_LAUNCH_CHECK(b, "gemm_grouped_kernel");
}

void launch_gemm_peer(Backend* b, bool cplx, int mode, const void* A, const void* B, const PeerBases& peers, const GemmTask* t, int nt,
                      const GemmSeg* s, int tiles) {
  if (tiles <= 0 || nt <= 0) return;
  if (cplx && mode == GEMM_TILE_MID_3M)
    gemm_grouped_peer_kernel<true, 96, true><<<tiles, GemmCfg<true, 96>::THREADS, GemmCfg<true, 96>::SMEM_BYTES, b->stream>>>((const char*)A, (const char*)B, t, nt, s, peers);
  else if (cplx && mode == GEMM_TILE_NARROW_3M)
    gemm_grouped_peer_kernel<true, 64, true><<<tiles, GemmCfg<true, 64>::THREADS, GemmCfg<true, 64>::SMEM_BYTES, b->stream>>>((const char*)A, (const char*)B, t, nt, s, peers);
  else if (mode != GEMM_TILE_WIDE) set_err(b, "launch_gemm_peer: unsupported tile mode", cudaErrorInvalidValue);
  else if (cplx) gemm_grouped_peer_kernel<true><<<tiles, GemmCfg<true, 128>::THREADS, GemmCfg<true, 128>::SMEM_BYTES, b->stream>>>((const char*)A, (const char*)B, t, nt, s, peers);
  else gemm_grouped_peer_kernel<false><<<tiles, GemmCfg<false, 128>::THREADS, GemmCfg<false, 128>::SMEM_BYTES, b->stream>>>((const char*)A, (const char*)B, t, nt, s, peers);
  CB_LAUNCH_CHECK(b, "gemm_grouped_peer_kernel");
}

// ------------------------------------------------------------------------------------------------ multi-GPU plumbing
int shard_alloc(Backend* b, int64_t landing_each, void* handle_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == SHARD_HANDLE_BYTES, "IPC handle size");
  if (b->sym) { set_err(b, "shard_alloc: already allocated", cudaErrorInvalidValue); return -1; }
  landing_each = (landing_each + 255) / 256 * 256;
  cudaError_t e = cudaMalloc(&b->sym, (size_t)(SYM_FLAG_BYTES + 2 * landing_each));
  if (e != cudaSuccess) { set_err(b, "shard_alloc: cudaMalloc", e); return -1; }
  b->landing_each = landing_each;
  CB_CHECK(b, cudaMemset(b->sym, 0, SYM_FLAG_BYTES));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, b->sym);
  if (e != cudaSuccess) { set_err(b, "shard_alloc: cudaIpcGetMemHandle", e); return -1; }
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

int shard_connect(Backend* b, int rank, int world, const void* all_handles) {
  if (!b->sym || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) {
    set_err(b, "shard_connect: bad arguments or no landing area", cudaErrorInvalidValue);
    return -1;
  }
  b->rank = rank; b->world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { b->peer_sym[r] = b->sym; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)all_handles + (size_t)r * SHARD_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_err(b, "shard_connect: cudaIpcOpenMemHandle", e); return -1; }
    b->peer_sym[r] = (char*)p;
    b->peer_opened[r] = true;
  }
  return 0;
}

int64_t shard_landing_bytes(Backend* b) { return b->sym ? b->landing_each : 0; }
void* shard_landing(Backend* b, int which) { return b->sym + SYM_FLAG_BYTES + (int64_t)(which & 1) * b->landing_each; }
void shard_peer_bases(Backend* b, int which, PeerBases* out) {
  out->n = b->world;
  for (int r = 0; r < MAX_PEERS; ++r) out->base[r] = r < b->world ? b->peer_sym[r] + SYM_FLAG_BYTES + (int64_t)(which & 1) * b->landing_each : nullptr;
}

struct PeerFlags { unsigned long long* f[MAX_PEERS]; };
// Thread t publishes this rank's arrival in rank t's flag row and waits for rank t's arrival in ours.  The release store
// follows a system-scope fence, so the peer stores of every kernel launched earlier on this stream (the R-stage tiles)
// are visible to whoever observes the flag; the acquire load orders our later kernels after the peers' data.
__global__ void shard_barrier_kernel(PeerFlags pf, int rank, int world, unsigned long long seq) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();
  unsigned long long* dst = pf.f[t] + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(seq) : "memory");
  const unsigned long long* src = pf.f[rank] + t;
  unsigned long long v = 0, t0 = 0, now = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (now - t0 > 60000000000ull) __trap();   // a peer died: fail this launch instead of hanging the box
  } while (v < seq);
}

void shard_barrier(Backend* b) {
  if (b->world <= 1) return;
  PeerFlags pf{};
  for (int r = 0; r < b->world; ++r) pf.f[r] = reinterpret_cast<unsigned long long*>(b->peer_sym[r]);
  shard_barrier_kernel<<<1, 32, 0, b->stream>>>(pf, b->rank, b->world, ++b->barrier_seq);
  CB_LAUNCH_CHECK(b, "shard_barrier_kernel");
}

static void nccl_fail(Backend* b, const char* what, ncclResult_t r);
// 16-byte grid-stride copy of one source buffer into every destination (the root's broadcast over NVLink)
__global__ void __launch_bounds__(256) peer_push_kernel(PeerBases dst, const uint4* __restrict__ src, long long n16) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n16; i += (long long)gridDim.x * 256) {
    const uint4 v = src[i];
    for (int r = 0; r < dst.n; ++r) reinterpret_cast<uint4*>(dst.base[r])[i] = v;
  }
}

void shard_bcast_p2p(Backend* b, void* buf, int64_t bytes, int root, int which) {
  if (b->world <= 1 || bytes <= 0) return;
  if (bytes > b->landing_each) { set_err(b, "shard_bcast_p2p: message larger than the landing area", cudaErrorInvalidValue); return; }
  const int64_t padded = (bytes + 15) / 16 * 16;     // buffers come from the pool allocator: at least 16-byte granular
  if (b->rank == root) {
    PeerBases pb;
    shard_peer_bases(b, which, &pb);
    const long long n16 = padded / 16;
    const int blocks = (int)std::min<long long>((n16 + 255) / 256, 148 * 8);
    peer_push_kernel<<<blocks, 256, 0, b->stream>>>(pb, (const uint4*)buf, n16);
    CB_LAUNCH_CHECK(b, "peer_push_kernel");
  }
  shard_barrier(b);
  if (b->rank != root) CB_CHECK(b, cudaMemcpyAsync(buf, shard_landing(b, which), (size_t)bytes, cudaMemcpyDeviceToDevice, b->stream));
}

// All-gather of a buffer whose slice [rank*chunk, min((rank+1)*chunk, total)) is valid on each rank: every rank pushes its slice into
// every landing half `which` (own included) over NVLink, one flag barrier, then the peers' slices are copied out of the own landing half.
void shard_allgather_p2p(Backend* b, void* buf, int64_t total_bytes, int64_t chunk_bytes, int which) {
  if (b->world <= 1 || total_bytes <= 0) return;
  if (total_bytes > b->landing_each || (chunk_bytes & 15)) { set_err(b, "shard_allgather_p2p: message larger than the landing area / chunk not 16-byte granular", cudaErrorInvalidValue); return; }
  const int64_t off = std::min<int64_t>((int64_t)b->rank * chunk_bytes, total_bytes);
  const int64_t mine = std::min<int64_t>(chunk_bytes, total_bytes - off);
  if (mine > 0) {
    PeerBases pb;
    shard_peer_bases(b, which, &pb);
    for (int r = 0; r < pb.n; ++r) pb.base[r] += off;
    const long long n16 = (mine + 15) / 16;
    const int blocks = (int)std::min<long long>((n16 + 255) / 256, 148 * 8);
    peer_push_kernel<<<blocks, 256, 0, b->stream>>>(pb, (const uint4*)((const char*)buf + off), n16);
    CB_LAUNCH_CHECK(b, "peer_push_kernel");
  }
  shard_barrier(b);
  const char* land = (const char*)shard_landing(b, which);
  if (off > 0) CB_CHECK(b, cudaMemcpyAsync(buf, land, (size_t)off, cudaMemcpyDeviceToDevice, b->stream));
  if (off + mine < total_bytes)
    CB_CHECK(b, cudaMemcpyAsync((char*)buf + off + mine, land + off + mine, (size_t)(total_bytes - off - mine), cudaMemcpyDeviceToDevice, b->stream));
}

void nccl_bcast(Backend* b, void* buf, int64_t bytes, int root) {
  if (!b->comm || !b->p_ncclBroadcast) { set_err(b, "nccl_bcast: no communicator", cudaErrorInvalidValue); return; }
  ncclResult_t r = b->p_ncclBroadcast(buf, buf, (size_t)bytes, ncclChar, root, b->comm, b->stream);
  if (r != ncclSuccess) nccl_fail(b, "ncclBroadcast", r);
  b->launches++;
}

static bool nccl_load(Backend* b) {
  if (b->nccl_lib) return true;
  b->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!b->nccl_lib) b->nccl_lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!b->nccl_lib) { set_err(b, "dlopen(libnccl.so.2) failed", cudaErrorUnknown); return false; }
#define CB_SYM(name) *(void**)(&b->p_##name) = dlsym(b->nccl_lib, #name)
  CB_SYM(ncclGetUniqueId); CB_SYM(ncclCommInitRank); CB_SYM(ncclAllReduce); CB_SYM(ncclBroadcast); CB_SYM(ncclCommDestroy); CB_SYM(ncclGetErrorString);
#undef CB_SYM
  if (!b->p_ncclGetUniqueId || !b->p_ncclCommInitRank || !b->p_ncclAllReduce) { set_err(b, "libnccl: missing symbols", cudaErrorUnknown); return false; }
  return true;
}
static void nccl_fail(Backend* b, const char* what, ncclResult_t r) {
  if (!b->has_err) {
    b->err = std::string(what) + ": " + (b->p_ncclGetErrorString ? b->p_ncclGetErrorString(r) : "nccl error");
    b->has_err = true;
  }
}
int nccl_unique_id(Backend* b, void* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (!nccl_load(b)) return -1;
  ncclUniqueId id;
  ncclResult_t r = b->p_ncclGetUniqueId(&id);
  if (r != ncclSuccess) { nccl_fail(b, "ncclGetUniqueId", r); return -1; }
  memcpy(out128, &id, sizeof(id));
  return 0;
}
int nccl_init(Backend* b, const void* uid128, int rank, int world) {
  if (!nccl_load(b)) return -1;
  ncclUniqueId id;
  memcpy(&id, uid128, sizeof(id));
  cudaSetDevice(b->device);
  ncclResult_t r = b->p_ncclCommInitRank(&b->comm, world, id, rank);
  if (r != ncclSuccess) { nccl_fail(b, "ncclCommInitRank", r); return -1; }
  b->rank = rank; b->world = world;
  return 0;
}
void nccl_allreduce_sum(Backend* b, void* buf, int64_t n_doubles) {
  if (!b->comm) { set_err(b, "nccl_allreduce_sum: no communicator", cudaErrorInvalidValue); return; }
  ncclResult_t r = b->p_ncclAllReduce(buf, buf, (size_t)n_doubles, ncclDouble, ncclSum, b->comm, b->stream);
  if (r != ncclSuccess) nccl_fail(b, "ncclAllReduce", r);
  b->launches++;
}

// ------------------------------------------------------------------------------------------------ mix
// out[row, oslot, r] = sum_e coef[e] * in[row, islot(e), r].  Three kernels:
//   mix_smem2_kernel: (default) a CTA stages ALL input slots of (row chunk, r) in shared memory once (coalesced, read exactly once from
//                    HBM) and forms every output slot from shared memory -> HBM traffic = read T1 + write T2.  The row chunk (32 or 16
//                    rows) is chosen so that at least TWO CTAs are resident per SM: one CTA's staging copies overlap another's output
//                    phase (the round-1 kernel, one 512-thread CTA per SM, ran load -> compute -> store back to back and reached 0.35 of
//                    the HBM copy rate); the descriptor of the next output slot is fetched while the current one is accumulated.
//   mix_smem_kernel : the round-1 form of the same kernel (CB200_MIX=v1).
//   mix_kernel      : generic fallback when the panel's inputs do not fit in shared memory (inputs re-read through L1/L2).
static constexpr int MIX_ROWS = 128;          // fallback kernel: rows per CTA
static constexpr int MIX_THREADS = 512;       // v1: 16 warps; each warp owns 32 rows x (every 16th output slot)
static constexpr int MIX_RB = 32;
static constexpr int MIX_SMEM_MAX = 208 * 1024;
static constexpr int MIX_CLUSTER_DEFAULT = 0;     // CTAs per cluster of mix_smem2_kernel (0 = plain launch); CB200_MIX_CLUSTER overrides
static constexpr int MIX_SMEM_SM = 227 * 1024;   // shared memory of one SM that CTAs can share (1 KB reserved per CTA on top of its request)

struct MixConfig {
  int version;      // 2 = mix_smem2_kernel, 1 = mix_smem_kernel, 0 = mix_kernel (generic fallback)
  int rb;           // rows per CTA
  int threads;
  size_t smem;
};
static MixConfig mix_config(bool cplx, int max_in) {
  // (read per call -- getenv is ~100 ns against a launch: tools/ab_mix.py switches variants inside one process)
  const char* ev = getenv("CB200_MIX");
  const bool v1 = ev && std::string(ev) == "v1";
  const char* er = getenv("CB200_MIX_RB");
  const int force_rb = er ? atoi(er) : 0;
  const char* et = getenv("CB200_MIX_THREADS");
  const int force_threads = et ? atoi(et) : 0;
  const size_t es = cplx ? 16 : 8;
  const size_t need32 = (size_t)max_in * 32 * es;
  if (need32 > (size_t)MIX_SMEM_MAX && (v1 || (size_t)max_in * 16 * es > (size_t)MIX_SMEM_MAX)) return {0, MIX_ROWS, MIX_ROWS, 0};
  if (v1) return {1, MIX_RB, MIX_THREADS, need32};
  // 32 rows when two CTAs of that size fit on an SM, else 16 rows (runs of 256 B for complex128, 128 B for real data)
  int rb = 2 * (need32 + 1024) <= (size_t)MIX_SMEM_SM ? 32 : 16;
  if ((force_rb == 16 || force_rb == 32) && (size_t)max_in * force_rb * es <= (size_t)MIX_SMEM_MAX) rb = force_rb;
  const size_t need = (size_t)max_in * rb * es;
  const int fit = (int)((size_t)MIX_SMEM_SM / (need + 1024));
  int threads = fit >= 4 ? 256 : 512;     // few resident CTAs: more warps per CTA keep the SM's load/store queues busy
  if (force_threads == 256 || force_threads == 512) threads = force_threads;
  return {2, rb, threads, need};
}
int mix_rows_per_block(bool cplx, int max_in) { return mix_config(cplx, max_in).rb; }

template <bool CPLX, int RB, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 4) mix_smem2_kernel(const void* __restrict__ in_, void* __restrict__ out_,
                                                                                     const MixPanel* __restrict__ panels, int npanels,
                                                                                     const MixIn* __restrict__ ins, const MixOut* __restrict__ outs,
                                                                                     const MixEntry* __restrict__ ents, int nblocks) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int G = THREADS / RB;              // groups of RB lanes; a group owns every G-th output slot
  extern __shared__ __align__(16) char msm[];
  T* sm = reinterpret_cast<T*>(msm);
  int lo = 0, hi = npanels - 1;
  const int blk = blockIdx.x;
  if (blk >= nblocks) return;                  // grid padded to a multiple of the cluster size (whole CTA leaves: no barrier is missed)
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (panels[mid].blk_begin <= blk) lo = mid; else hi = mid - 1;
  }
  const MixPanel p = panels[lo];
  const int local = blk - p.blk_begin;
  const int row0 = (local % p.row_chunks) * RB;
  const int64_t r = local / p.row_chunks;
  const int lrow = threadIdx.x % RB, grp = threadIdx.x / RB;
  const bool row_ok = row0 + lrow < p.rows;
  const T* in = reinterpret_cast<const T*>(in_);
  // stage every input slot of this (row chunk, r) once: fire-and-forget async copies (zero-filled past the last row)
  for (int s = grp; s < p.in_count; s += G) {
    const MixIn mi = ins[p.in_begin + s];
    const T* src = row_ok ? (in + mi.off + r * mi.rstride + row0 + lrow) : in;
    if constexpr (CPLX) cp_async16(sm + s * RB + lrow, src, row_ok ? 16 : 0);
    else cp_async8(sm + s * RB + lrow, src, row_ok ? 8 : 0);
  }
  cp_async_commit();
  // the first output descriptor travels while the copies are in flight
  int o = p.out_begin + grp;
  MixOut mo{};
  if (o < p.out_end) mo = outs[o];
  cp_async_wait<0>();
  __syncthreads();
  T* out = reinterpret_cast<T*>(out_);
  const T* smr = sm + lrow;
  while (o < p.out_end) {
    const int on = o + G;
    MixOut mn{};
    if (on < p.out_end) mn = outs[on];          // next descriptor: its latency hides behind this slot's arithmetic
    double ar = 0.0, ai = 0.0;
    int e = mo.e_begin;
    // entries are read as broadcast loads (every lane of the group the same address; L1-resident: all CTAs of a panel share them);
    // two entries per trip so that both pairs of loads are in flight together; the accumulation order is the entry order
    for (; e + 1 < mo.e_end; e += 2) {
      const double2 c0 = __ldg(reinterpret_cast<const double2*>(&ents[e].cre));
      const int s0 = __ldg(&ents[e].slot);
      const double2 c1 = __ldg(reinterpret_cast<const double2*>(&ents[e + 1].cre));
      const int s1 = __ldg(&ents[e + 1].slot);
      const T v0 = smr[s0 * RB];
      const T v1 = smr[s1 * RB];
      if constexpr (!CPLX) {
        ar = fma(c0.x, v0, ar);
        ar = fma(c1.x, v1, ar);
      } else {
        ar = fma(c0.x, v0.x, ar); ar = fma(-c0.y, v0.y, ar);
        ai = fma(c0.x, v0.y, ai); ai = fma(c0.y, v0.x, ai);
        ar = fma(c1.x, v1.x, ar); ar = fma(-c1.y, v1.y, ar);
        ai = fma(c1.x, v1.y, ai); ai = fma(c1.y, v1.x, ai);
      }
    }
    if (e < mo.e_end) {
      const double2 c0 = __ldg(reinterpret_cast<const double2*>(&ents[e].cre));
      const int s0 = __ldg(&ents[e].slot);
      const T v0 = smr[s0 * RB];
      if constexpr (!CPLX) {
        ar = fma(c0.x, v0, ar);
      } else {
        ar = fma(c0.x, v0.x, ar); ar = fma(-c0.y, v0.y, ar);
        ai = fma(c0.x, v0.y, ai); ai = fma(c0.y, v0.x, ai);
      }
    }
    if (row_ok) {
      if constexpr (!CPLX) out[mo.off + r * mo.rstride + row0 + lrow] = ar;
      else out[mo.off + r * mo.rstride + row0 + lrow] = make_double2(ar, ai);
    }
    mo = mn;
    o = on;
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(MIX_THREADS, 1) mix_smem_kernel(const void* __restrict__ in_, void* __restrict__ out_,
                                                                  const MixPanel* __restrict__ panels, int npanels,
                                                                  const MixIn* __restrict__ ins, const MixOut* __restrict__ outs,
                                                                  const MixEntry* __restrict__ ents) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int RB = MIX_RB;
  constexpr int G = MIX_THREADS / RB;
  extern __shared__ __align__(16) char msm[];
  T* sm = reinterpret_cast<T*>(msm);
  int lo = 0, hi = npanels - 1;
  const int blk = blockIdx.x;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (panels[mid].blk_begin <= blk) lo = mid; else hi = mid - 1;
  }
  const MixPanel p = panels[lo];
  const int local = blk - p.blk_begin;
  const int row0 = (local % p.row_chunks) * RB;
  const int64_t r = local / p.row_chunks;
  const int lrow = threadIdx.x % RB, grp = threadIdx.x / RB;
  const bool row_ok = row0 + lrow < p.rows;
  const T* in = reinterpret_cast<const T*>(in_);
  // stage every input slot of this (row chunk, r) once: fire-and-forget async copies (zero-filled past the last row)
  for (int s = grp; s < p.in_count; s += G) {
    const MixIn mi = ins[p.in_begin + s];
    const T* src = row_ok ? (in + mi.off + r * mi.rstride + row0 + lrow) : in;
    if constexpr (CPLX) cp_async16(sm + s * RB + lrow, src, row_ok ? 16 : 0);
    else cp_async8(sm + s * RB + lrow, src, row_ok ? 8 : 0);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  T* out = reinterpret_cast<T*>(out_);
  // One warp per output slot: the warp reads 32 entries with one coalesced load, parks them in a warp-private strip of shared memory and
  // then reads them back one by one as BROADCAST loads (every lane the same address: one wavefront each).  The inner loop is bound by the
  // LSU / shuffle unit, not by HBM: the round-1 form broadcast (cre, cim, slot) with five 32-bit shuffles per entry.
  const int lane = threadIdx.x & 31;
  __shared__ __align__(16) double2 s_coef[MIX_THREADS];     // [warp][32]
  __shared__ int s_slot[MIX_THREADS];
  double2* wc = s_coef + (threadIdx.x & ~31);
  int* ws = s_slot + (threadIdx.x & ~31);
  for (int o = p.out_begin + grp; o < p.out_end; o += G) {
    const MixOut mo = outs[o];
    double ar = 0.0, ai = 0.0;
    for (int base = mo.e_begin; base < mo.e_end; base += 32) {
      const int cnt = min(32, mo.e_end - base);
      __syncwarp();
      if (lane < cnt) {
        wc[lane] = __ldg(reinterpret_cast<const double2*>(&ents[base + lane].cre));
        ws[lane] = __ldg(&ents[base + lane].slot);
      }
      __syncwarp();
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const double2 c = wc[j];
        const int sl = ws[j];
        if constexpr (!CPLX) {
          ar = fma(c.x, sm[sl * RB + lrow], ar);
        } else {
          const double2 v = sm[sl * RB + lrow];
          ar = fma(c.x, v.x, ar); ar = fma(-c.y, v.y, ar);
          ai = fma(c.x, v.y, ai); ai = fma(c.y, v.x, ai);
        }
      }
    }
    if (row_ok) {
      if constexpr (!CPLX) out[mo.off + r * mo.rstride + row0 + lrow] = ar;
      else out[mo.off + r * mo.rstride + row0 + lrow] = make_double2(ar, ai);
    }
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(MIX_ROWS) mix_kernel(const void* __restrict__ in_, void* __restrict__ out_,
                                                       const MixPanel* __restrict__ panels, int npanels, const MixIn* __restrict__ ins,
                                                       const MixOut* __restrict__ outs, const MixEntry* __restrict__ ents) {
  int lo = 0, hi = npanels - 1;
  const int blk = blockIdx.x;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (panels[mid].blk_begin <= blk) lo = mid; else hi = mid - 1;
  }
  const MixPanel p = panels[lo];
  const int local = blk - p.blk_begin;
  const int row = (local % p.row_chunks) * MIX_ROWS + threadIdx.x;
  const int64_t r = local / p.row_chunks;
  if (row >= p.rows) return;
  for (int o = p.out_begin; o < p.out_end; ++o) {
    const MixOut mo = outs[o];
    if constexpr (!CPLX) {
      const double* in = (const double*)in_;
      double acc = 0.0;
      for (int e = mo.e_begin; e < mo.e_end; ++e) {
        const MixEntry me = ents[e];
        const MixIn mi = ins[p.in_begin + me.slot];
        acc = fma(me.cre, __ldg(in + mi.off + r * mi.rstride + row), acc);
      }
      ((double*)out_)[mo.off + r * mo.rstride + row] = acc;
    } else {
      const double2* in = (const double2*)in_;
      double ar = 0.0, ai = 0.0;
      for (int e = mo.e_begin; e < mo.e_end; ++e) {
        const MixEntry me = ents[e];
        const MixIn mi = ins[p.in_begin + me.slot];
        const double2 v = __ldg(in + mi.off + r * mi.rstride + row);
        ar = fma(me.cre, v.x, ar); ar = fma(-me.cim, v.y, ar);
        ai = fma(me.cre, v.y, ai); ai = fma(me.cim, v.x, ai);
      }
      ((double2*)out_)[mo.off + r * mo.rstride + row] = make_double2(ar, ai);
    }
  }
}

template <bool CPLX, int RB, int THREADS>
static void launch_mix2(Backend* b, const void* in, void* out, const MixPanel* p, int np, const MixIn* mi, const MixOut* o, const MixEntry* e,
                        int blocks, size_t smem) {
  static bool attr_done = false;     // one device per process (backend_create enforces it)
  if (!attr_done) {
    cudaFuncSetAttribute(mix_smem2_kernel<CPLX, RB, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_MAX);
    attr_done = true;
  }
  // ask for a shared-memory carve-out that holds every CTA the register budget admits (2 x 512 or 4 x 256 threads): the overlap between
  // CTAs is the point of this kernel; the rest of the 256 KB stays L1 for the descriptor / entry lists every CTA of a panel re-reads
  static int carve_set = -1;
  const int resident = std::min<int>((int)((size_t)MIX_SMEM_SM / (smem + 1024)), THREADS == 512 ? 2 : 4);
  const int carve = std::min<int>(100, (int)(((size_t)std::max(resident, 1) * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)));
  if (carve != carve_set) {
    cudaFuncSetAttribute(mix_smem2_kernel<CPLX, RB, THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    carve_set = carve;
  }
  // CB200_MIX_CLUSTER=c (2, 4, 8): consecutive CTAs -- adjacent row chunks of the same (slot, r) runs -- are co-scheduled as one thread-block
  // cluster, so that their requests for neighbouring 256-byte pieces reach the memory system together.  No distributed shared memory is used.
  const char* ec = getenv("CB200_MIX_CLUSTER");
  const int cl = ec ? atoi(ec) : MIX_CLUSTER_DEFAULT;
  if (cl == 2 || cl == 4 || cl == 8) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)((blocks + cl - 1) / cl * cl));
    lc.blockDim = dim3(THREADS);
    lc.dynamicSmemBytes = smem;
    lc.stream = b->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    if (cudaLaunchKernelEx(&lc, mix_smem2_kernel<CPLX, RB, THREADS>, in, out, p, np, mi, o, e, blocks) == cudaSuccess) return;
    (void)cudaGetLastError();     // a cluster shape the device cannot place: plain launch below
  }
  mix_smem2_kernel<CPLX, RB, THREADS><<<blocks, THREADS, smem, b->stream>>>(in, out, p, np, mi, o, e, blocks);
}

void launch_mix(Backend* b, bool cplx, const void* in, void* out, const MixPanel* p, int np, const MixIn* mi, const MixOut* o,
                const MixEntry* e, int blocks, int max_in) {
  if (blocks <= 0) return;
  const MixConfig cfg = mix_config(cplx, max_in);
  if (cfg.version == 2) {
#define CB_MIX2(C, R, T) launch_mix2<C, R, T>(b, in, out, p, np, mi, o, e, blocks, cfg.smem)
    if (cplx) {
      if (cfg.rb == 32) { if (cfg.threads == 512) CB_MIX2(true, 32, 512); else CB_MIX2(true, 32, 256); }
      else { if (cfg.threads == 512) CB_MIX2(true, 16, 512); else CB_MIX2(true, 16, 256); }
    } else {
      if (cfg.rb == 32) { if (cfg.threads == 512) CB_MIX2(false, 32, 512); else CB_MIX2(false, 32, 256); }
      else { if (cfg.threads == 512) CB_MIX2(false, 16, 512); else CB_MIX2(false, 16, 256); }
    }
#undef CB_MIX2
    CB_LAUNCH_CHECK(b, "mix_smem2_kernel");
  } else if (cfg.version == 1) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(mix_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_MAX);
      cudaFuncSetAttribute(mix_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_MAX);
      attr_done = true;
    }
    if (cplx) mix_smem_kernel<true><<<blocks, MIX_THREADS, cfg.smem, b->stream>>>(in, out, p, np, mi, o, e);
    else mix_smem_kernel<false><<<blocks, MIX_THREADS, cfg.smem, b->stream>>>(in, out, p, np, mi, o, e);
    CB_LAUNCH_CHECK(b, "mix_smem_kernel");
  } else {
    if (cplx) mix_kernel<true><<<blocks, MIX_ROWS, 0, b->stream>>>(in, out, p, np, mi, o, e);
    else mix_kernel<false><<<blocks, MIX_ROWS, 0, b->stream>>>(in, out, p, np, mi, o, e);
    CB_LAUNCH_CHECK(b, "mix_kernel");
  }
}

// ------------------------------------------------------------------------------------------------ copy
static constexpr int COPY_THREADS = 256;
static constexpr int COPY_PER_BLOCK = 2048;
int copy_elems_per_block() { return COPY_PER_BLOCK; }

template <bool CPLX>
__global__ void __launch_bounds__(COPY_THREADS) copy_kernel(const void* __restrict__ src_, void* __restrict__ dst_,
                                                            const CopyTask* __restrict__ tasks, int ntasks) {
  int lo = 0, hi = ntasks - 1;
  const int blk = blockIdx.x;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (tasks[mid].blk_begin <= blk) lo = mid; else hi = mid - 1;
  }
  const CopyTask t = tasks[lo];
  const int64_t total = (int64_t)t.n0 * t.n1 * t.n2;
  const int64_t base = (int64_t)(blk - t.blk_begin) * COPY_PER_BLOCK;
  for (int u = threadIdx.x; u < COPY_PER_BLOCK; u += COPY_THREADS) {
    const int64_t idx = base + u;
    if (idx >= total) break;
    const int64_t i = idx % t.n0;
    const int64_t jk = idx / t.n0;
    const int64_t j = jk % t.n1, k = jk / t.n1;
    const int64_t so = t.src_off + i * t.ss0 + j * t.ss1 + k * t.ss2;
    const int64_t d_o = t.dst_off + i * t.ds0 + j * t.ds1 + k * t.ds2;
    if constexpr (!CPLX) {
      ((double*)dst_)[d_o] = ((const double*)src_)[so];
    } else {
      double2 v = ((const double2*)src_)[so];
      if (t.conj) v.y = -v.y;
      ((double2*)dst_)[d_o] = v;
    }
  }
}

void launch_copy(Backend* b, bool cplx, const void* src, void* dst, const CopyTask* t, int nt, int blocks) {
  if (blocks <= 0) return;
  if (cplx) copy_kernel<true><<<blocks, COPY_THREADS, 0, b->stream>>>(src, dst, t, nt);
  else copy_kernel<false><<<blocks, COPY_THREADS, 0, b->stream>>>(src, dst, t, nt);
  CB_LAUNCH_CHECK(b, "copy_kernel");
}

// ------------------------------------------------------------------------------------------------ level 1
struct PtrPack { const void* p[MAX_LC]; };
struct CoefPack { double re[MAX_LC], im[MAX_LC]; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// partial[(2*j+c)*DOT_BLOCKS + blk] = block-local sum; fixed traversal order => bitwise reproducible
template <bool CPLX, int NV>
__global__ void __launch_bounds__(DOT_THREADS) dot_partial_kernel(PtrPack xs, const void* __restrict__ y_, int64_t n,
                                                                  double* __restrict__ partial) {
  double sr[NV], si[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) { sr[j] = 0.0; si[j] = 0.0; }
  const int64_t stride = (int64_t)gridDim.x * DOT_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * DOT_THREADS + threadIdx.x; i < n; i += stride) {
    if constexpr (!CPLX) {
      const double yv = ((const double*)y_)[i];
#pragma unroll
      for (int j = 0; j < NV; ++j) sr[j] = fma(((const double*)xs.p[j])[i], yv, sr[j]);
    } else {
      const double2 yv = ((const double2*)y_)[i];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const double2 xv = ((const double2*)xs.p[j])[i];
        sr[j] = fma(xv.x, yv.x, sr[j]); sr[j] = fma(xv.y, yv.y, sr[j]);
        si[j] = fma(xv.x, yv.y, si[j]); si[j] = fma(-xv.y, yv.x, si[j]);
      }
    }
  }
  __shared__ double red[2 * MAX_LC][DOT_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double a = warp_sum(sr[j]);
    double c = warp_sum(si[j]);
    if (lane == 0) { red[2 * j][w] = a; red[2 * j + 1][w] = c; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * NV) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < DOT_THREADS / 32; ++k) s += red[threadIdx.x][k];
    partial[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
}

__global__ void dot_final_kernel(const double* __restrict__ partial, int nblocks, int nout, double* __restrict__ out) {
  // one warp per output, fixed-order tree => deterministic
  const int o = blockIdx.x;
  if (o >= nout) return;
  double s = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) s += partial[(int64_t)o * nblocks + i];
  s = warp_sum(s);
  if (threadIdx.x == 0) out[o] = s;
}

template <bool CPLX>
static void dot_dispatch(Backend* b, int nv, const PtrPack& xs, const void* y, int64_t n, int blocks) {
  switch (nv) {
#define CASE(N) case N: dot_partial_kernel<CPLX, N><<<blocks, DOT_THREADS, 0, b->stream>>>(xs, y, n, b->d_partial); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
}

void multi_dot(Backend* b, bool cplx, int nv, const void* const* x, const void* y, int64_t n, double* out) {
  if (nv <= 0) return;
  if (nv > MAX_LC) { set_err(b, "multi_dot: too many vectors", cudaErrorInvalidValue); return; }
  PtrPack xs{};
  for (int j = 0; j < nv; ++j) xs.p[j] = x[j];
  int blocks = (int)((n + DOT_THREADS * 4 - 1) / (DOT_THREADS * 4));
  if (blocks > DOT_BLOCKS) blocks = DOT_BLOCKS;
  if (blocks < 1) blocks = 1;
  if (cplx) dot_dispatch<true>(b, nv, xs, y, n, blocks); else dot_dispatch<false>(b, nv, xs, y, n, blocks);
  CB_LAUNCH_CHECK(b, "dot_partial_kernel");
  dot_final_kernel<<<2 * nv, 32, 0, b->stream>>>(b->d_partial, blocks, 2 * nv, b->d_dot_out);
  CB_LAUNCH_CHECK(b, "dot_final_kernel");
  CB_CHECK(b, cudaMemcpyAsync(b->h_dot_out, b->d_dot_out, sizeof(double) * 2 * nv, cudaMemcpyDeviceToHost, b->stream));
  CB_CHECK(b, cudaStreamSynchronize(b->stream));
  for (int j = 0; j < 2 * nv; ++j) out[j] = b->h_dot_out[j];
  if (!cplx) for (int j = 0; j < nv; ++j) out[2 * j + 1] = 0.0;
}

template <bool CPLX, int NV>
__global__ void __launch_bounds__(256) lincomb_kernel(PtrPack xs, CoefPack c, double br, double bi, void* __restrict__ dst_, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool has_beta = (br != 0.0) || (bi != 0.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if constexpr (!CPLX) {
      double acc = has_beta ? br * ((double*)dst_)[i] : 0.0;
#pragma unroll
      for (int j = 0; j < NV; ++j) acc = fma(c.re[j], ((const double*)xs.p[j])[i], acc);
      ((double*)dst_)[i] = acc;
    } else {
      double ar = 0.0, ai = 0.0;
      if (has_beta) {
        const double2 d = ((double2*)dst_)[i];
        ar = br * d.x - bi * d.y; ai = br * d.y + bi * d.x;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const double2 v = ((const double2*)xs.p[j])[i];
        ar = fma(c.re[j], v.x, ar); ar = fma(-c.im[j], v.y, ar);
        ai = fma(c.re[j], v.y, ai); ai = fma(c.im[j], v.x, ai);
      }
      ((double2*)dst_)[i] = make_double2(ar, ai);
    }
  }
}

void lincomb(Backend* b, bool cplx, int nv, const void* const* x, const double* c, double br, double bi, void* dst, int64_t n) {
  if (n <= 0) return;
  if (nv > MAX_LC || nv < 1) { set_err(b, "lincomb: bad vector count", cudaErrorInvalidValue); return; }
  PtrPack xs{};
  CoefPack cp{};
  for (int j = 0; j < nv; ++j) { xs.p[j] = x[j]; cp.re[j] = c[2 * j]; cp.im[j] = c[2 * j + 1]; }
  int blocks = (int)((n + 256 * 4 - 1) / (256 * 4));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
#define CASE(N) case N: if (cplx) lincomb_kernel<true, N><<<blocks, 256, 0, b->stream>>>(xs, cp, br, bi, dst, n); \
                        else lincomb_kernel<false, N><<<blocks, 256, 0, b->stream>>>(xs, cp, br, bi, dst, n); break;
  switch (nv) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) }
#undef CASE
  CB_LAUNCH_CHECK(b, "lincomb_kernel");
}

// ------------------------------------------------------------------------------------------------ Jacobi
// Cyclic two-sided Jacobi on an NJ x NJ Hermitian matrix held in shared memory, round-robin parallel ordering
// (NJ/2 disjoint rotations per step).  Used as the inner solver of the block one-sided Jacobi SVD that replaces
// ITensor's diagHermitian(rho) (LAPACK dsyev/zheev) in MPS::svdBond.
static constexpr int NJ = 2 * JB;
static constexpr int LDJ = NJ + 1;
static constexpr int JAC_THREADS = 256;
static constexpr int JAC_MAX_SWEEPS = 60;

template <bool CPLX>
__global__ void __launch_bounds__(JAC_THREADS) jacobi_eig_kernel(const void* __restrict__ G_, void* __restrict__ Q_, double tol, double floor2, double abs_g2,
                                                                 int max_sweeps, unsigned long long* __restrict__ rot_total) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  extern __shared__ __align__(16) char jsm[];
  T* G = reinterpret_cast<T*>(jsm);
  T* Q = G + NJ * LDJ;
  __shared__ double rc[NJ / 2], rs[NJ / 2], rer[NJ / 2], rei[NJ / 2];
  __shared__ int rp[NJ / 2], rq[NJ / 2];
  __shared__ int sweep_rot;
  __shared__ int any_big;
  const int tid = threadIdx.x;
  if (tid == 0) any_big = 0;
  const T* Gg = reinterpret_cast<const T*>(G_) + (size_t)blockIdx.x * NJ * NJ;
  T* Qg = reinterpret_cast<T*>(Q_) + (size_t)blockIdx.x * NJ * NJ;

  // load the upper triangle and mirror it so that G is exactly Hermitian
  for (int idx = tid; idx < NJ * NJ; idx += JAC_THREADS) {
    const int r = idx % NJ, c = idx / NJ;
    T v;
    if (r <= c) v = Gg[r + c * NJ];
    else {
      v = Gg[c + r * NJ];
      if constexpr (CPLX) v.y = -v.y;
    }
    if constexpr (CPLX) { if (r == c) v.y = 0.0; }
    G[r + c * LDJ] = v;
    T one;
    if constexpr (CPLX) one = make_double2(r == c ? 1.0 : 0.0, 0.0); else one = (r == c ? 1.0 : 0.0);
    Q[r + c * LDJ] = one;
  }
  __syncthreads();

  const double tol2 = tol * tol;
  // a pair that is already orthogonal to working precision (the common case in the last outer sweeps) costs one scan, not a sweep
  for (int idx = tid; idx < NJ * NJ; idx += JAC_THREADS) {
    const int r = idx % NJ, c = idx / NJ;
    if (r >= c) continue;
    double gpp, gqq, g2;
    if constexpr (CPLX) { gpp = G[r + r * LDJ].x; gqq = G[c + c * LDJ].x; const T v = G[r + c * LDJ]; g2 = v.x * v.x + v.y * v.y; }
    else { gpp = G[r + r * LDJ]; gqq = G[c + c * LDJ]; const T v = G[r + c * LDJ]; g2 = v * v; }
    if (g2 > tol2 * fabs(gpp * gqq) && g2 > abs_g2 && gpp > floor2 && gqq > floor2) any_big = 1;
  }
  __syncthreads();
  if (!any_big) max_sweeps = 0;
  unsigned long long my_total = 0;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (tid == 0) sweep_rot = 0;
    __syncthreads();
    for (int step = 0; step < NJ - 1; ++step) {
      // ---- rotation parameters for the NJ/2 disjoint pairs of this step
      if (tid < NJ / 2) {
        int p, q;
        if (tid == 0) { p = step; q = NJ - 1; }
        else { p = (step + tid) % (NJ - 1); q = (step - tid + (NJ - 1)) % (NJ - 1); }
        if (p > q) { int t_ = p; p = q; q = t_; }
        double gpp, gqq, gr, gi;
        if constexpr (CPLX) { gpp = G[p + p * LDJ].x; gqq = G[q + q * LDJ].x; gr = G[p + q * LDJ].x; gi = G[p + q * LDJ].y; }
        else { gpp = G[p + p * LDJ]; gqq = G[q + q * LDJ]; gr = G[p + q * LDJ]; gi = 0.0; }
        const double g2 = gr * gr + gi * gi;
        double c = 1.0, s = 0.0, er = 1.0, ei = 0.0;
        if (g2 > tol2 * fabs(gpp * gqq) && g2 > abs_g2 && gpp > floor2 && gqq > floor2) {
          const double ag = sqrt(g2);
          er = gr / ag; ei = gi / ag;
          const double tau = (gqq - gpp) / (2.0 * ag);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t);
          s = t * c;
          atomicAdd(&sweep_rot, 1);
        }
        rp[tid] = p; rq[tid] = q; rc[tid] = c; rs[tid] = s; rer[tid] = er; rei[tid] = ei;
      }
      __syncthreads();
      // ---- column phase: G <- G J, Q <- Q J   (J[p,p]=J[q,q]=c, J[p,q]=s e^{i th}, J[q,p]=-s e^{-i th})
      for (int idx = tid; idx < (NJ / 2) * NJ; idx += JAC_THREADS) {
        const int k = idx / NJ, r = idx % NJ;
        const double s = rs[k];
        if (s == 0.0) continue;
        const double c = rc[k], er = rer[k], ei = rei[k];
        const int p = rp[k], q = rq[k];
        if constexpr (!CPLX) {
          const double sg = s * er;   // er = +-1
          double a = G[r + p * LDJ], b = G[r + q * LDJ];
          G[r + p * LDJ] = c * a - sg * b; G[r + q * LDJ] = sg * a + c * b;
          a = Q[r + p * LDJ]; b = Q[r + q * LDJ];
          Q[r + p * LDJ] = c * a - sg * b; Q[r + q * LDJ] = sg * a + c * b;
        } else {
          // new_p = c a - s e^{-i th} b ; new_q = s e^{i th} a + c b
          const double sr = s * er, si = s * ei;
          double2 a = G[r + p * LDJ], b = G[r + q * LDJ];
          G[r + p * LDJ] = make_double2(c * a.x - (sr * b.x + si * b.y), c * a.y - (sr * b.y - si * b.x));
          G[r + q * LDJ] = make_double2((sr * a.x - si * a.y) + c * b.x, (sr * a.y + si * a.x) + c * b.y);
          a = Q[r + p * LDJ]; b = Q[r + q * LDJ];
          Q[r + p * LDJ] = make_double2(c * a.x - (sr * b.x + si * b.y), c * a.y - (sr * b.y - si * b.x));
          Q[r + q * LDJ] = make_double2((sr * a.x - si * a.y) + c * b.x, (sr * a.y + si * a.x) + c * b.y);
        }
      }
      __syncthreads();
      // ---- row phase: G <- J^H G   (row_p' = c row_p - s e^{i th} row_q ; row_q' = s e^{-i th} row_p + c row_q)
      for (int idx = tid; idx < (NJ / 2) * NJ; idx += JAC_THREADS) {
        const int k = idx / NJ, cc = idx % NJ;
        const double s = rs[k];
        if (s == 0.0) continue;
        const double c = rc[k], er = rer[k], ei = rei[k];
        const int p = rp[k], q = rq[k];
        if constexpr (!CPLX) {
          const double sg = s * er;
          const double a = G[p + cc * LDJ], b = G[q + cc * LDJ];
          G[p + cc * LDJ] = c * a - sg * b; G[q + cc * LDJ] = sg * a + c * b;
        } else {
          const double sr = s * er, si = s * ei;
          const double2 a = G[p + cc * LDJ], b = G[q + cc * LDJ];
          G[p + cc * LDJ] = make_double2(c * a.x - (sr * b.x - si * b.y), c * a.y - (sr * b.y + si * b.x));
          G[q + cc * LDJ] = make_double2((sr * a.x + si * a.y) + c * b.x, (sr * a.y - si * a.x) + c * b.y);
        }
      }
      __syncthreads();
      // ---- clean the annihilated entries
      if (tid < NJ / 2 && rs[tid] != 0.0) {
        const int p = rp[tid], q = rq[tid];
        if constexpr (CPLX) {
          G[p + q * LDJ] = make_double2(0.0, 0.0); G[q + p * LDJ] = make_double2(0.0, 0.0);
          G[p + p * LDJ].y = 0.0; G[q + q * LDJ].y = 0.0;
        } else { G[p + q * LDJ] = 0.0; G[q + p * LDJ] = 0.0; }
      }
      __syncthreads();
    }
    const int nrot = sweep_rot;
    my_total += (unsigned long long)nrot;
    __syncthreads();
    if (nrot == 0) break;
  }
  for (int idx = tid; idx < NJ * NJ; idx += JAC_THREADS) {
    const int r = idx % NJ, c = idx / NJ;
    Qg[r + c * NJ] = Q[r + c * LDJ];
  }
  if (tid == 0 && my_total) atomicAdd(rot_total, my_total);
}

void launch_jacobi_eig(Backend* b, bool cplx, const void* G, void* Q, int npairs, double tol, double floor2, double abs_g2,
                       int max_sweeps, long long* d_rot) {
  if (max_sweeps <= 0 || max_sweeps > JAC_MAX_SWEEPS) max_sweeps = JAC_MAX_SWEEPS;
  if (npairs <= 0) return;
  const size_t smem = (size_t)2 * NJ * LDJ * (cplx ? 16 : 8);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(jacobi_eig_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * NJ * LDJ * 16);
    cudaFuncSetAttribute(jacobi_eig_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * NJ * LDJ * 8);
    attr_done = true;
  }
  if (cplx) jacobi_eig_kernel<true><<<npairs, JAC_THREADS, smem, b->stream>>>(G, Q, tol, floor2, abs_g2, max_sweeps, (unsigned long long*)d_rot);
  else jacobi_eig_kernel<false><<<npairs, JAC_THREADS, smem, b->stream>>>(G, Q, tol, floor2, abs_g2, max_sweeps, (unsigned long long*)d_rot);
  CB_LAUNCH_CHECK(b, "jacobi_eig_kernel");
}

template <bool CPLX>
__global__ void __launch_bounds__(256) colnorm2_kernel(const void* __restrict__ X_, int64_t ld, int64_t rows, double* __restrict__ out) {
  const int64_t col = blockIdx.x;
  double s = 0.0;
  if constexpr (!CPLX) {
    const double* x = (const double*)X_ + col * ld;
    for (int64_t i = threadIdx.x; i < rows; i += 256) s = fma(x[i], x[i], s);
  } else {
    const double2* x = (const double2*)X_ + col * ld;
    for (int64_t i = threadIdx.x; i < rows; i += 256) { const double2 v = x[i]; s = fma(v.x, v.x, s); s = fma(v.y, v.y, s); }
  }
  __shared__ double red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    out[col] = t;
  }
}

void launch_colnorm2(Backend* b, bool cplx, const void* X, int64_t ld, int64_t rows, int64_t cols, double* d_out) {
  if (cols <= 0) return;
  if (cplx) colnorm2_kernel<true><<<(unsigned)cols, 256, 0, b->stream>>>(X, ld, rows, d_out);
  else colnorm2_kernel<false><<<(unsigned)cols, 256, 0, b->stream>>>(X, ld, rows, d_out);
  CB_LAUNCH_CHECK(b, "colnorm2_kernel");
}

template <bool CPLX>
__global__ void set_identity_kernel(void* X_, int64_t ld, int64_t n) {
  const int64_t total = n * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx % n, c = idx / n;
    if constexpr (CPLX) ((double2*)X_)[r + c * ld] = make_double2(r == c ? 1.0 : 0.0, 0.0);
    else ((double*)X_)[r + c * ld] = (r == c ? 1.0 : 0.0);
  }
}

void launch_set_identity(Backend* b, bool cplx, void* X, int64_t ld, int64_t n) {
  if (n <= 0) return;
  int64_t blocks = (n * n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (cplx) set_identity_kernel<true><<<(unsigned)blocks, 256, 0, b->stream>>>(X, ld, n);
  else set_identity_kernel<false><<<(unsigned)blocks, 256, 0, b->stream>>>(X, ld, n);
  CB_LAUNCH_CHECK(b, "set_identity_kernel");
}

// ------------------------------------------------------------------------------------------------ Hermitian eigensolver
static bool tridiag_unblocked(Backend* b, bool cplx, void* A, int64_t n, double* d, double* e, void* tau, void* scl, void* p) {
  static int grid_blocks[2] = {0, 0};
  const void* fn = cplx ? (const void*)tridiag_kernel<true> : (const void*)tridiag_kernel<false>;
  if (!grid_blocks[cplx]) {
    int per_sm = 0, sms = 0;
    CB_CHECK(b, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, TRI_THREADS, 0));
    CB_CHECK(b, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
    grid_blocks[cplx] = std::max(1, per_sm) * std::max(1, sms);
  }
  // no more CTAs than there are columns to own (one warp per column): fewer CTAs = cheaper grid barriers on small problems
  int blocks = (int)std::min<int64_t>(grid_blocks[cplx], std::max<int64_t>(1, (n + TRI_WARPS - 1) / TRI_WARPS));
  long long nn = n;
  void* args[] = {&A, &nn, &d, &e, &tau, &scl, &p};
  CB_CHECK(b, cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(TRI_THREADS), args, 0, b->stream));
  CB_LAUNCH_CHECK(b, "tridiag_kernel");
  return true;
}

bool launch_tridiag(Backend* b, bool cplx, void* A, int64_t n, double* d, double* e, void* tau, void* scl, void* p) {
  if (n <= 0) return true;
  static const bool unblocked = [] { const char* v = getenv("CB200_TRIDIAG"); return v && std::string(v) == "unblocked"; }();
  if (unblocked || n < 3) return tridiag_unblocked(b, cplx, A, n, d, e, tau, scl, p);
  // ---- blocked: one cooperative launch per panel of <= 64 reflectors + one GEMM launch for the deferred rank-2nb update
  const int NB = 64;
  const size_t es = cplx ? 16 : 8;
  static int grid_blocks[2] = {0, 0};
  const void* fn = cplx ? (const void*)tridiag_panel_kernel<true> : (const void*)tridiag_panel_kernel<false>;
  if (!grid_blocks[cplx]) {
    int per_sm = 0, sms = 0;
    CB_CHECK(b, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, TRI_THREADS, 0));
    CB_CHECK(b, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
    const char* v = getenv("CB200_TRI_CTAS_PER_SM");
    if (v && atoi(v) > 0) per_sm = std::min(per_sm, atoi(v));
    grid_blocks[cplx] = std::max(1, per_sm) * std::max(1, sms);
  }
  int blocks = (int)std::min<int64_t>(grid_blocks[cplx], std::max<int64_t>(1, (n + TRI_WARPS - 1) / TRI_WARPS));
  {
    static const int cap = [] { const char* v = getenv("CB200_TRI_MAX_CTAS"); return v ? atoi(v) : 0; }();
    if (cap > 0) blocks = std::min(blocks, cap);
  }
  const int64_t nref = n - 1;
  const int npanel = (int)((nref + NB - 1) / NB);
  char* pan = (char*)dev_alloc(b, 6 * es * (size_t)n * NB);          // Vp | W | Vn | Wn (n x NB, ld n) | Vt | Wt (NB x n, ld NB)
  char* gbuf = (char*)dev_alloc(b, es * 128);
  double* partial = (double*)dev_alloc(b, sizeof(double) * (size_t)blocks);
  GemmTask* d_tasks = (GemmTask*)dev_alloc(b, sizeof(GemmTask) * (size_t)npanel);
  GemmSeg* d_segs = (GemmSeg*)dev_alloc(b, sizeof(GemmSeg) * 2 * (size_t)npanel);
  if (!pan || !gbuf || !partial || !d_tasks || !d_segs) return true;   // the allocation error is already recorded
  void* Vp = pan; void* W = pan + es * (size_t)n * NB; void* Vn = pan + 2 * es * (size_t)n * NB; void* Wn = pan + 3 * es * (size_t)n * NB;
  void* Vt = pan + 4 * es * (size_t)n * NB; void* Wt = pan + 5 * es * (size_t)n * NB;
  // task list of the rank-2nb updates: A[t0:, t0:] += conj( conj(Vn) W^T + conj(Wn) Vp^T ) = -(V W^H + W V^H)
  std::vector<GemmTask> tasks(npanel);
  std::vector<GemmSeg> segs(2 * (size_t)npanel);
  std::vector<int> tiles(npanel, 0);
  const int bm = gemm_tile_m(cplx), bn = gemm_tile_n(cplx, GEMM_TILE_WIDE);
  for (int pi = 0; pi < npanel; ++pi) {
    const int64_t j0 = (int64_t)pi * NB;
    const int nbk = (int)std::min<int64_t>(NB, nref - j0);
    const int64_t t0 = j0 + nbk, mt = n - t0;
    GemmTask t{};
    t.c_off = t0 + t0 * n; t.c_off2 = t.c_off; t.ldc = n;
    t.M = (int32_t)mt; t.N = (int32_t)mt; t.n_split = (int32_t)mt;
    t.seg_begin = 2 * pi; t.seg_count = 2;
    t.tile_begin = 0; t.tiles_m = (int32_t)((mt + bm - 1) / bm);
    t.flags = GEMM_ACCUM | (cplx ? (GEMM_CONJ_A | GEMM_CONJ_OUT) : 0u);
    tiles[pi] = t.tiles_m * (int)((mt + bn - 1) / bn);
    tasks[pi] = t;
    GemmSeg s1{}, s2{};
    // A-operands are addressed from Vn (the panel base handed to the launch); Wn sits 1 panel further, W / Vp are the B-operands
    s1.a_off = t0; s1.lda = n; s1.b_off = t0; s1.ldb = n; s1.K = nbk;                       // conj(Vn) * W^T
    s2.a_off = t0 + (int64_t)n * NB; s2.lda = n; s2.b_off = t0 - (int64_t)n * NB; s2.ldb = n; s2.K = nbk;   // conj(Wn) * Vp^T
    segs[2 * pi] = s1; segs[2 * pi + 1] = s2;
  }
  dev_h2d(b, d_tasks, tasks.data(), sizeof(GemmTask) * tasks.size());
  dev_h2d(b, d_segs, segs.data(), sizeof(GemmSeg) * segs.size());
  long long nn = n;
  static const int fsplit_env = [] { const char* v = getenv("CB200_TRI_FSPLIT"); return v ? atoi(v) : 1; }();
  int fsplit = std::max(1, std::min(fsplit_env, TRI_WARPS));
  for (int pi = 0; pi < npanel; ++pi) {
    long long j0 = (long long)pi * NB;
    int nbk = (int)std::min<int64_t>(NB, nref - j0);
    void* args[] = {&A, &nn, &j0, &nbk, &d, &e, &tau, &scl, &p, &gbuf, &partial, &Vp, &W, &Vn, &Wn, &Vt, &Wt, &fsplit};
    CB_CHECK(b, cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(TRI_THREADS), args, 0, b->stream));
    CB_LAUNCH_CHECK(b, "tridiag_panel_kernel");
    const int64_t t0 = j0 + nbk;
    if (t0 < n - 1 && tiles[pi] > 0)    // (the last panel updates the final diagonal element itself)
      launch_gemm(b, cplx, false, true, GEMM_TILE_WIDE, Vn, W, A, d_tasks + pi, 1, d_segs, tiles[pi]);
  }
  dev_free(b, pan); dev_free(b, gbuf); dev_free(b, partial); dev_free(b, d_tasks); dev_free(b, d_segs);
  return true;
}

bool launch_tridiag_batched(Backend* b, bool cplx, const TriProblem* probs, int nprob) {
  static const bool enabled = [] { const char* v = getenv("CB200_TRI_BATCH"); return !(v && atoi(v) == 0); }();
  static const bool unblocked = [] { const char* v = getenv("CB200_TRIDIAG"); return v && std::string(v) == "unblocked"; }();
  // matrices that take part in the batched launches; the rest (tiny ones) go through the single-matrix path
  std::vector<int> big;
  for (int k = 0; k < nprob; ++k)
    if (probs[k].n >= 3) big.push_back(k);
  constexpr int MAX_BATCH = 8;
  if (!enabled || unblocked || big.size() < 2 || big.size() > MAX_BATCH) {
    for (int k = 0; k < nprob; ++k)
      if (!launch_tridiag(b, cplx, probs[k].A, probs[k].n, probs[k].d, probs[k].e, probs[k].tau, probs[k].scl, probs[k].p_work)) return false;
    return true;
  }
  for (int k = 0; k < nprob; ++k)
    if (probs[k].n > 0 && probs[k].n < 3)
      launch_tridiag(b, cplx, probs[k].A, probs[k].n, probs[k].d, probs[k].e, probs[k].tau, probs[k].scl, probs[k].p_work);
  const int NB = 64;
  const size_t es = cplx ? 16 : 8;
  const int np = (int)big.size();
  static int capacity[2] = {0, 0};
  const void* fn = cplx ? (const void*)tridiag_panel_batched_kernel<true> : (const void*)tridiag_panel_batched_kernel<false>;
  if (!capacity[cplx]) {
    int per_sm = 0, sms = 0;
    CB_CHECK(b, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, TRI_THREADS, 0));
    CB_CHECK(b, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
    capacity[cplx] = std::max(1, per_sm) * std::max(1, sms);
  }
  if (capacity[cplx] < np) {   // cannot give every matrix a CTA: one after the other
    for (int k : big)
      if (!launch_tridiag(b, cplx, probs[k].A, probs[k].n, probs[k].d, probs[k].e, probs[k].tau, probs[k].scl, probs[k].p_work)) return false;
    return true;
  }
  static const int fsplit_env = [] { const char* v = getenv("CB200_TRI_FSPLIT"); return v ? atoi(v) : 1; }();
  const int fsplit = std::max(1, std::min(fsplit_env, TRI_WARPS));
  // CTAs per matrix: an equal share of the co-resident capacity, but no more than one warp per column needs
  struct Work {
    int64_t n = 0, nref = 0;
    int npanel = 0, blocks = 0;
    char* pan = nullptr; char* gbuf = nullptr; double* partial = nullptr;
    GemmTask* d_tasks = nullptr; GemmSeg* d_segs = nullptr;
    std::vector<int> tiles;
  };
  std::vector<Work> w(np);
  int npanel_max = 0;
  const int share = capacity[cplx] / np;
  const int bm = gemm_tile_m(cplx), bn = gemm_tile_n(cplx, GEMM_TILE_WIDE);
  bool ok = true;
  for (int q = 0; q < np; ++q) {
    const TriProblem& P = probs[big[q]];
    Work& x = w[q];
    x.n = P.n; x.nref = P.n - 1;
    x.npanel = (int)((x.nref + NB - 1) / NB);
    npanel_max = std::max(npanel_max, x.npanel);
    x.blocks = (int)std::min<int64_t>(share, std::max<int64_t>(1, (P.n + TRI_WARPS - 1) / TRI_WARPS));
    x.pan = (char*)dev_alloc(b, 6 * es * (size_t)P.n * NB);          // Vp | W | Vn | Wn (n x NB, ld n) | Vt | Wt (NB x n, ld NB)
    x.gbuf = (char*)dev_alloc(b, es * 128);
    x.partial = (double*)dev_alloc(b, sizeof(double) * (size_t)x.blocks);
    x.d_tasks = (GemmTask*)dev_alloc(b, sizeof(GemmTask) * (size_t)x.npanel);
    x.d_segs = (GemmSeg*)dev_alloc(b, sizeof(GemmSeg) * 2 * (size_t)x.npanel);
    if (!x.pan || !x.gbuf || !x.partial || !x.d_tasks || !x.d_segs) { ok = false; break; }
    // rank-2nb updates after each panel: A[t0:, t0:] += conj( conj(Vn) W^T + conj(Wn) Vp^T ) = -(V W^H + W V^H)   (as in launch_tridiag)
    std::vector<GemmTask> tasks(x.npanel);
    std::vector<GemmSeg> segs(2 * (size_t)x.npanel);
    x.tiles.assign(x.npanel, 0);
    const int64_t n = P.n;
    for (int pi = 0; pi < x.npanel; ++pi) {
      const int64_t j0 = (int64_t)pi * NB;
      const int nbk = (int)std::min<int64_t>(NB, x.nref - j0);
      const int64_t t0 = j0 + nbk, mt = n - t0;
      GemmTask t{};
      t.c_off = t0 + t0 * n; t.c_off2 = t.c_off; t.ldc = n;
      t.M = (int32_t)mt; t.N = (int32_t)mt; t.n_split = (int32_t)mt;
      t.seg_begin = 2 * pi; t.seg_count = 2;
      t.tile_begin = 0; t.tiles_m = (int32_t)((mt + bm - 1) / bm);
      t.flags = GEMM_ACCUM | (cplx ? (GEMM_CONJ_A | GEMM_CONJ_OUT) : 0u);
      x.tiles[pi] = t.tiles_m * (int)((mt + bn - 1) / bn);
      tasks[pi] = t;
      GemmSeg s1{}, s2{};
      s1.a_off = t0; s1.lda = n; s1.b_off = t0; s1.ldb = n; s1.K = nbk;
      s2.a_off = t0 + (int64_t)n * NB; s2.lda = n; s2.b_off = t0 - (int64_t)n * NB; s2.ldb = n; s2.K = nbk;
      segs[2 * pi] = s1; segs[2 * pi + 1] = s2;
    }
    dev_h2d(b, x.d_tasks, tasks.data(), sizeof(GemmTask) * tasks.size());
    dev_h2d(b, x.d_segs, segs.data(), sizeof(GemmSeg) * segs.size());
  }
  // descriptors of every (panel step, matrix) + one arrival counter each, uploaded once
  std::vector<TriDesc> descs;
  std::vector<int> first(npanel_max + 1, 0);
  TriDesc* d_descs = nullptr;
  unsigned int* d_ctr = nullptr;
  if (ok) {
    d_ctr = (unsigned int*)dev_alloc(b, sizeof(unsigned int) * (size_t)npanel_max * np);
    if (d_ctr) dev_memset0(b, d_ctr, sizeof(unsigned int) * (size_t)npanel_max * np);
    for (int pi = 0; pi < npanel_max; ++pi) {
      first[pi] = (int)descs.size();
      int blk = 0;
      for (int q = 0; q < np; ++q) {
        const Work& x = w[q];
        if (pi >= x.npanel) continue;
        const TriProblem& P = probs[big[q]];
        const size_t pn = es * (size_t)x.n * NB;
        TriDesc D{};
        D.A = P.A; D.n = x.n; D.j0 = (long long)pi * NB;
        D.nbk = (int)std::min<int64_t>(NB, x.nref - D.j0);
        D.d = P.d; D.e = P.e; D.tau = P.tau; D.scl = P.scl; D.p = P.p_work; D.g = x.gbuf; D.partial = x.partial;
        D.Vp = x.pan; D.W = x.pan + pn; D.Vn = x.pan + 2 * pn; D.Wn = x.pan + 3 * pn; D.Vt = x.pan + 4 * pn; D.Wt = x.pan + 5 * pn;
        D.ctr = d_ctr ? d_ctr + descs.size() : nullptr;
        D.blk_begin = blk; D.blk_count = x.blocks; D.fsplit = fsplit;
        blk += x.blocks;
        descs.push_back(D);
      }
    }
    first[npanel_max] = (int)descs.size();
    d_descs = (TriDesc*)dev_alloc(b, sizeof(TriDesc) * descs.size());
    if (!d_ctr || !d_descs) ok = false;
    else dev_h2d(b, d_descs, descs.data(), sizeof(TriDesc) * descs.size());
  }
  if (ok) {
    for (int pi = 0; pi < npanel_max; ++pi) {
      const int nd = first[pi + 1] - first[pi];
      const TriDesc* dp = d_descs + first[pi];
      const TriDesc& last = descs[first[pi + 1] - 1];
      int grid = last.blk_begin + last.blk_count;
      int ndv = nd;
      void* args[] = {(void*)&dp, &ndv};
      CB_CHECK(b, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(TRI_THREADS), args, 0, b->stream));
      CB_LAUNCH_CHECK(b, "tridiag_panel_batched_kernel");
      for (int q = 0; q < np; ++q) {
        const Work& x = w[q];
        if (pi >= x.npanel) continue;
        const int64_t j0 = (int64_t)pi * NB;
        const int nbk = (int)std::min<int64_t>(NB, x.nref - j0);
        const int64_t t0 = j0 + nbk;
        if (t0 < x.n - 1 && x.tiles[pi] > 0) {
          const size_t pn = es * (size_t)x.n * NB;
          launch_gemm(b, cplx, false, true, GEMM_TILE_WIDE, x.pan + 2 * pn, x.pan + pn, probs[big[q]].A, x.d_tasks + pi, 1, x.d_segs, x.tiles[pi]);
        }
      }
    }
  }
  for (Work& x : w) {
    if (x.pan) dev_free(b, x.pan);
    if (x.gbuf) dev_free(b, x.gbuf);
    if (x.partial) dev_free(b, x.partial);
    if (x.d_tasks) dev_free(b, x.d_tasks);
    if (x.d_segs) dev_free(b, x.d_segs);
  }
  if (d_ctr) dev_free(b, d_ctr);
  if (d_descs) dev_free(b, d_descs);
  if (!ok) set_err(b, "launch_tridiag_batched: allocation of the panel / descriptor buffers failed", cudaErrorMemoryAllocation);
  return ok;   // false: the caller must not read d / e (they were never written)
}

static double* tri_bounds(Backend* b, const double* d, const double* e, int64_t n) {
  if (!b->d_bounds) CB_CHECK(b, cudaMalloc(&b->d_bounds, 8 * sizeof(double)));
  tri_bounds_kernel<<<1, 256, 0, b->stream>>>(d, e, (long long)n, b->d_bounds);
  CB_LAUNCH_CHECK(b, "tri_bounds_kernel");
  return b->d_bounds;
}

void launch_bisect(Backend* b, const double* d, const double* e, int64_t n, double* w) {
  if (n <= 0) return;
  double* bounds = tri_bounds(b, d, e, n);
  bisect_kernel<<<(unsigned)((n + 3) / 4), 128, 0, b->stream>>>(d, e, (long long)n, bounds, w);
  CB_LAUNCH_CHECK(b, "bisect_kernel");
}

void launch_invit(Backend* b, const double* d, const double* e, int64_t n, const double* lam, int64_t m, double* Zt, double* work) {
  if (n <= 0 || m <= 0) return;
  double* bounds = tri_bounds(b, d, e, n);
  static const bool ref = [] { const char* v = getenv("CB200_INVIT"); return v && std::string(v) == "ref"; }();
  if (ref) invit_ref_kernel<<<(unsigned)((m + 127) / 128), 128, 0, b->stream>>>(d, e, (long long)n, lam, (long long)m, bounds, Zt, work);
  else invit_kernel<<<(unsigned)((m + 127) / 128), 128, 0, b->stream>>>(d, e, (long long)n, lam, (long long)m, bounds, Zt, work);
  CB_LAUNCH_CHECK(b, "invit_kernel");
}

void launch_ns_coeff(Backend* b, const double* G, int64_t m, double alpha, double* Cout, double* stats) {
  if (m <= 0) return;
  CB_CHECK(b, cudaMemsetAsync(stats, 0, 16, b->stream));
  ns_coeff_kernel<<<(unsigned)m, 256, 0, b->stream>>>(G, (long long)m, alpha, Cout, (unsigned long long*)stats);
  CB_LAUNCH_CHECK(b, "ns_coeff_kernel");
}

void launch_backtransform(Backend* b, bool cplx, const void* A, int64_t n, const void* tau, const void* scl, const double* Zt, int64_t m, void* V) {
  if (n <= 0 || m <= 0) return;
  const size_t col_bytes = (size_t)n * (cplx ? 16 : 8);
  const size_t budget = 200 * 1024;
  int cpc = (int)std::min<size_t>(16, budget / col_bytes);
  const int use_smem = cpc >= 1;
  if (!use_smem) cpc = 8;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(backtransform_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    cudaFuncSetAttribute(backtransform_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    attr_done = true;
  }
  const unsigned blocks = (unsigned)((m + cpc - 1) / cpc);
  const size_t smem = use_smem ? (size_t)cpc * col_bytes : 0;
  if (cplx) backtransform_kernel<true><<<blocks, 32 * cpc, smem, b->stream>>>(A, (long long)n, tau, scl, Zt, (long long)m, V, use_smem);
  else backtransform_kernel<false><<<blocks, 32 * cpc, smem, b->stream>>>(A, (long long)n, tau, scl, Zt, (long long)m, V, use_smem);
  CB_LAUNCH_CHECK(b, "backtransform_kernel");
}

void launch_zt_to_v(Backend* b, bool cplx, const double* Zt, int64_t n, int64_t m, void* V) {
  if (n <= 0 || m <= 0) return;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((m + 31) / 32));
  if (cplx) zt_to_v_kernel<true><<<grid, 256, 0, b->stream>>>(Zt, (long long)n, (long long)m, V);
  else zt_to_v_kernel<false><<<grid, 256, 0, b->stream>>>(Zt, (long long)n, (long long)m, V);
  CB_LAUNCH_CHECK(b, "zt_to_v_kernel");
}

void launch_form_y(Backend* b, bool cplx, const void* A, int64_t n, const void* tau, const void* scl, int64_t j0, int nbk, void* Y) {
  const int64_t len = n - j0 - 1;
  if (len <= 0 || nbk <= 0) return;
  dim3 grid((unsigned)std::min<int64_t>((len + 255) / 256, 64), (unsigned)nbk);
  if (cplx) form_y_kernel<true><<<grid, 256, 0, b->stream>>>(A, (long long)n, tau, scl, (long long)j0, nbk, Y);
  else form_y_kernel<false><<<grid, 256, 0, b->stream>>>(A, (long long)n, tau, scl, (long long)j0, nbk, Y);
  CB_LAUNCH_CHECK(b, "form_y_kernel");
}

void launch_wy_t_batched(Backend* b, bool cplx, const void* S, const void* tau, int npanel, int nb_last, void* Tneg) {
  if (npanel <= 0) return;
  if (nb_last <= 0 || nb_last > 64) { set_err(b, "launch_wy_t_batched: block size out of range", cudaErrorInvalidValue); return; }
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(wy_t_batched_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 * 16);
    cudaFuncSetAttribute(wy_t_batched_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 * 8);
    attr_done = true;
  }
  const size_t smem = (size_t)64 * 64 * (cplx ? 16 : 8);
  if (cplx) wy_t_batched_kernel<true><<<npanel, 64, smem, b->stream>>>(S, tau, npanel, nb_last, Tneg);
  else wy_t_batched_kernel<false><<<npanel, 64, smem, b->stream>>>(S, tau, npanel, nb_last, Tneg);
  CB_LAUNCH_CHECK(b, "wy_t_batched_kernel");
}

void launch_wy_t(Backend* b, bool cplx, const void* S, const void* tau_j0, int nbk, void* Tneg) {
  if (nbk <= 0 || nbk > 64) { set_err(b, "launch_wy_t: block size out of range", cudaErrorInvalidValue); return; }
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(wy_t_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 * 16);
    cudaFuncSetAttribute(wy_t_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 * 8);
    attr_done = true;
  }
  const size_t smem = (size_t)nbk * nbk * (cplx ? 16 : 8);
  if (cplx) wy_t_kernel<true><<<1, 64, smem, b->stream>>>(S, tau_j0, nbk, Tneg);
  else wy_t_kernel<false><<<1, 64, smem, b->stream>>>(S, tau_j0, nbk, Tneg);
  CB_LAUNCH_CHECK(b, "wy_t_kernel");
}

// ---- two-stage reduction (heig2.cuh) ------------------------------------------------------------------------------------------
static int sm_count(Backend* b) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device);
  return std::max(1, sms);
}
static void opt_in_smem(const void* fn, size_t bytes) { cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); }

void launch_panel_qr(Backend* b, bool cplx, void* A, int64_t n, int64_t j0, void* Y, void* Tn, void* Tnh) {
  const int64_t len = n - (j0 + SB_B);
  if (len < 2) { set_err(b, "launch_panel_qr: panel with fewer than two rows", cudaErrorInvalidValue); return; }
  const size_t es = cplx ? 16 : 8;
  const int rmax = cplx ? PQ_MAXROWS_CPLX : PQ_MAXROWS_REAL;
  const void* fn = cplx ? (const void*)panel_qr_kernel<true> : (const void*)panel_qr_kernel<false>;
  opt_in_smem(fn, cplx ? pq_smem_bytes<true>(rmax) : pq_smem_bytes<false>(rmax));     // per device; cheap
  const int cap = sm_count(b);                                                           // 1 CTA per SM (launch bounds), co-resident
  int nblk = (int)std::min<int64_t>(cap, std::max<int64_t>(1, (len + 63) / 64));
  int R = (int)((len + nblk - 1) / nblk);
  if (R > rmax) { nblk = (int)((len + rmax - 1) / rmax); R = (int)((len + nblk - 1) / nblk); }
  if (nblk > cap) { set_err(b, "launch_panel_qr: matrix too large for the co-resident panel kernel", cudaErrorInvalidValue); return; }
  const size_t smem = cplx ? pq_smem_bytes<true>(R) : pq_smem_bytes<false>(R);
  char* scratch = (char*)dev_alloc(b, es * (2 * (size_t)nblk * SB_B + 2 * SB_B));
  if (!scratch) return;
  void* part = scratch;
  void* rowc = scratch + es * 2 * (size_t)nblk * SB_B;
  long long nn = n, jj = j0;
  void* args[] = {&A, &nn, &jj, &R, &Y, &Tn, &Tnh, &part, &rowc};
  CB_CHECK(b, cudaLaunchCooperativeKernel(fn, dim3(nblk), dim3(PQ_THREADS), args, smem, b->stream));
  CB_LAUNCH_CHECK(b, "panel_qr_kernel");
  dev_free(b, scratch);
}

void launch_band_extract(Backend* b, bool cplx, const void* A, int64_t n, void* AB) {
  const int blocks = (int)std::min<int64_t>((n * SB_LD + 255) / 256, 148 * 16);
  if (cplx) band_extract_kernel<true><<<blocks, 256, 0, b->stream>>>(A, n, AB);
  else band_extract_kernel<false><<<blocks, 256, 0, b->stream>>>(A, n, AB);
  CB_LAUNCH_CHECK(b, "band_extract_kernel");
}

int64_t chase_blocks(int64_t n) {
  const int64_t K0 = (n - 1 + SB_B - 1) / SB_B;
  return std::max<int64_t>(K0 * (K0 + 1) / 2, 1);
}

void launch_bulge_chase(Backend* b, bool cplx, const ChaseProblem* probs, int nprob) {
  if (nprob <= 0) return;
  std::vector<ChaseDesc> descs(nprob);
  int64_t max_sweeps = 0;
  for (int p = 0; p < nprob; ++p) {
    descs[p] = ChaseDesc{probs[p].AB, (long long)probs[p].n, probs[p].d, probs[p].e, probs[p].V2, probs[p].tau2, probs[p].prog};
    max_sweeps = std::max<int64_t>(max_sweeps, probs[p].n - 1);
  }
  if (max_sweeps <= 0) return;
  const void* fn = cplx ? (const void*)bulge_chase_kernel<true> : (const void*)bulge_chase_kernel<false>;
  const size_t smem = cplx ? bc_smem_bytes<true>() : bc_smem_bytes<false>();
  opt_in_smem(fn, smem);
  int per_sm = 0;
  CB_CHECK(b, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, BC_THREADS, smem));
  const int64_t cap = (int64_t)std::max(1, per_sm) * sm_count(b);
  // at most ~n/(2 SB_B) sweeps of one matrix are in flight at a time (each waits for its predecessor to be two tasks ahead)
  int64_t useful = 0;
  for (int p = 0; p < nprob; ++p) useful += probs[p].n / (2 * SB_B) + 2;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(cap, std::min<int64_t>(useful, max_sweeps * nprob)));
  char* scratch = (char*)dev_alloc(b, sizeof(ChaseDesc) * (size_t)nprob + 64);
  if (!scratch) return;
  dev_memset0(b, scratch, 64);
  dev_h2d(b, scratch + 64, descs.data(), sizeof(ChaseDesc) * (size_t)nprob);
  const ChaseDesc* dp = (const ChaseDesc*)(scratch + 64);
  unsigned long long* ticket = (unsigned long long*)scratch;
  long long ms = max_sweeps;
  int np = nprob;
  void* args[] = {(void*)&dp, &np, &ms, &ticket};
  CB_CHECK(b, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(BC_THREADS), args, smem, b->stream));
  CB_LAUNCH_CHECK(b, "bulge_chase_kernel");
  dev_free(b, scratch);
}

void launch_q2_apply(Backend* b, bool cplx, const void* V2, const void* tau2, int64_t n, void* Z, int64_t m) {
  if (n < 2 || m <= 0) return;
  const size_t es = cplx ? 16 : 8;
  const int64_t nblk = chase_blocks(n);
  // T factors of the groups of 8 reflectors (one warp per group), then the blockwise application
  const int G = cplx ? q2_group<true>() : q2_group<false>();
  void* T8 = dev_alloc(b, es * ((size_t)nblk * SB_B * G + 16));
  if (!T8) return;
  {
    const long long warps = (long long)nblk * (SB_B / G);
    const int blocks = (int)((warps * 32 + 255) / 256);
    if (cplx) q2_tprep_kernel<true><<<blocks, 256, 0, b->stream>>>(V2, tau2, nblk, T8);
    else q2_tprep_kernel<false><<<blocks, 256, 0, b->stream>>>(V2, tau2, nblk, T8);
    CB_LAUNCH_CHECK(b, "q2_tprep_kernel");
  }
  const void* fn = cplx ? (const void*)q2_apply_kernel<true> : (const void*)q2_apply_kernel<false>;
  const size_t smem = cplx ? q2_smem_bytes<true>() : q2_smem_bytes<false>();
  opt_in_smem(fn, smem);
  // a CTA applies ALL reflector blocks to its columns, so only the columns are parallel: 4 columns per warp, as few warps per CTA as it
  // takes to give every SM a CTA (the per-block time is shared-memory / shuffle bound and grows with the warp count)
  const int sms = sm_count(b);
  const int warps = (int)std::max<int64_t>(1, std::min<int64_t>(Q2_THREADS / 32, (m + 4 * (int64_t)sms - 1) / (4 * (int64_t)sms)));
  const int grid = (int)((m + 4 * warps - 1) / (4 * warps));
  if (cplx) q2_apply_kernel<true><<<grid, 32 * warps, smem, b->stream>>>(V2, T8, n, Z, m);
  else q2_apply_kernel<false><<<grid, 32 * warps, smem, b->stream>>>(V2, T8, n, Z, m);
  CB_LAUNCH_CHECK(b, "q2_apply_kernel");
  dev_free(b, T8);
}

void launch_truncate_select(Backend* b, const double* pooled, int64_t n, int64_t maxdim, int64_t mindim, double cutoff, double* sorted,
                            double* out3) {
  if (n <= 0) return;
  rank_sort_desc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(pooled, (long long)n, sorted);
  CB_LAUNCH_CHECK(b, "rank_sort_desc_kernel");
  truncate_rule_kernel<<<1, 32, 0, b->stream>>>(sorted, (long long)n, (long long)maxdim, (long long)mindim, cutoff, out3);
  CB_LAUNCH_CHECK(b, "truncate_rule_kernel");
}

// ------------------------------------------------------------------------------------------------ FP64 pipe calibration
// Pure issue-rate loops (no memory traffic): the roofline denominators for the DMMA GEMM (MEASURED_PEAKS.json has no FP64 entry).
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i * 1e-9; }
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;
}
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* sink) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) sink[0] = s;
}

// which: 0 = DMMA.8x8x4 loop, 1 = DFMA loop; returns TFLOP/s
double fp64_pipe_peak(Backend* b, int which, int iters) {
  double* sink = nullptr;
  CB_CHECK(b, cudaMalloc(&sink, 8));
  const int blocks = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    if (rep == 1) cudaEventRecord(e0, b->stream);
    if (which == 0) dmma_peak_kernel<<<blocks, 256, 0, b->stream>>>(iters, sink);
    else dfma_peak_kernel<<<blocks, 256, 0, b->stream>>>(iters, sink);
    CB_LAUNCH_CHECK(b, "fp64_peak_kernel");
  }
  cudaEventRecord(e1, b->stream);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  const double warps = (double)blocks * 8;
  const double flops = which == 0 ? warps * iters * 16.0 * 512.0 : warps * 32.0 * iters * 16.0 * 2.0;
  return flops / (ms * 1e-3) / 1e12;
}

}  // namespace cb200
