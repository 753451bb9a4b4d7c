// extern "C" boundary (include/chain_b200.h).  Marshals plain pointers into the engine and turns C++ exceptions
// into status codes; nothing here touches the CPU for arithmetic.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "../../include/chain_b200.h"
#include "engine.h"

using namespace cb200;

struct cb200_ctx {
  Ctx c;
};
static thread_local std::string g_create_err;

struct cb200_result {
  std::unique_ptr<DmrgResult> r;
};

struct cb200_bond {
  cb200_ctx* ctx = nullptr;
  bool cplx = false;
  int nq = 1;
  SiteInfo S;
  IndexMap l, r, kl, km, kr;
  HostW W1, W2;
  DevEnv L, R;
  Bond b;
  std::vector<Buf> pen;
};

template <class F>
static int guard(cb200_ctx* ctx, F f) {
  try {
    f();
    const char* e = dev_last_error(ctx->c.be);
    if (e) { ctx->c.err = e; return CB200_ERR_CUDA; }
    return CB200_OK;
  } catch (const EngineError& e) {
    ctx->c.err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    ctx->c.err = std::string("internal error: ") + e.what();
    return CB200_ERR_INVALID;
  }
}

extern "C" {

int cb200_create(cb200_ctx** out, int device) {
  if (!out) return CB200_ERR_INVALID;
  *out = nullptr;
  char err[512] = {0};
  Backend* be = backend_create(device, err, sizeof(err));
  if (!be) { g_create_err = err; return CB200_ERR_NO_DEVICE; }
  cb200_ctx* c = new cb200_ctx();
  c->c.be = be;
  *out = c;
  return CB200_OK;
}
void cb200_destroy(cb200_ctx* ctx) {
  if (!ctx) return;
  ctx->c.ws.clear();                 // device buffers go before the backend (stream, pool) they belong to
  backend_destroy(ctx->c.be);
  ctx->c.be = nullptr;
  delete ctx;
}
const char* cb200_last_error(const cb200_ctx* ctx) { return ctx ? ctx->c.err.c_str() : g_create_err.c_str(); }
const char* cb200_backend_name(void) { return backend_name(); }
uint64_t cb200_launch_count(const cb200_ctx* ctx) { return ctx ? dev_launch_count(ctx->c.be) : 0; }

// ---------------------------------------------------------------------------------------------- multi-GPU
int cb200_nccl_unique_id(cb200_ctx* ctx, void* out128) {
  if (!ctx || !out128) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    if (nccl_unique_id(ctx->c.be, out128) != 0) fail(CB200_ERR_CUDA, std::string("nccl: ") + (dev_last_error(ctx->c.be) ? dev_last_error(ctx->c.be) : "?"));
  });
}
int cb200_shard_alloc(cb200_ctx* ctx, int64_t landing_bytes, void* handle_out64) {
  if (!ctx || !handle_out64 || landing_bytes <= 0) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    if (shard_alloc(ctx->c.be, landing_bytes, handle_out64) != 0)
      fail(CB200_ERR_CUDA, std::string("shard_alloc: ") + (dev_last_error(ctx->c.be) ? dev_last_error(ctx->c.be) : "?"));
  });
}
int cb200_shard_setup(cb200_ctx* ctx, int32_t rank, int32_t world, int32_t mode, const void* nccl_uid128, const void* handles,
                      int64_t min_dim) {
  if (!ctx) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || mode < 0 || mode > 2) fail(CB200_ERR_INVALID, "shard_setup: bad rank/world/mode");
    if (mode == 1 && world > 1) {
      if (!nccl_uid128) fail(CB200_ERR_INVALID, "shard_setup: mode 1 needs the NCCL unique id");
      if (nccl_init(ctx->c.be, nccl_uid128, rank, world) != 0)
        fail(CB200_ERR_CUDA, std::string("nccl: ") + (dev_last_error(ctx->c.be) ? dev_last_error(ctx->c.be) : "?"));
    }
    if (mode == 2 && world > 1) {
      if (!handles) fail(CB200_ERR_INVALID, "shard_setup: mode 2 needs the ranks' landing handles");
      if (shard_connect(ctx->c.be, rank, world, handles) != 0)
        fail(CB200_ERR_CUDA, std::string("shard_connect: ") + (dev_last_error(ctx->c.be) ? dev_last_error(ctx->c.be) : "?"));
    }
    ctx->c.shard.rank = rank;
    ctx->c.shard.world = world;
    ctx->c.shard.mode = mode;
    ctx->c.shard.flip = 0;
    if (min_dim >= 0) ctx->c.shard.min_dim = min_dim;
  });
}

// ---------------------------------------------------------------------------------------------- argument validation
// Everything a caller can get wrong WITHOUT lying about how much memory it owns is refused here with CB200_ERR_INVALID, in front of any
// allocation or launch (tools/fuzz_abi.py drives these under ASan/UBSan).  CB200_MAX_NQ / CB200_MAX_SITES / CB200_MAX_SITE_DIM are stated
// in include/chain_b200.h.
static void check_group(int nq, int d, const char* who) {
  if (nq < 1 || nq > CB200_MAX_NQ) fail(CB200_ERR_INVALID, std::string(who) + ": nq must be in [1, " + std::to_string(CB200_MAX_NQ) + "]");
  if (d < 1 || d > CB200_MAX_SITE_DIM) fail(CB200_ERR_INVALID, std::string(who) + ": site dimension d must be in [1, " + std::to_string(CB200_MAX_SITE_DIM) + "]");
}
static void check_chain(int n, const void* site, const void* link, const char* who) {
  if (n < 1 || n > CB200_MAX_SITES) fail(CB200_ERR_INVALID, std::string(who) + ": n must be in [1, " + std::to_string(CB200_MAX_SITES) + "]");
  if (!site || !link) fail(CB200_ERR_INVALID, std::string(who) + ": null site / link array");
}
static void check_index(const cb200_index* ix, const char* who) {
  if (!ix) fail(CB200_ERR_INVALID, std::string(who) + ": null index");
  if (ix->dim < 0 || ix->dim > ((int64_t)1 << 31)) fail(CB200_ERR_INVALID, std::string(who) + ": index dimension out of range");
}
static void check_tensor(const cb200_tensor& t, int rank, const std::string& who) {
  if (t.rank != rank) fail(CB200_ERR_INVALID, who + " has rank " + std::to_string(t.rank) + ", expected " + std::to_string(rank));
  if (t.dtype != 0 && t.dtype != 1) fail(CB200_ERR_INVALID, who + ": dtype must be 0 (float64) or 1 (complex128)");
  if (!t.data) fail(CB200_ERR_INVALID, who + " has no data");
}
static void check_site_charges(int nq, int d, const int32_t* site_charge) {
  if (!site_charge) return;                       // NULL = all zero (the dense case)
  for (int s = 0; s < d; ++s)
    if (site_charge[s] < 0 || site_charge[s] >= nq) fail(CB200_ERR_INVALID, "site charge label out of range");
}

// the MPO of a C-ABI call as the engine takes it (shared by cb200_dmrg and cb200_mps_mpo_element)
static MpoIn to_mpo_in(const cb200_mpo* H, int nq, const int32_t* site_charge) {
  check_group(nq, H->d, "MPO");
  check_chain(H->n, H->site, H->link, "MPO");
  check_site_charges(nq, H->d, site_charge);
  MpoIn mi;
  mi.n = H->n; mi.nq = nq; mi.d = H->d; mi.site_charge = site_charge;
  mi.link.resize(H->n + 1);
  for (int j = 0; j <= H->n; ++j) {
    check_index(&H->link[j], "MPO link");
    mi.link[j] = IndexMap::from_labels(nq, H->link[j].dim, H->link[j].charge);
  }
  mi.site.resize(H->n);
  for (int j = 0; j < H->n; ++j) {
    const cb200_tensor& t = H->site[j];
    check_tensor(t, 4, "MPO site tensor " + std::to_string(j));
    if (t.dims[0] != H->link[j].dim || t.dims[1] != H->d || t.dims[2] != H->d || t.dims[3] != H->link[j + 1].dim)
      fail(CB200_ERR_INVALID, "MPO site tensor " + std::to_string(j) + " has inconsistent dimensions");
    if (j == 0) mi.cplx = t.dtype == 1;
    else if ((t.dtype == 1) != mi.cplx) fail(CB200_ERR_INVALID, "mixed dtypes inside the MPO");
    mi.site[j] = t.data;
  }
  return mi;
}

// ---------------------------------------------------------------------------------------------- coarse: dmrg
static HostMPS to_host_mps(const cb200_mps* m, int nq, int d, bool* cplx) {
  if (!m) fail(CB200_ERR_INVALID, "null MPS");
  check_group(nq, d, "MPS");
  check_chain(m->n, m->site, m->link, "MPS");
  if (m->nq != nq) fail(CB200_ERR_INVALID, "MPS: nq differs from the symmetry group of the call");
  if (m->ortho_center < 0 || m->ortho_center > m->n) fail(CB200_ERR_INVALID, "MPS: ortho_center must be 0 (unknown) or a 1-based site number");
  for (int j = 0; j <= m->n; ++j) check_index(&m->link[j], "MPS link");
  HostMPS h;
  h.n = m->n;
  h.center = m->ortho_center;
  h.link.resize(m->n + 1);
  for (int j = 0; j <= m->n; ++j) h.link[j] = IndexMap::from_labels(nq, m->link[j].dim, m->link[j].charge);
  h.site.resize(m->n);
  if (m->layout != 0 && m->layout != 1) fail(CB200_ERR_INVALID, "MPS layout must be 0 (dense) or 1 (block-packed)");
  h.packed = m->layout == 1;
  *cplx = false;
  for (int j = 0; j < m->n; ++j) {
    const cb200_tensor& t = m->site[j];
    check_tensor(t, 3, "MPS site tensor " + std::to_string(j));
    if (t.dims[0] != m->link[j].dim || t.dims[1] != d || t.dims[2] != m->link[j + 1].dim)
      fail(CB200_ERR_INVALID, "MPS site tensor " + std::to_string(j) + " has inconsistent dimensions");
    if (j == 0) *cplx = t.dtype == 1;
    else if ((t.dtype == 1) != *cplx) fail(CB200_ERR_INVALID, "mixed dtypes inside one MPS");
    h.site[j] = t.data;
  }
  return h;
}

int cb200_dmrg(cb200_ctx* ctx, const cb200_mpo* H, const cb200_mps* prev, int32_t nprev, const cb200_mps* psi0,
               const cb200_sweep* sweeps, int32_t nsweep, double weight, double* energy, cb200_result** out) {
  if (!ctx) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    if (!H || !psi0 || !sweeps || !energy || !out || nsweep < 0 || nprev < 0 || (nprev > 0 && !prev)) fail(CB200_ERR_INVALID, "null argument");
    if (H->n != psi0->n) fail(CB200_ERR_INVALID, "MPO and MPS have different lengths");
    for (int i = 0; i < nsweep; ++i) {
      const cb200_sweep& w = sweeps[i];
      if (w.maxdim < 1 || w.mindim < 0 || w.niter < 0) fail(CB200_ERR_INVALID, "sweep " + std::to_string(i) + ": maxdim >= 1, mindim >= 0 and niter >= 0 are required");
      if (!(w.cutoff >= 0) || std::isinf(w.cutoff)) fail(CB200_ERR_INVALID, "sweep " + std::to_string(i) + ": cutoff must be a finite number >= 0");
      if (std::isnan(w.noise)) fail(CB200_ERR_INVALID, "sweep " + std::to_string(i) + ": noise is NaN");
    }
    if (!std::isfinite(weight)) fail(CB200_ERR_INVALID, "weight must be finite");
    MpoIn mi = to_mpo_in(H, H->nq, H->site_charge);
    bool pc = false;
    HostMPS p0 = to_host_mps(psi0, H->nq, H->d, &pc);
    std::vector<HostMPS> pv;
    std::vector<bool> pvc;
    for (int p = 0; p < nprev; ++p) {
      bool cx = false;
      pv.push_back(to_host_mps(&prev[p], H->nq, H->d, &cx));
      pvc.push_back(cx);
    }
    std::vector<SweepParams> sw(nsweep);
    for (int i = 0; i < nsweep; ++i) sw[i] = {sweeps[i].maxdim, sweeps[i].mindim, sweeps[i].niter, sweeps[i].cutoff, sweeps[i].noise};
    auto r = run_dmrg(&ctx->c, mi, pv, pvc, p0, pc, sw, weight);
    if (!std::isfinite(r->energy)) fail(CB200_ERR_NOCONV, "dmrg: the energy is not finite (NaN / Inf in the MPO or the start state?)");
    *energy = r->energy;
    cb200_result* res = new cb200_result();
    res->r = std::move(r);
    *out = res;
  });
}

int cb200_mps_entropies(cb200_ctx* ctx, const cb200_mps* psi, int32_t d, const int32_t* site_charge, double* out) {
  if (!ctx || !psi || !out) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    bool cx = false;
    HostMPS h = to_host_mps(psi, psi->nq, d, &cx);
    check_site_charges(psi->nq, d, site_charge);
    std::vector<double> ee = mps_entropies(&ctx->c, psi->nq, d, site_charge, h, cx);
    for (size_t i = 0; i < ee.size(); ++i) out[i] = ee[i];
  });
}

int cb200_mps_translate(cb200_ctx* ctx, const cb200_mps* psi, int32_t d, const int32_t* site_charge, double cutoff, cb200_result** out) {
  if (!ctx || !psi || !out) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    bool cx = false;
    HostMPS h = to_host_mps(psi, psi->nq, d, &cx);
    check_site_charges(psi->nq, d, site_charge);
    if (!(cutoff >= 0)) fail(CB200_ERR_INVALID, "translate: cutoff must be >= 0");
    auto r = mps_translate(&ctx->c, psi->nq, d, site_charge, h, cx, cutoff);
    cb200_result* res = new cb200_result();
    res->r = std::move(r);
    *out = res;
  });
}

int cb200_mps_overlap(cb200_ctx* ctx, const cb200_mps* a, const cb200_mps* b, int32_t d, const int32_t* site_charge, double* re_im) {
  if (!ctx || !a || !b || !re_im) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    if (a->nq != b->nq) fail(CB200_ERR_INVALID, "overlap: the two states use different symmetry groups");
    bool ca = false, cbb = false;
    HostMPS ha = to_host_mps(a, a->nq, d, &ca), hb = to_host_mps(b, b->nq, d, &cbb);
    check_site_charges(a->nq, d, site_charge);
    const cplx_t z = mps_overlap(&ctx->c, a->nq, d, site_charge, ha, ca, hb, cbb);
    re_im[0] = z.real(); re_im[1] = z.imag();
  });
}

int cb200_mps_mpo_element(cb200_ctx* ctx, const cb200_mps* bra, const cb200_mpo* O, const cb200_mps* ket, double* re_im) {
  if (!ctx || !bra || !O || !ket || !re_im) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    if (O->nq != 1 || bra->nq != 1 || ket->nq != 1) fail(CB200_ERR_INVALID, "mpo_element: dense (nq = 1) operands only");
    MpoIn mi = to_mpo_in(O, 1, nullptr);
    bool cbr = false, ck = false;
    HostMPS hb = to_host_mps(bra, 1, O->d, &cbr), hk = to_host_mps(ket, 1, O->d, &ck);
    const cplx_t z = mps_mpo_element(&ctx->c, O->d, hb, cbr, mi, hk, ck);
    re_im[0] = z.real(); re_im[1] = z.imag();
  });
}

int32_t cb200_result_nsites(const cb200_result* r) { return r ? r->r->psi.n : 0; }
int32_t cb200_result_dtype(const cb200_result* r) { return r && r->r->cplx ? 1 : 0; }
int64_t cb200_result_link_dim(const cb200_result* r, int32_t link) {
  if (!r || link < 0 || link > r->r->psi.n) return -1;
  return r->r->psi.link[link].g.total();
}
int cb200_result_link_charges(const cb200_result* r, int32_t link, int32_t* out) {
  if (!r || !out || link < 0 || link > r->r->psi.n) return CB200_ERR_INVALID;
  const auto& m = r->r->psi.link[link];
  for (size_t i = 0; i < m.charge_ext.size(); ++i) out[i] = m.charge_ext[i];
  return CB200_OK;
}
int cb200_result_site(const cb200_result* r, int32_t site, void* out) {
  if (!r || !out || site < 0 || site >= r->r->psi.n || !r->r->ctx) return CB200_ERR_INVALID;
  Ctx* c = r->r->ctx;
  try {
    result_site_export(*r->r, site, out);
    return CB200_OK;
  } catch (const EngineError& e) {
    c->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    c->err = std::string("internal error: ") + e.what();
    return CB200_ERR_INVALID;
  }
}
int64_t cb200_site_blocks_elems(int32_t nq, int32_t d, const int32_t* site_charge, const cb200_index* l, const cb200_index* r) {
  if (nq < 1 || d < 1 || !l || !r) return -1;
  try {
    SiteInfo S = SiteInfo::make(nq, d, site_charge);
    IndexMap ml = IndexMap::from_labels(nq, l->dim, l->charge), mr = IndexMap::from_labels(nq, r->dim, r->charge);
    return site3_block_tasks(S, ml, mr, nullptr, true, nullptr);
  } catch (...) {
    return -1;
  }
}
int64_t cb200_result_site_blocks_elems(const cb200_result* r, int32_t site) {
  if (!r || site < 0 || site >= r->r->psi.n) return -1;
  return result_site_blocks_elems(*r->r, site);
}
int cb200_result_site_blocks(const cb200_result* r, int32_t site, void* out) {
  if (!r || !out || site < 0 || site >= r->r->psi.n || !r->r->ctx) return CB200_ERR_INVALID;
  Ctx* c = r->r->ctx;
  try {
    result_site_export_blocks(*r->r, site, out);
    return CB200_OK;
  } catch (const EngineError& e) {
    c->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    c->err = std::string("internal error: ") + e.what();
    return CB200_ERR_INVALID;
  }
}
int32_t cb200_result_nstats(const cb200_result* r) { return r ? (int32_t)r->r->stats.size() : 0; }
int cb200_result_stats(const cb200_result* r, cb200_bond_stat* out) {
  if (!r || !out) return CB200_ERR_INVALID;
  for (size_t i = 0; i < r->r->stats.size(); ++i) {
    const BondStat& s = r->r->stats[i];
    out[i].sweep = s.sweep; out[i].bond = s.bond; out[i].dir = s.dir; out[i].m = s.m;
    out[i].energy = s.energy; out[i].truncerr = s.truncerr;
  }
  return CB200_OK;
}
int cb200_result_timers(const cb200_result* r, double* o) {
  if (!r || !o) return CB200_ERR_INVALID;
  const Timers& t = r->r->tm;
  o[0] = t.env; o[1] = t.matvec; o[2] = t.level1; o[3] = t.trunc; o[4] = t.other; o[5] = t.nmatvec; o[6] = t.matvec_flops;
  return CB200_OK;
}
void cb200_result_free(cb200_result* r) { delete r; }

// ---------------------------------------------------------------------------------------------- fine: one bond
int cb200_bond_create(cb200_ctx* ctx, int32_t dtype, int32_t nq, int32_t d, const int32_t* site_charge, const cb200_index* l,
                      const cb200_index* r, const cb200_index* kl, const cb200_index* km, const cb200_index* kr, const void* L,
                      const void* W1, const void* W2, const void* R, cb200_bond** out) {
  if (!ctx || !out) return CB200_ERR_INVALID;
  *out = nullptr;
  cb200_bond* bp = new cb200_bond();
  int rc = guard(ctx, [&] {
    cb200_bond& B = *bp;
    if (dtype != 0 && dtype != 1) fail(CB200_ERR_INVALID, "bond_create: dtype must be 0 (float64) or 1 (complex128)");
    check_group(nq, d, "bond_create");
    check_site_charges(nq, d, site_charge);
    for (const cb200_index* ix : {l, r, kl, km, kr}) check_index(ix, "bond_create");
    if (!L || !W1 || !W2 || !R) fail(CB200_ERR_INVALID, "bond_create: null environment / MPO tensor");
    B.ctx = ctx; B.cplx = dtype == 1; B.nq = nq;
    B.S = SiteInfo::make(nq, d, site_charge);
    B.l = IndexMap::from_labels(nq, l->dim, l->charge);
    B.r = IndexMap::from_labels(nq, r->dim, r->charge);
    B.kl = IndexMap::from_labels(nq, kl->dim, kl->charge);
    B.km = IndexMap::from_labels(nq, km->dim, km->charge);
    B.kr = IndexMap::from_labels(nq, kr->dim, kr->charge);
    B.W1 = import_w(nq, d, B.S, B.kl, B.km, B.cplx, W1);
    B.W2 = import_w(nq, d, B.S, B.km, B.kr, B.cplx, W2);
    import_env(&ctx->c, B.cplx, B.l, B.kl, L, B.L);
    import_env(&ctx->c, B.cplx, B.r, B.kr, R, B.R);
    Bond& b = B.b;
    b.ctx = &ctx->c; b.cplx = B.cplx; b.nq = nq; b.S = &B.S;
    b.Dl = B.l.g; b.Dr = B.r.g; b.Kl = B.kl.g; b.Km = B.km.g; b.Kr = B.kr.g;
    b.L = B.L.buf.p; b.R = B.R.buf.p; b.W1 = &B.W1; b.W2 = &B.W2;
    b.plan();
  });
  if (rc != CB200_OK) { delete bp; return rc; }
  *out = bp;
  return CB200_OK;
}
void cb200_bond_free(cb200_bond* b) { delete b; }

int cb200_bond_set_penalty(cb200_bond* B, int32_t nprev, const void* const* o, double weight) {
  if (!B || nprev < 0 || (nprev > 0 && !o)) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    for (int p = 0; p < nprev; ++p)
      if (!o[p]) fail(CB200_ERR_INVALID, "set_penalty: null vector");
    if (!std::isfinite(weight)) fail(CB200_ERR_INVALID, "set_penalty: weight must be finite");
    B->pen.clear();
    B->b.pen.clear();
    const size_t es = B->cplx ? 16 : 8;
    for (int p = 0; p < nprev; ++p) {
      B->pen.emplace_back(B->ctx->c.be, (int64_t)es * std::max<int64_t>(B->b.philay.total, 1), true);
      import_phi(&B->ctx->c, B->cplx, B->S, B->l, B->r, B->b.philay, o[p], B->pen.back().p);
      B->b.pen.push_back(B->pen.back().p);
    }
    B->b.weight = weight;
  });
}
int64_t cb200_bond_phi_elems(const cb200_bond* B) { return B ? B->l.g.total() * B->S.d * B->S.d * B->r.g.total() : 0; }
double cb200_bond_matvec_flops(const cb200_bond* B) { return B ? B->b.flops_alg_full : 0; }

int cb200_bond_apply(cb200_bond* B, const void* phi, void* out) {
  if (!B || !phi || !out) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    Ctx* c = &B->ctx->c;
    const size_t es = B->cplx ? 16 : 8;
    const int64_t n = std::max<int64_t>(B->b.philay.total, 1);
    Buf x(c->be, (int64_t)es * n, true), y(c->be, (int64_t)es * n, true);
    import_phi(c, B->cplx, B->S, B->l, B->r, B->b.philay, phi, x.p);
    B->b.apply(x.p, y.p);
    export_phi(c, B->cplx, B->S, B->l, B->r, B->b.philay, y.p, out);
  });
}

int64_t cb200_bond_phi_blocks_elems(const cb200_bond* B) { return B ? B->b.philay.total : 0; }

int cb200_bond_phi_blocks(const cb200_bond* B, int32_t* nblocks, int64_t* offsets, int64_t* dims4) {
  if (!B || !nblocks) return CB200_ERR_INVALID;
  // the block list of the packed layout in storage order: offsets[i] (elements) and dims4[4*i..] = (len_l, len_s1, len_s2, len_r)
  int32_t n = 0;
  int64_t off = 0;
  auto runs = [](const IndexMap& m) {
    std::vector<std::pair<int64_t, int>> v;   // (length, charge)
    for (size_t i = 0; i < m.charge_ext.size();) {
      size_t j = i;
      while (j < m.charge_ext.size() && m.charge_ext[j] == m.charge_ext[i]) ++j;
      v.push_back({(int64_t)(j - i), m.charge_ext[i]});
      i = j;
    }
    return v;
  };
  const auto rl = runs(B->l), rs = runs(B->S.map), rr = runs(B->r);
  for (auto& ur : rr)
    for (auto& u2 : rs)
      for (auto& u1 : rs)
        for (auto& ul : rl) {
          if ((ul.second + u1.second + u2.second - ur.second) % B->nq != 0) continue;
          if (offsets) offsets[n] = off;
          if (dims4) { dims4[4 * n] = ul.first; dims4[4 * n + 1] = u1.first; dims4[4 * n + 2] = u2.first; dims4[4 * n + 3] = ur.first; }
          off += ul.first * u1.first * u2.first * ur.first;
          ++n;
        }
  *nblocks = n;
  return CB200_OK;
}

int cb200_bond_apply_blocks(cb200_bond* B, const void* phi_blocks, void* out_blocks) {
  if (!B || !phi_blocks || !out_blocks) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    Ctx* c = &B->ctx->c;
    const size_t es = B->cplx ? 16 : 8;
    const int64_t n = std::max<int64_t>(B->b.philay.total, 1);
    Buf x(c->be, (int64_t)es * n), y(c->be, (int64_t)es * n);
    import_phi_blocks(c, B->cplx, B->S, B->l, B->r, B->b.philay, phi_blocks, x.p);
    B->b.apply(x.p, y.p);
    export_phi_blocks(c, B->cplx, B->S, B->l, B->r, B->b.philay, y.p, out_blocks);
  });
}

// slice of the block-packed vector that rank `rank` of `world` carries across PCIe in cb200_bond_apply_blocks_sharded
static void phi_slice(const cb200_bond* B, int rank, int world, int64_t* begin, int64_t* count) {
  const int64_t n = B->b.philay.total;
  int64_t chunk = (n + world - 1) / world;
  chunk = (chunk + 1) / 2 * 2;                              // 16-byte granular for f64 as well
  *begin = std::min<int64_t>((int64_t)rank * chunk, n);
  *count = std::min<int64_t>(chunk, n - *begin);
}
int cb200_bond_phi_blocks_slice(const cb200_bond* B, int64_t* begin, int64_t* count) {
  if (!B || !begin || !count) return CB200_ERR_INVALID;
  const ShardState& sh = B->ctx->c.shard;
  phi_slice(B, sh.rank, sh.world, begin, count);
  return CB200_OK;
}

int cb200_bond_apply_blocks_sharded(cb200_bond* B, const void* phi_slice_in, void* out_slice) {
  if (!B || !phi_slice_in || !out_slice) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    Ctx* c = &B->ctx->c;
    const ShardState& sh = c->shard;
    if (sh.world > 1 && sh.mode == 0) fail(CB200_ERR_INVALID, "cb200_bond_apply_blocks_sharded needs an exchange mode (nccl or p2p)");
    const size_t es = B->cplx ? 16 : 8;
    const int64_t n = std::max<int64_t>(B->b.philay.total, 1);
    int64_t b0, cnt;
    phi_slice(B, sh.rank, sh.world, &b0, &cnt);
    int64_t chunk = (B->b.philay.total + sh.world - 1) / sh.world;
    chunk = (chunk + 1) / 2 * 2;
    Buf x(c->be, (int64_t)es * n), y(c->be, (int64_t)es * n), stage(c->be, (int64_t)es * n);
    std::vector<CopyTask> tin, tout;
    if (phi_block_tasks(B->S, B->l, B->r, B->b.philay, true, &tin) != B->b.philay.total) fail(-2, "block-packed phi: element count does not match the charge structure");
    phi_block_tasks(B->S, B->l, B->r, B->b.philay, false, &tout);
    // rank g uploads ONLY its slice; the slices are all-gathered on the device side (NVLink peer stores or ncclBroadcast per root)
    if (cnt > 0) dev_h2d(c->be, (char*)stage.p + es * (size_t)b0, phi_slice_in, es * (size_t)cnt);
    if (sh.world > 1) {
      if (sh.mode == 2) {
        shard_allgather_p2p(c->be, stage.p, (int64_t)es * B->b.philay.total, (int64_t)es * chunk, c->shard.flip);
        c->shard.flip ^= 1;
      } else {
        for (int g = 0; g < sh.world; ++g) {
          const int64_t gb = std::min<int64_t>((int64_t)g * chunk, B->b.philay.total), gc = std::min<int64_t>(chunk, B->b.philay.total - gb);
          if (gc > 0) nccl_bcast(c->be, (char*)stage.p + es * (size_t)gb, (int64_t)es * gc, g);
        }
      }
    }
    run_copy_tasks(c, B->cplx, tin, stage.p, x.p);
    B->b.apply(x.p, y.p);
    // every rank holds the whole product (the exchange is fused into the R-stage); each downloads its slice: the result crosses PCIe once
    run_copy_tasks(c, B->cplx, tout, y.p, stage.p);
    if (cnt > 0) dev_d2h(c->be, out_slice, (char*)stage.p + es * (size_t)b0, es * (size_t)cnt); else dev_sync(c->be);
  });
}

int cb200_bond_apply_timed(cb200_bond* B, const void* phi, int32_t reps, int32_t warmup, double* ms_mean) {
  if (!B || !phi || !ms_mean || reps < 1) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    Ctx* c = &B->ctx->c;
    const size_t es = B->cplx ? 16 : 8;
    const int64_t n = std::max<int64_t>(B->b.philay.total, 1);
    Buf x(c->be, (int64_t)es * n, true), y(c->be, (int64_t)es * n, true);
    import_phi(c, B->cplx, B->S, B->l, B->r, B->b.philay, phi, x.p);
    for (int i = 0; i < warmup; ++i) B->b.apply(x.p, y.p);
    dev_sync(c->be);
    dev_event_record(c->be, 0);
    for (int i = 0; i < reps; ++i) B->b.apply(x.p, y.p);
    dev_event_record(c->be, 1);
    *ms_mean = dev_event_elapsed_ms(c->be, 0, 1) / reps;
  });
}

int cb200_bond_stage_times(cb200_bond* B, const void* phi, int32_t reps, double* ms3, double* flops3) {
  if (!B || !phi || !ms3 || reps < 1) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    Ctx* c = &B->ctx->c;
    const size_t es = B->cplx ? 16 : 8;
    const int64_t n = std::max<int64_t>(B->b.philay.total, 1);
    Buf x(c->be, (int64_t)es * n, true), y(c->be, (int64_t)es * n, true);
    import_phi(c, B->cplx, B->S, B->l, B->r, B->b.philay, phi, x.p);
    B->b.apply(x.p, y.p);
    dev_sync(c->be);
    B->b.stage_times(x.p, y.p, reps, ms3);
    if (flops3) { flops3[0] = B->b.g1.flops; flops3[1] = B->b.mix.flops; flops3[2] = B->b.g2.flops; }
  });
}

int cb200_bond_stage_bytes(const cb200_bond* B, double* bytes3) {
  if (!B || !bytes3) return CB200_ERR_INVALID;
  // algorithmic HBM traffic of the three stages of one product: every operand read once, every result written once
  const double es = B->cplx ? 16.0 : 8.0;
  const double phi = (double)B->b.philay.total, t1 = (double)B->b.t1_elems, t2 = (double)B->b.t2_elems;
  bytes3[0] = es * ((double)B->L.lay.total + phi + t1);
  bytes3[1] = es * (t1 + t2);
  bytes3[2] = es * (t2 + (double)B->R.lay.total + phi);
  return CB200_OK;
}

int cb200_bond_update_timed(cb200_bond* B, const void* phi, int32_t dir, int32_t maxdim, double cutoff, int32_t niter,
                            double* energy, int32_t* m_out, int32_t* jacobi_sweeps, double* seconds4) {
  if (!B || !phi || !seconds4) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    Ctx* c = &B->ctx->c;
    const size_t es = B->cplx ? 16 : 8;
    const int64_t n = std::max<int64_t>(B->b.philay.total, 1);
    Buf x(c->be, (int64_t)es * n, true);
    import_phi(c, B->cplx, B->S, B->l, B->r, B->b.philay, phi, x.p);
    dev_sync(c->be);
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    c->tm = Timers{};
    double t0 = now();
    DavidsonResult dr = davidson(B->b, x.p, niter);
    dev_sync(c->be);
    const double t_dav = now() - t0;
    t0 = now();
    SplitResult sr = split_bond(c, B->cplx, &B->S, B->b.philay, x.p, dir, maxdim, 1, cutoff);
    dev_sync(c->be);
    const double t_trunc = now() - t0;
    t0 = now();
    DevEnv E;
    if (dir == 0) env_update_left(c, B->cplx, &B->S, B->L, sr.A, B->W1, E);
    else env_update_right(c, B->cplx, &B->S, B->R, sr.B, B->W2, E);
    dev_sync(c->be);
    const double t_env = now() - t0;
    seconds4[0] = c->tm.matvec; seconds4[1] = t_dav - c->tm.matvec; seconds4[2] = t_trunc; seconds4[3] = t_env;
    if (energy) *energy = dr.eigenvalue;
    if (m_out) *m_out = (int32_t)sr.mid.total();
    if (jacobi_sweeps) *jacobi_sweeps = sr.sweeps;
  });
}

int cb200_bond_davidson(cb200_bond* B, void* phi, int32_t niter, double* eigenvalue, int32_t* nmatvec) {
  if (!B || !phi || !eigenvalue) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    Ctx* c = &B->ctx->c;
    const size_t es = B->cplx ? 16 : 8;
    const int64_t n = std::max<int64_t>(B->b.philay.total, 1);
    Buf x(c->be, (int64_t)es * n, true);
    import_phi(c, B->cplx, B->S, B->l, B->r, B->b.philay, phi, x.p);
    DavidsonResult dr = davidson(B->b, x.p, niter);
    *eigenvalue = dr.eigenvalue;
    if (nmatvec) *nmatvec = dr.nmatvec;
    export_phi(c, B->cplx, B->S, B->l, B->r, B->b.philay, x.p, phi);
  });
}

int cb200_bond_truncate(cb200_bond* B, const void* phi, int32_t dir, int32_t maxdim, int32_t mindim, double cutoff, int32_t* m,
                        int32_t* charges_mid, double* truncerr, double* spectrum, void* A, void* Bout) {
  if (!B || !phi || !m) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    Ctx* c = &B->ctx->c;
    const size_t es = B->cplx ? 16 : 8;
    const int64_t n = std::max<int64_t>(B->b.philay.total, 1);
    Buf x(c->be, (int64_t)es * n, true);
    import_phi(c, B->cplx, B->S, B->l, B->r, B->b.philay, phi, x.p);
    SplitResult sr = split_bond(c, B->cplx, &B->S, B->b.philay, x.p, dir, maxdim, mindim, cutoff);
    *m = (int32_t)sr.mid.total();
    IndexMap mid = IndexMap::from_grading(sr.mid);
    if (charges_mid)
      for (int64_t i = 0; i < sr.mid.total(); ++i) charges_mid[i] = mid.charge_ext[i];
    if (truncerr) *truncerr = sr.truncerr;
    if (spectrum)
      for (size_t i = 0; i < sr.spectrum.size(); ++i) spectrum[i] = sr.spectrum[i];
    if (A) export_site3(c, B->cplx, B->S, B->l, mid, sr.A, A);
    if (Bout) export_site3(c, B->cplx, B->S, mid, B->r, sr.B, Bout);
  });
}

int cb200_bond_env_update(cb200_bond* B, int32_t dir, int32_t m, const int32_t* charges_mid, const void* AorB, void* out) {
  if (!B || !AorB || !out) return CB200_ERR_INVALID;
  return guard(B->ctx, [&] {
    Ctx* c = &B->ctx->c;
    if ((dir != 0 && dir != 1) || m < 1) fail(CB200_ERR_INVALID, "env_update: dir must be 0 / 1 and m >= 1");
    IndexMap mid = IndexMap::from_labels(B->nq, m, charges_mid);
    DevSite3 T;
    DevEnv E;
    if (dir == 0) {
      import_site3(c, B->cplx, B->S, B->l, mid, AorB, T);
      env_update_left(c, B->cplx, &B->S, B->L, T, B->W1, E);
    } else {
      import_site3(c, B->cplx, B->S, mid, B->r, AorB, T);
      env_update_right(c, B->cplx, &B->S, B->R, T, B->W2, E);
    }
    export_env(c, B->cplx, mid, B->km, E, out);
  });
}

// ---------------------------------------------------------------------------------------------- kernel-level
int cb200_k_gemm(cb200_ctx* ctx, int32_t dtype, int32_t trans_a, int32_t trans_b, int32_t conj_a, int32_t accumulate, int32_t M,
                 int32_t N, int32_t nseg, const int32_t* K, const void* A, int64_t a_elems, const int64_t* a_off, int64_t lda,
                 const void* Bm, int64_t b_elems, const int64_t* b_off, int64_t ldb, void* C, int64_t c_elems, int64_t c_off,
                 int64_t ldc, int32_t reps, double* ms_mean) {
  if (!ctx) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    Ctx* c = &ctx->c;
    const bool cx = dtype == 1;
    const size_t es = cx ? 16 : 8;
    Buf dA(c->be, (int64_t)es * a_elems), dB(c->be, (int64_t)es * b_elems), dC(c->be, (int64_t)es * c_elems);
    dev_h2d(c->be, dA.p, A, es * a_elems);
    dev_h2d(c->be, dB.p, Bm, es * b_elems);
    dev_h2d(c->be, dC.p, C, es * c_elems);
    std::vector<GemmTask> tasks(1);
    std::vector<GemmSeg> segs(nseg);
    GemmTask& t = tasks[0];
    t = GemmTask{};
    t.c_off = c_off; t.c_off2 = c_off; t.ldc = ldc; t.M = M; t.N = N; t.n_split = N;
    t.seg_begin = 0; t.seg_count = nseg;
    t.flags = (trans_a ? GEMM_TRANS_A : 0) | (trans_b ? GEMM_TRANS_B : 0) | (conj_a ? GEMM_CONJ_A : 0) | (accumulate ? GEMM_ACCUM : 0);
    // tile mode of this test hook: CB200_K_GEMM_MODE = 0 wide (default), 1 narrow, 2 narrow + 3M, 3 mid (96) + 3M (3M: complex only)
    const char* me = getenv("CB200_K_GEMM_MODE");
    int mode = me ? atoi(me) : GEMM_TILE_WIDE;
    if (!cx && (mode == GEMM_TILE_NARROW_3M || mode == GEMM_TILE_MID_3M)) mode = GEMM_TILE_WIDE;
    const int tm = gemm_tile_m(cx), tn = gemm_tile_n(cx, mode);
    t.tile_begin = 0; t.tiles_m = (M + tm - 1) / tm;
    const int tiles = t.tiles_m * ((N + tn - 1) / tn);
    for (int s = 0; s < nseg; ++s) { segs[s] = GemmSeg{}; segs[s].a_off = a_off[s]; segs[s].b_off = b_off[s]; segs[s].lda = lda; segs[s].ldb = ldb; segs[s].K = K[s]; }
    DevList<GemmTask> dt; DevList<GemmSeg> dsg;
    dt.upload(c->be, tasks); dsg.upload(c->be, segs);
    launch_gemm(c->be, cx, trans_a != 0, trans_b != 0, mode, dA.p, dB.p, dC.p, dt.ptr(), 1, dsg.ptr(), tiles);
    dev_d2h(c->be, C, dC.p, es * c_elems);
    if (reps > 0 && ms_mean) {
      dev_event_record(c->be, 0);
      for (int i = 0; i < reps; ++i) launch_gemm(c->be, cx, trans_a != 0, trans_b != 0, mode, dA.p, dB.p, dC.p, dt.ptr(), 1, dsg.ptr(), tiles);
      dev_event_record(c->be, 1);
      *ms_mean = dev_event_elapsed_ms(c->be, 0, 1) / reps;
    }
  });
}

int cb200_k_jacobi_eig(cb200_ctx* ctx, int32_t dtype, int32_t npairs, const void* G, void* Q, double tol, int64_t* rotations) {
  if (!ctx) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    Ctx* c = &ctx->c;
    const bool cx = dtype == 1;
    const size_t bytes = (size_t)(cx ? 16 : 8) * 4 * JB * JB * npairs;
    Buf dG(c->be, (int64_t)bytes), dQ(c->be, (int64_t)bytes), rot(c->be, 8, true);
    dev_h2d(c->be, dG.p, G, bytes);
    launch_jacobi_eig(c->be, cx, dG.p, dQ.p, npairs, tol, 0.0, 0.0, 0, (long long*)rot.p);
    dev_d2h(c->be, Q, dQ.p, bytes);
    long long r = 0;
    dev_d2h(c->be, &r, rot.p, 8);
    if (rotations) *rotations = r;
  });
}

int cb200_k_svd(cb200_ctx* ctx, int32_t dtype, int64_t rows, int64_t cols, void* X, void* V, double* w, int32_t* sweeps) {
  if (!ctx) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    Ctx* c = &ctx->c;
    const bool cx = dtype == 1;
    const size_t es = cx ? 16 : 8;
    Buf dX(c->be, (int64_t)es * std::max<int64_t>(rows * cols, 1));
    dev_h2d(c->be, dX.p, X, es * rows * cols);
    JacobiOut jo = jacobi_svd(c, cx, dX.p, rows, cols);
    // return the `cols` heaviest slots, in descending order of weight
    std::vector<int64_t> order(jo.w.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int64_t)i;
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return jo.w[a] > jo.w[b]; });
    std::vector<char> hx(es * (size_t)jo.ldx * jo.w.size()), hv(es * (size_t)jo.ldv * jo.w.size());
    dev_d2h(c->be, hx.data(), jo.XV.p, hx.size());
    dev_d2h(c->be, hv.data(), jo.V.p, hv.size());
    for (int64_t k = 0; k < cols; ++k) {
      const int64_t s = order[k];
      w[k] = jo.w[s];
      std::memcpy((char*)X + es * (size_t)(k * rows), hx.data() + es * (size_t)(s * jo.ldx), es * (size_t)rows);
      std::memcpy((char*)V + es * (size_t)(k * cols), hv.data() + es * (size_t)(s * jo.ldv), es * (size_t)cols);
    }
    if (sweeps) *sweeps = jo.sweeps;
  });
}

int cb200_k_heig(cb200_ctx* ctx, int32_t dtype, int64_t n, const void* A, int64_t m, double* w, void* V, int32_t* ns_iters) {
  if (!ctx) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    Ctx* c = &ctx->c;
    const bool cx = dtype == 1;
    const size_t es = cx ? 16 : 8;
    if (n < 1 || m < 0 || m > n) fail(CB200_ERR_INVALID, "k_heig: bad sizes");
    Buf dA(c->be, (int64_t)es * n * n);
    dev_h2d(c->be, dA.p, A, es * (size_t)n * n);
    Heig h;
    if (!heig_values(c, cx, std::move(dA), n, h)) fail(CB200_ERR_INVALID, "k_heig: this backend has no tridiagonalisation kernel");
    for (int64_t i = 0; i < n; ++i) w[i] = h.w[i];
    if (m > 0) {
      Buf dV = heig_vectors(c, h, m);
      dev_d2h(c->be, V, dV.p, es * (size_t)n * m);
    }
    if (ns_iters) *ns_iters = h.ns_iters;
  });
}

int cb200_k_heig_batched(cb200_ctx* ctx, int32_t dtype, int32_t nprob, const int64_t* n, const void* const* A, const int64_t* m,
                         double* const* w, void* const* V) {
  if (!ctx || !n || !A || !m || !w || !V || nprob < 1) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    Ctx* c = &ctx->c;
    const bool cx = dtype == 1;
    const size_t es = cx ? 16 : 8;
    std::vector<Buf> dA(nprob);
    std::vector<int64_t> ns(nprob);
    std::vector<Heig> hs(nprob);
    std::vector<Heig*> hp(nprob);
    for (int k = 0; k < nprob; ++k) {
      if (n[k] < 1 || m[k] < 0 || m[k] > n[k]) fail(CB200_ERR_INVALID, "k_heig_batched: bad sizes");
      dA[k] = Buf(c->be, (int64_t)es * n[k] * n[k]);
      dev_h2d(c->be, dA[k].p, A[k], es * (size_t)n[k] * n[k]);
      ns[k] = n[k];
      hp[k] = &hs[k];
    }
    if (!heig_values_batched(c, cx, dA, ns, hp)) fail(CB200_ERR_INVALID, "k_heig_batched: this backend has no tridiagonalisation kernel");
    for (int k = 0; k < nprob; ++k) {
      for (int64_t i = 0; i < n[k]; ++i) w[k][i] = hs[k].w[i];
      if (m[k] > 0) {
        Buf dV = heig_vectors(c, hs[k], m[k]);
        dev_d2h(c->be, V[k], dV.p, es * (size_t)n[k] * m[k]);
      }
    }
  });
}

int cb200_k_dots(cb200_ctx* ctx, int32_t dtype, int32_t nv, int64_t n, const void* x, const void* y, double* out) {
  if (!ctx) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    Ctx* c = &ctx->c;
    const bool cx = dtype == 1;
    const size_t es = cx ? 16 : 8;
    Buf dx(c->be, (int64_t)es * n * nv), dy(c->be, (int64_t)es * n);
    dev_h2d(c->be, dx.p, x, es * n * nv);
    dev_h2d(c->be, dy.p, y, es * n);
    std::vector<const void*> xs(nv);
    for (int j = 0; j < nv; ++j) xs[j] = (const char*)dx.p + es * (size_t)n * j;
    multi_dot(c->be, cx, nv, xs.data(), dy.p, n, out);
  });
}

int cb200_k_lincomb(cb200_ctx* ctx, int32_t dtype, int32_t nv, int64_t n, const void* x, const double* coef, double beta_re,
                    double beta_im, void* dst) {
  if (!ctx) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    Ctx* c = &ctx->c;
    const bool cx = dtype == 1;
    const size_t es = cx ? 16 : 8;
    Buf dx(c->be, (int64_t)es * n * nv), dd(c->be, (int64_t)es * n);
    dev_h2d(c->be, dx.p, x, es * n * nv);
    dev_h2d(c->be, dd.p, dst, es * n);
    std::vector<const void*> xs(nv);
    for (int j = 0; j < nv; ++j) xs[j] = (const char*)dx.p + es * (size_t)n * j;
    lincomb(c->be, cx, nv, xs.data(), coef, beta_re, beta_im, dd.p, n);
    dev_d2h(c->be, dst, dd.p, es * n);
  });
}

int cb200_k_level1_timed(cb200_ctx* ctx, int32_t dtype, int32_t nv, int64_t n, int32_t reps, double* ms3) {
  if (!ctx || !ms3 || nv < 1 || nv > MAX_LC || n < 1 || reps < 1) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    Ctx* c = &ctx->c;
    const bool cx = dtype == 1;
    const size_t es = cx ? 16 : 8;
    Buf dx(c->be, (int64_t)es * n * nv, true), dy(c->be, (int64_t)es * n, true);
    std::vector<const void*> xs(nv);
    for (int j = 0; j < nv; ++j) xs[j] = (const char*)dx.p + es * (size_t)n * j;
    std::vector<double> out(2 * (size_t)nv), coef(2 * (size_t)nv, 0.0);
    for (int j = 0; j < nv; ++j) coef[2 * j] = 1e-3;
    multi_dot(c->be, cx, nv, xs.data(), dy.p, n, out.data());                     // warm-up of all three
    lincomb(c->be, cx, nv, xs.data(), coef.data(), 1.0, 0.0, dy.p, n);
    dev_d2d(c->be, dy.p, dx.p, es * n);
    dev_event_record(c->be, 0);
    for (int r = 0; r < reps; ++r) multi_dot(c->be, cx, nv, xs.data(), dy.p, n, out.data());
    dev_event_record(c->be, 1);
    ms3[0] = dev_event_elapsed_ms(c->be, 0, 1) / reps;
    dev_event_record(c->be, 0);
    for (int r = 0; r < reps; ++r) lincomb(c->be, cx, nv, xs.data(), coef.data(), 1.0, 0.0, dy.p, n);
    dev_event_record(c->be, 1);
    ms3[1] = dev_event_elapsed_ms(c->be, 0, 1) / reps;
    dev_event_record(c->be, 0);
    for (int r = 0; r < reps; ++r) dev_d2d(c->be, dy.p, dx.p, es * n);
    dev_event_record(c->be, 1);
    ms3[2] = dev_event_elapsed_ms(c->be, 0, 1) / reps;
  });
}

int cb200_k_gemm_peak(cb200_ctx* ctx, int32_t dtype, int32_t n, int32_t reps, double* tflops) {
  if (!ctx || !tflops) return CB200_ERR_INVALID;
  return guard(ctx, [&] {
    Ctx* c = &ctx->c;
    const bool cx = dtype == 1;
    const size_t es = cx ? 16 : 8;
    const int64_t ne = (int64_t)n * n;
    Buf dA(c->be, (int64_t)es * ne), dB(c->be, (int64_t)es * ne), dC(c->be, (int64_t)es * ne, true);
    std::vector<double> h((size_t)ne * (cx ? 2 : 1));
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
    dev_h2d(c->be, dA.p, h.data(), es * ne);
    dev_h2d(c->be, dB.p, h.data(), es * ne);
    std::vector<GemmTask> tasks(1);
    std::vector<GemmSeg> segs(1);
    GemmTask& t = tasks[0];
    t = GemmTask{};
    t.ldc = n; t.M = n; t.N = n; t.n_split = n; t.seg_count = 1;
    const int tm = gemm_tile_m(cx), tn = gemm_tile_n(cx, false);
    t.tiles_m = (n + tm - 1) / tm;
    const int tiles = t.tiles_m * ((n + tn - 1) / tn);
    segs[0] = GemmSeg{};
    segs[0].lda = n; segs[0].ldb = n; segs[0].K = n;
    DevList<GemmTask> dt; DevList<GemmSeg> dsg;
    dt.upload(c->be, tasks); dsg.upload(c->be, segs);
    for (int i = 0; i < 3; ++i) launch_gemm(c->be, cx, false, false, false, dA.p, dB.p, dC.p, dt.ptr(), 1, dsg.ptr(), tiles);
    dev_sync(c->be);
    dev_event_record(c->be, 0);
    for (int i = 0; i < reps; ++i) launch_gemm(c->be, cx, false, false, false, dA.p, dB.p, dC.p, dt.ptr(), 1, dsg.ptr(), tiles);
    dev_event_record(c->be, 1);
    const double ms = dev_event_elapsed_ms(c->be, 0, 1) / reps;
    *tflops = (cx ? 8.0 : 2.0) * (double)n * n * n / (ms * 1e-3) / 1e12;
  });
}

int cb200_k_fp64_peak(cb200_ctx* ctx, int32_t which, int32_t iters, double* tflops) {
  if (!ctx || !tflops) return CB200_ERR_INVALID;
  return guard(ctx, [&] { *tflops = fp64_pipe_peak(ctx->c.be, which, iters); });
}

}  // extern "C"
