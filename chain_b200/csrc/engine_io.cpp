// Import/export between the dense external tensors that cross the C ABI (ITensor's column-major storage,
// arbitrary charge order per index) and the internal block layouts; and the sweep driver itensor::dmrg().
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "engine.h"

namespace cb200 {

static size_t esz(bool cplx) { return cplx ? 16 : 8; }
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void check_dev(Ctx* c) {
  const char* e = dev_last_error(c->be);
  if (e) fail(-3, std::string("device error: ") + e);
}

// read element i of an external array of given dtype as complex
static inline cplx_t rd(const void* p, bool cplx, int64_t i) {
  if (cplx) return cplx_t(((const double*)p)[2 * i], ((const double*)p)[2 * i + 1]);
  return cplx_t(((const double*)p)[i], 0.0);
}
static inline void wr(void* p, bool cplx, int64_t i, cplx_t v) {
  if (cplx) { ((double*)p)[2 * i] = v.real(); ((double*)p)[2 * i + 1] = v.imag(); }
  else ((double*)p)[i] = v.real();
}

HostW import_w(int nq, int d, const SiteInfo& S, const IndexMap& kl, const IndexMap& kr, bool cplx, const void* data) {
  HostW W;
  W.KL = kl.g; W.KR = kr.g; W.d = d;
  const int64_t nl = kl.g.total(), nr = kr.g.total();
  W.w.assign((size_t)nl * d * d * nr, cplx_t(0, 0));
  W.qa_l.resize(nl); W.qa_r.resize(nr);
  for (int64_t a = 0; a < nl; ++a) W.qa_l[kl.ext2int[a]] = kl.charge_ext[a];
  for (int64_t a = 0; a < nr; ++a) W.qa_r[kr.ext2int[a]] = kr.charge_ext[a];
  for (int64_t ap = 0; ap < nr; ++ap)
    for (int s = 0; s < d; ++s)
      for (int t = 0; t < d; ++t)
        for (int64_t a = 0; a < nl; ++a) {
          const cplx_t v = rd(data, cplx, a + nl * (t + (int64_t)d * (s + (int64_t)d * ap)));
          if (v == cplx_t(0, 0)) continue;
          const int64_t ai = kl.ext2int[a], api = kr.ext2int[ap];
          const int ti = (int)S.map.ext2int[t], si = (int)S.map.ext2int[s];
          // charge rule q(a') = q(a) + q(s_out) - q(s_in)
          if (qadd(W.qa_l[ai], S.q_int[ti], nq) != qadd(W.qa_r[api], S.q_int[si], nq))
            fail(-2, "MPO tensor has a non-zero element that violates q(a)+q(s_out) = q(s_in)+q(a')");
          W.w[ai + nl * (ti + (int64_t)d * (si + (int64_t)d * api))] = v;
        }
  return W;
}

// site tensors with at least this many dense elements are (un)packed on the device (below it the host loop is cheaper than a launch)
// (CB200_PACK_MIN overrides the threshold; read per call so that a test can exercise both paths on the same tensors)
static int64_t site3_pack_min() { const char* e = std::getenv("CB200_PACK_MIN"); return e ? std::atoll(e) : (int64_t)4096; }
// generic 3-index transfer: ext[xe + Dx*(se + d*ye)] <-> block (qx,qs) element xi_loc + Dx_q*(s_loc + ds*y_loc)
template <class F>
static void for_site3(const SiteInfo& S, const IndexMap& l, const IndexMap& r, const Site3Layout& lay, F f) {
  const int nq = S.nq, d = S.d;
  const int64_t Dx = l.g.total();
  for (int qx = 0; qx < nq; ++qx)
    for (int qs = 0; qs < nq; ++qs) {
      const int qy = qadd(qx, qs, nq);
      const int64_t dx = l.g.dim[qx], dsq = S.map.g.dim[qs], dy = r.g.dim[qy];
      const int64_t bo = lay.o(qx, qs);
      for (int64_t y = 0; y < dy; ++y) {
        const int64_t ye = r.int2ext[r.g.off(qy) + y];
        for (int64_t s = 0; s < dsq; ++s) {
          const int64_t se = S.map.int2ext[S.map.g.off(qs) + s];
          for (int64_t x = 0; x < dx; ++x) {
            const int64_t xe = l.int2ext[l.g.off(qx) + x];
            f(xe + Dx * (se + (int64_t)d * ye), bo + x + dx * (s + dsq * y));
          }
        }
      }
    }
}


// Device-side (un)packing of a dense [x, s, y] site tensor: one strided-copy task per (sector of x, site value, sector of y),
// sectors = maximal runs of equal charge label (internal order is a stable sort by charge, so a run is a contiguous internal
// range).  Returns false when the labels are so interleaved that the task list would be longer than the host loop is slow.
static bool site3_copy_tasks(const SiteInfo& S, const IndexMap& l, const IndexMap& r, const Site3Layout& lay, bool to_internal,
                             std::vector<CopyTask>& tasks) {
  const int nq = S.nq, d = S.d;
  const int64_t Dx = l.g.total();
  struct Run { int64_t start, len; int q; };
  auto runs = [](const IndexMap& m) {
    std::vector<Run> v;
    const int64_t n = (int64_t)m.charge_ext.size();
    for (int64_t i = 0; i < n;) {
      int64_t j = i;
      while (j < n && m.charge_ext[j] == m.charge_ext[i]) ++j;
      v.push_back({i, j - i, m.charge_ext[i]});
      i = j;
    }
    return v;
  };
  const std::vector<Run> rl = runs(l), rr = runs(r);
  if ((int64_t)rl.size() * (int64_t)rr.size() * d > 16384) return false;
  for (const Run& uy : rr)
    for (int se = 0; se < d; ++se) {
      const int si = (int)S.map.ext2int[se], qs = S.q_int[si];
      const int64_t s_loc = S.loc_int[si], dsq = S.map.g.dim[qs];
      for (const Run& ux : rl) {
        if (qadd(ux.q, qs, nq) != uy.q) continue;
        const int64_t dx = l.g.dim[ux.q];
        const int64_t x_loc = l.ext2int[ux.start] - l.g.off(ux.q), y_loc = r.ext2int[uy.start] - r.g.off(uy.q);
        const int64_t eoff = ux.start + Dx * (se + (int64_t)d * uy.start), es1 = Dx * d;
        const int64_t ioff = lay.o(ux.q, qs) + x_loc + dx * (s_loc + dsq * y_loc), is1 = dx * dsq;
        CopyTask t{};
        t.n0 = (int32_t)ux.len; t.n1 = (int32_t)uy.len; t.n2 = 1;
        if (to_internal) { t.src_off = eoff; t.dst_off = ioff; t.ss0 = 1; t.ss1 = es1; t.ds0 = 1; t.ds1 = is1; }
        else { t.src_off = ioff; t.dst_off = eoff; t.ss0 = 1; t.ss1 = is1; t.ds0 = 1; t.ds1 = es1; }
        tasks.push_back(t);
      }
    }
  return true;
}

static void import_site3_t(Ctx* c, bool cplx, bool src_cplx, const SiteInfo& S, const IndexMap& l, const IndexMap& r, const void* ext,
                           DevSite3& out) {
  out.lay.build(l.g, r.g, &S);
  const int64_t tot = l.g.total() * S.d * r.g.total();
  std::vector<CopyTask> tasks;
  if (cplx == src_cplx && tot >= site3_pack_min() && site3_copy_tasks(S, l, r, out.lay, true, tasks)) {
    // the dense array crosses PCIe as it is and the block gather is one launch (no host pass over the data)
    out.buf = Buf(c->be, (int64_t)esz(cplx) * std::max<int64_t>(out.lay.total, 1));
    Buf stage(c->be, (int64_t)esz(cplx) * tot);
    dev_h2d(c->be, stage.p, ext, esz(cplx) * (size_t)tot);
    run_copy_tasks(c, cplx, tasks, stage.p, out.buf.p);
    return;
  }
  std::vector<char> h(esz(cplx) * (size_t)std::max<int64_t>(out.lay.total, 1), 0);
  for_site3(S, l, r, out.lay, [&](int64_t e, int64_t i) { wr(h.data(), cplx, i, rd(ext, src_cplx, e)); });
  out.buf = Buf(c->be, (int64_t)h.size());
  dev_h2d(c->be, out.buf.p, h.data(), h.size());
}
void import_site3(Ctx* c, bool cplx, const SiteInfo& S, const IndexMap& l, const IndexMap& r, const void* ext, DevSite3& out) {
  import_site3_t(c, cplx, cplx, S, l, r, ext, out);
}

// ---- block-packed ("QDense order") MPS site tensors: cb200_mps.layout = 1 --------------------------------------------------------
// Sectors = maximal runs of equal charge label on l, s and r; the stored blocks are the zero-flux sector triples q(l)+q(s) = q(r) in
// increasing block number i_l + n_l*(i_s + n_s*i_r); each block dense column-major [len_l, len_s, len_r]; no gaps.  One strided-copy
// task per (block, site value inside the block's site sector).  tasks == nullptr: only the element count is returned.
int64_t site3_block_tasks(const SiteInfo& S, const IndexMap& l, const IndexMap& r, const Site3Layout* lay, bool to_internal,
                          std::vector<CopyTask>* tasks) {
  const int nq = S.nq;
  struct Run { int64_t start, len; int q; };
  auto runs = [](const IndexMap& m) {
    std::vector<Run> v;
    const int64_t n = (int64_t)m.charge_ext.size();
    for (int64_t i = 0; i < n;) {
      int64_t j = i;
      while (j < n && m.charge_ext[j] == m.charge_ext[i]) ++j;
      v.push_back({i, j - i, m.charge_ext[i]});
      i = j;
    }
    return v;
  };
  const std::vector<Run> rl = runs(l), rs = runs(S.map), rr = runs(r);
  int64_t boff = 0;
  for (const Run& uy : rr)
    for (const Run& us : rs)
      for (const Run& ux : rl) {
        if (qadd(ux.q, us.q, nq) != uy.q) continue;
        if (tasks) {
          const int64_t dx = l.g.dim[ux.q];
          const int64_t x_loc = l.ext2int[ux.start] - l.g.off(ux.q), y_loc = r.ext2int[uy.start] - r.g.off(uy.q);
          for (int64_t b1 = 0; b1 < us.len; ++b1) {
            const int si = (int)S.map.ext2int[us.start + b1], qs = S.q_int[si];
            const int64_t s_loc = S.loc_int[si], dsq = S.map.g.dim[qs];
            const int64_t eoff = boff + ux.len * b1, es1 = ux.len * us.len;
            const int64_t ioff = lay->o(ux.q, qs) + x_loc + dx * (s_loc + dsq * y_loc), is1 = dx * dsq;
            CopyTask t{};
            t.n0 = (int32_t)ux.len; t.n1 = (int32_t)uy.len; t.n2 = 1;
            if (to_internal) { t.src_off = eoff; t.dst_off = ioff; t.ss0 = 1; t.ss1 = es1; t.ds0 = 1; t.ds1 = is1; }
            else { t.src_off = ioff; t.dst_off = eoff; t.ss0 = 1; t.ss1 = is1; t.ds0 = 1; t.ds1 = es1; }
            tasks->push_back(t);
          }
        }
        boff += ux.len * us.len * uy.len;
      }
  return boff;
}
static void import_site3_blocks(Ctx* c, bool cplx, bool src_cplx, const SiteInfo& S, const IndexMap& l, const IndexMap& r, const void* ext,
                                DevSite3& out) {
  out.lay.build(l.g, r.g, &S);
  std::vector<CopyTask> tasks;
  const int64_t tot = site3_block_tasks(S, l, r, &out.lay, true, &tasks);
  if (tot != out.lay.total) fail(-2, "block-packed site tensor: element count does not match the charge structure");
  out.buf = Buf(c->be, (int64_t)esz(cplx) * std::max<int64_t>(tot, 1));
  if (tot == 0) return;
  Buf stage(c->be, (int64_t)esz(cplx) * tot);
  if (cplx == src_cplx) {
    dev_h2d(c->be, stage.p, ext, esz(cplx) * (size_t)tot);
  } else {   // real data into a complex run (a real start state under a complex Hamiltonian): widen on the host
    std::vector<char> h(esz(cplx) * (size_t)tot);
    for (int64_t i = 0; i < tot; ++i) wr(h.data(), cplx, i, rd(ext, src_cplx, i));
    dev_h2d(c->be, stage.p, h.data(), h.size());
  }
  run_copy_tasks(c, cplx, tasks, stage.p, out.buf.p);
}
// site j of an MPS that crossed the boundary in either layout
static void import_mps_site(Ctx* c, bool cplx, bool src_cplx, const SiteInfo& S, const HostMPS& m, int j, DevSite3& out) {
  if (m.packed) import_site3_blocks(c, cplx, src_cplx, S, m.link[j], m.link[j + 1], m.site[j], out);
  else import_site3_t(c, cplx, src_cplx, S, m.link[j], m.link[j + 1], m.site[j], out);
}
int64_t result_site_blocks_elems(const DmrgResult& r, int site) { return r.dev[site].lay.total; }
void result_site_export_blocks(const DmrgResult& r, int site, void* out) {
  Ctx* c = r.ctx;
  const DevSite3& in = r.dev[site];
  std::vector<CopyTask> tasks;
  const int64_t tot = site3_block_tasks(r.S, r.psi.link[site], r.psi.link[site + 1], &in.lay, false, &tasks);
  if (tot != in.lay.total) fail(-2, "block-packed site tensor: element count does not match the charge structure");
  if (tot == 0) return;
  Buf stage(c->be, (int64_t)esz(r.cplx) * tot);
  run_copy_tasks(c, r.cplx, tasks, in.buf.p, stage.p);
  dev_d2h(c->be, out, stage.p, esz(r.cplx) * (size_t)tot);
  check_dev(c);
}
void export_site3(Ctx* c, bool cplx, const SiteInfo& S, const IndexMap& l, const IndexMap& r, const DevSite3& in, void* ext) {
  {
    const int64_t tot = l.g.total() * S.d * r.g.total();
    std::vector<CopyTask> tasks;
    if (tot >= site3_pack_min() && site3_copy_tasks(S, l, r, in.lay, false, tasks)) {
      Buf stage(c->be, (int64_t)esz(cplx) * tot, true);
      run_copy_tasks(c, cplx, tasks, in.buf.p, stage.p);
      dev_d2h(c->be, ext, stage.p, esz(cplx) * (size_t)tot);
      return;
    }
  }
  std::vector<char> h(esz(cplx) * (size_t)std::max<int64_t>(in.lay.total, 1));
  dev_d2h(c->be, h.data(), in.buf.p, esz(cplx) * (size_t)in.lay.total);
  const int64_t tot = l.g.total() * S.d * r.g.total();
  std::memset(ext, 0, esz(cplx) * (size_t)tot);
  for_site3(S, l, r, in.lay, [&](int64_t e, int64_t i) { wr(ext, cplx, e, rd(h.data(), cplx, i)); });
}

template <class F>
static void for_env(const IndexMap& x, const IndexMap& k, const EnvLayout& lay, F f) {
  const int nq = x.g.nq;
  const int64_t Dx = x.g.total(), Kt = k.g.total();
  for (int qxp = 0; qxp < nq; ++qxp)
    for (int qa = 0; qa < nq; ++qa) {
      const int qx = qsub(qxp, qa, nq);
      const int64_t dxp = x.g.dim[qxp], ka = k.g.dim[qa], dx = x.g.dim[qx];
      const int64_t bo = lay.o(qxp, qa);
      for (int64_t xi = 0; xi < dx; ++xi) {
        const int64_t xe = x.int2ext[x.g.off(qx) + xi];
        for (int64_t a = 0; a < ka; ++a) {
          const int64_t ae = k.int2ext[k.g.off(qa) + a];
          for (int64_t xp = 0; xp < dxp; ++xp) {
            const int64_t xpe = x.int2ext[x.g.off(qxp) + xp];
            f(xpe + Dx * (ae + Kt * xe), bo + xp + dxp * (a + ka * xi));
          }
        }
      }
    }
}
void import_env(Ctx* c, bool cplx, const IndexMap& x, const IndexMap& k, const void* ext, DevEnv& out) {
  out.lay.build(x.g, k.g);
  std::vector<char> h(esz(cplx) * (size_t)std::max<int64_t>(out.lay.total, 1), 0);
  for_env(x, k, out.lay, [&](int64_t e, int64_t i) { wr(h.data(), cplx, i, rd(ext, cplx, e)); });
  out.buf = Buf(c->be, (int64_t)h.size());
  dev_h2d(c->be, out.buf.p, h.data(), h.size());
}
void export_env(Ctx* c, bool cplx, const IndexMap& x, const IndexMap& k, const DevEnv& in, void* ext) {
  std::vector<char> h(esz(cplx) * (size_t)std::max<int64_t>(in.lay.total, 1));
  dev_d2h(c->be, h.data(), in.buf.p, esz(cplx) * (size_t)in.lay.total);
  const int64_t tot = x.g.total() * k.g.total() * x.g.total();
  std::memset(ext, 0, esz(cplx) * (size_t)tot);
  for_env(x, k, in.lay, [&](int64_t e, int64_t i) { wr(ext, cplx, e, rd(h.data(), cplx, i)); });
}

template <class F>
static void for_phi(const SiteInfo& S, const IndexMap& l, const IndexMap& r, const PhiLayout& pl, F f) {
  const int nq = S.nq, d = S.d;
  const int64_t Dl = l.g.total();
  for (int ql = 0; ql < nq; ++ql)
    for (int qr = 0; qr < nq; ++qr) {
      const int qd = qsub(qr, ql, nq);
      const int64_t dl = l.g.dim[ql], dr = r.g.dim[qr], ns = S.ns(qd);
      const int64_t bo = pl.o(ql, qr);
      for (int64_t rr = 0; rr < dr; ++rr) {
        const int64_t re = r.int2ext[r.g.off(qr) + rr];
        for (int64_t sr = 0; sr < ns; ++sr) {
          const int64_t s1e = S.map.int2ext[S.srow[qd][sr].first], s2e = S.map.int2ext[S.srow[qd][sr].second];
          for (int64_t ll = 0; ll < dl; ++ll) {
            const int64_t le = l.int2ext[l.g.off(ql) + ll];
            f(le + Dl * (s1e + (int64_t)d * (s2e + (int64_t)d * re)), bo + ll + dl * (sr + ns * rr));
          }
        }
      }
    }
}
// Device-side (un)packing when both link indices arrive already sorted by charge: the dense [l,s1,s2,r] array is
// moved over PCIe as is and the block gather / scatter is one strided-copy kernel (no host pass over the data).
static void phi_copy_tasks(const SiteInfo& S, const IndexMap& l, const IndexMap& r, const PhiLayout& pl, bool to_internal,
                           std::vector<CopyTask>& tasks) {
  const int nq = S.nq, d = S.d;
  const int64_t Dl = l.g.total();
  for (int ql = 0; ql < nq; ++ql)
    for (int qr = 0; qr < nq; ++qr) {
      const int qd = qsub(qr, ql, nq);
      const int64_t dl = l.g.dim[ql], dr = r.g.dim[qr], ns = S.ns(qd);
      if (dl == 0 || dr == 0) continue;
      for (int64_t sr = 0; sr < ns; ++sr) {
        const int64_t s1e = S.map.int2ext[S.srow[qd][sr].first], s2e = S.map.int2ext[S.srow[qd][sr].second];
        CopyTask t{};
        const int64_t eoff = l.g.off(ql) + Dl * (s1e + (int64_t)d * (s2e + (int64_t)d * r.g.off(qr)));
        const int64_t ioff = pl.o(ql, qr) + dl * sr;
        t.n0 = (int32_t)dl; t.n1 = (int32_t)dr; t.n2 = 1;
        if (to_internal) { t.src_off = eoff; t.dst_off = ioff; t.ss0 = 1; t.ss1 = Dl * d * d; t.ds0 = 1; t.ds1 = dl * ns; }
        else { t.src_off = ioff; t.dst_off = eoff; t.ss0 = 1; t.ss1 = dl * ns; t.ds0 = 1; t.ds1 = Dl * d * d; }
        tasks.push_back(t);
      }
    }
}
void run_copy_tasks(Ctx* c, bool cplx, std::vector<CopyTask>& tasks, const void* src, void* dst) {
  if (tasks.empty()) return;
  const int epb = copy_elems_per_block();
  int blocks = 0;
  for (auto& t : tasks) {
    t.blk_begin = blocks;
    blocks += (int)(((int64_t)t.n0 * t.n1 * t.n2 + epb - 1) / epb);
  }
  DevList<CopyTask> d;
  d.upload(c->be, tasks);
  launch_copy(c->be, cplx, src, dst, d.ptr(), d.n, blocks);
}

void import_phi(Ctx* c, bool cplx, const SiteInfo& S, const IndexMap& l, const IndexMap& r, const PhiLayout& pl, const void* ext, void* dev) {
  if (S.nq == 1 && l.identity && r.identity && S.map.identity) {   // dense: the phi layout IS the column-major [l,s1,s2,r] array
    dev_h2d(c->be, dev, ext, esz(cplx) * (size_t)pl.total);
    return;
  }
  const int64_t tot = l.g.total() * S.d * S.d * r.g.total();
  if (l.identity && r.identity) {
    Buf stage(c->be, (int64_t)esz(cplx) * std::max<int64_t>(tot, 1));
    dev_h2d(c->be, stage.p, ext, esz(cplx) * (size_t)tot);
    std::vector<CopyTask> tasks;
    phi_copy_tasks(S, l, r, pl, true, tasks);
    run_copy_tasks(c, cplx, tasks, stage.p, dev);
    return;
  }
  std::vector<char> h(esz(cplx) * (size_t)std::max<int64_t>(pl.total, 1), 0);
  for_phi(S, l, r, pl, [&](int64_t e, int64_t i) { wr(h.data(), cplx, i, rd(ext, cplx, e)); });
  dev_h2d(c->be, dev, h.data(), esz(cplx) * (size_t)pl.total);
}
void export_phi(Ctx* c, bool cplx, const SiteInfo& S, const IndexMap& l, const IndexMap& r, const PhiLayout& pl, const void* dev, void* ext) {
  if (S.nq == 1 && l.identity && r.identity && S.map.identity) {
    dev_d2h(c->be, ext, dev, esz(cplx) * (size_t)pl.total);
    return;
  }
  const int64_t tot = l.g.total() * S.d * S.d * r.g.total();
  if (l.identity && r.identity) {
    Buf stage(c->be, (int64_t)esz(cplx) * std::max<int64_t>(tot, 1), true);
    std::vector<CopyTask> tasks;
    phi_copy_tasks(S, l, r, pl, false, tasks);
    run_copy_tasks(c, cplx, tasks, dev, stage.p);
    dev_d2h(c->be, ext, stage.p, esz(cplx) * (size_t)tot);
    return;
  }
  std::vector<char> h(esz(cplx) * (size_t)std::max<int64_t>(pl.total, 1));
  dev_d2h(c->be, h.data(), dev, esz(cplx) * (size_t)pl.total);
  std::memset(ext, 0, esz(cplx) * (size_t)tot);
  for_phi(S, l, r, pl, [&](int64_t e, int64_t i) { wr(ext, cplx, e, rd(h.data(), cplx, i)); });
}

// ---- block-packed boundary ("QDense order") -----------------------------------------------------------------------------
// ITensor keeps a QN tensor as a list of dense blocks (QDense), never as a dense array with zeros.  The block-packed host layout
// accepted here is that storage: every index is cut into sectors = maximal runs of equal charge label; the blocks are the sector
// tuples (i_l, i_s1, i_s2, i_r) of zero flux, q(l)+q(s1)+q(s2) = q(r), enumerated with i_l fastest and i_r slowest; each block is
// dense column-major [len_l, len_s1, len_s2, len_r]; blocks follow each other without gaps.  Only 1/nq of the dense array crosses
// PCIe, and the (un)packing is one strided-copy launch on the device (internal order is a stable sort by charge, so a sector is a
// contiguous internal range).
namespace {
struct Run { int64_t start, len; int q; };
std::vector<Run> runs_of(const IndexMap& m) {
  std::vector<Run> v;
  const int64_t n = (int64_t)m.charge_ext.size();
  for (int64_t i = 0; i < n;) {
    int64_t j = i;
    while (j < n && m.charge_ext[j] == m.charge_ext[i]) ++j;
    v.push_back({i, j - i, m.charge_ext[i]});
    i = j;
  }
  return v;
}
}  // namespace
int64_t phi_block_tasks(const SiteInfo& S, const IndexMap& l, const IndexMap& r, const PhiLayout& pl, bool to_internal,
                        std::vector<CopyTask>* tasks) {
  const int nq = S.nq, d = S.d;
  const std::vector<Run> rl = runs_of(l), rs = runs_of(S.map), rr = runs_of(r);
  int64_t boff = 0;
  for (const Run& ur : rr)
    for (const Run& u2 : rs)
      for (const Run& u1 : rs)
        for (const Run& ul : rl) {
          if (qsub(qadd(qadd(ul.q, u1.q, nq), u2.q, nq), ur.q, nq) != 0) continue;
          const int64_t bsz = ul.len * u1.len * u2.len * ur.len;
          if (tasks) {
            const int ql = ul.q, qr = ur.q, qd = qsub(qr, ql, nq);
            const int64_t dlq = l.g.dim[ql], ns = S.ns(qd);
            const int64_t l_loc = l.ext2int[ul.start] - l.g.off(ql), r_loc = r.ext2int[ur.start] - r.g.off(qr);
            for (int64_t b2 = 0; b2 < u2.len; ++b2)
              for (int64_t b1 = 0; b1 < u1.len; ++b1) {
                const int s1i = (int)S.map.ext2int[u1.start + b1], s2i = (int)S.map.ext2int[u2.start + b2];
                const int64_t sr = S.srow_index[(size_t)s1i * d + s2i];
                const int64_t eoff = boff + ul.len * (b1 + u1.len * b2);
                const int64_t ioff = pl.o(ql, qr) + l_loc + dlq * (sr + ns * r_loc);
                CopyTask t{};
                t.n0 = (int32_t)ul.len; t.n1 = (int32_t)ur.len; t.n2 = 1;
                const int64_t es1 = ul.len * u1.len * u2.len, is1 = dlq * ns;
                if (to_internal) { t.src_off = eoff; t.dst_off = ioff; t.ss0 = 1; t.ss1 = es1; t.ds0 = 1; t.ds1 = is1; }
                else { t.src_off = ioff; t.dst_off = eoff; t.ss0 = 1; t.ss1 = is1; t.ds0 = 1; t.ds1 = es1; }
                tasks->push_back(t);
              }
          }
          boff += bsz;
        }
  return boff;
}
void import_phi_blocks(Ctx* c, bool cplx, const SiteInfo& S, const IndexMap& l, const IndexMap& r, const PhiLayout& pl, const void* ext, void* dev) {
  std::vector<CopyTask> tasks;
  const int64_t tot = phi_block_tasks(S, l, r, pl, true, &tasks);
  if (tot != pl.total) fail(-2, "block-packed phi: element count does not match the charge structure");
  if (tot == 0) return;
  Buf stage(c->be, (int64_t)esz(cplx) * tot);
  dev_h2d(c->be, stage.p, ext, esz(cplx) * (size_t)tot);
  run_copy_tasks(c, cplx, tasks, stage.p, dev);
}
void export_phi_blocks(Ctx* c, bool cplx, const SiteInfo& S, const IndexMap& l, const IndexMap& r, const PhiLayout& pl, const void* dev, void* ext) {
  std::vector<CopyTask> tasks;
  const int64_t tot = phi_block_tasks(S, l, r, pl, false, &tasks);
  if (tot != pl.total) fail(-2, "block-packed phi: element count does not match the charge structure");
  if (tot == 0) return;
  Buf stage(c->be, (int64_t)esz(cplx) * tot);
  run_copy_tasks(c, cplx, tasks, dev, stage.p);
  dev_d2h(c->be, ext, stage.p, esz(cplx) * (size_t)tot);
}

// ------------------------------------------------------------------------------------------------ overlap environments
namespace {

struct DevMat {   // block-diagonal matrix F[x, j] (same charge on both sides): block q is X[q] x J[q], column-major
  Grading X, J;
  std::vector<int64_t> off;
  int64_t total = 0;
  Buf buf;
  void build(const Grading& x, const Grading& j) {
    X = x; J = j;
    off.assign(x.nq, 0);
    int64_t t = 0;
    for (int q = 0; q < x.nq; ++q) { off[q] = t; t += X.dim[q] * J.dim[q]; }
    total = t;
  }
};

struct PlanBits {
  std::vector<GemmTask> tasks;
  std::vector<GemmSeg> segs;
};

// tiny local copy of the GEMM builder (the one in engine.cpp is file-local)
struct GB {
  std::vector<GemmTask> tasks;
  std::vector<GemmSeg> segs;
  bool cplx;
  GemmTask cur{};
  explicit GB(bool c) : cplx(c) {}
  void begin(int64_t c_off, int64_t ldc, int64_t M, int64_t N, uint32_t flags) {
    cur = GemmTask{};
    cur.c_off = c_off; cur.c_off2 = c_off; cur.ldc = ldc; cur.M = (int32_t)M; cur.N = (int32_t)N; cur.n_split = (int32_t)N;
    cur.seg_begin = (int32_t)segs.size(); cur.flags = flags;
  }
  void seg(int64_t a_off, int64_t lda, int64_t b_off, int64_t ldb, int64_t K) {
    if (K <= 0) return;
    GemmSeg s{};
    s.a_off = a_off; s.b_off = b_off; s.lda = lda; s.ldb = ldb; s.K = (int32_t)K;
    segs.push_back(s);
  }
  void end() {
    cur.seg_count = (int32_t)segs.size() - cur.seg_begin;
    if (cur.M <= 0 || cur.N <= 0) { segs.resize(cur.seg_begin); return; }
    tasks.push_back(cur);
  }
  void run(Ctx* c, bool ta, bool tb, const void* A, const void* B, void* C) {
    if (tasks.empty()) return;
    const int tm = gemm_tile_m(cplx), tn = gemm_tile_n(cplx, GEMM_TILE_WIDE);
    int tiles = 0;
    for (auto& t : tasks) {
      t.tile_begin = tiles;
      t.tiles_m = (t.M + tm - 1) / tm;
      tiles += t.tiles_m * ((t.N + tn - 1) / tn);
    }
    DevList<GemmTask> dt;
    DevList<GemmSeg> dsg;
    dt.upload(c->be, tasks);
    dsg.upload(c->be, segs);
    launch_gemm(c->be, cplx, ta, tb, GEMM_TILE_WIDE, A, B, C, dt.ptr(), dt.n, dsg.ptr(), tiles);
  }
};

// Y[x, s, n] = sum_j F[x, j] P[j, s, n]      (Site3 with left grading F.X, right grading P.Y)
void mat_times_site3(Ctx* c, bool cplx, const SiteInfo* S, const DevMat& F, const DevSite3& P, DevSite3& Y) {
  const int nq = S->nq;
  Y.lay.build(F.X, P.lay.Y, S);
  Y.buf = Buf(c->be, (int64_t)esz(cplx) * std::max<int64_t>(Y.lay.total, 1), true);
  GB gb(cplx);
  for (int q = 0; q < nq; ++q) {
    const int64_t M = F.X.dim[q], K = F.J.dim[q], N = P.lay.ncols(q);
    if (M == 0 || N == 0) continue;
    gb.begin(Y.lay.o(q, 0), M, M, N, 0);
    gb.seg(F.off[q], M, P.lay.o(q, 0), K, K);
    gb.end();
  }
  gb.run(c, false, false, F.buf.p, P.buf.p, Y.buf.p);
}

// Y[n, s, r] = sum_p P[n, s, p] F[r, p]      (Site3 with left grading P.X, right grading F.X)
void site3_times_matT(Ctx* c, bool cplx, const SiteInfo* S, const DevSite3& P, const DevMat& F, DevSite3& Y) {
  const int nq = S->nq;
  const Grading& ds = S->map.g;
  Y.lay.build(P.lay.X, F.X, S);
  Y.buf = Buf(c->be, (int64_t)esz(cplx) * std::max<int64_t>(Y.lay.total, 1), true);
  GB gb(cplx);
  for (int qn = 0; qn < nq; ++qn)
    for (int qs = 0; qs < nq; ++qs) {
      const int qr = qadd(qn, qs, nq);
      const int64_t M = P.lay.X.dim[qn] * ds.dim[qs], N = F.X.dim[qr], K = F.J.dim[qr];
      if (M == 0 || N == 0) continue;
      gb.begin(Y.lay.o(qn, qs), M, M, N, 0);
      gb.seg(P.lay.o(qn, qs), M, F.off[qr], N, K);   // B^T: element (k=p, n=r) at r + N*p
      gb.end();
    }
  gb.run(c, false, true, P.buf.p, F.buf.p, Y.buf.p);
}

// F'[m, n] = sum_{l,s} conj(A[l,s,m]) X[l,s,n]   (A: Dl->Dm, X: Dl->Dn)
void site3H_times_site3(Ctx* c, bool cplx, const SiteInfo* S, const DevSite3& A, const DevSite3& X, DevMat& F) {
  const int nq = S->nq;
  const Grading& ds = S->map.g;
  F.build(A.lay.Y, X.lay.Y);
  F.buf = Buf(c->be, (int64_t)esz(cplx) * std::max<int64_t>(F.total, 1), true);
  GB gb(cplx);
  for (int qm = 0; qm < nq; ++qm) {
    const int64_t M = A.lay.Y.dim[qm], N = X.lay.Y.dim[qm];
    if (M == 0 || N == 0) continue;
    gb.begin(F.off[qm], M, M, N, GEMM_TRANS_A | GEMM_CONJ_A);
    for (int ql = 0; ql < nq; ++ql) {
      const int qs = qsub(qm, ql, nq);
      const int64_t K = A.lay.X.dim[ql] * ds.dim[qs];
      gb.seg(A.lay.o(ql, qs), K, X.lay.o(ql, qs), K, K);
    }
    gb.end();
  }
  gb.run(c, true, false, A.buf.p, X.buf.p, F.buf.p);
}

// F'[m, n] = sum_{s,r} conj(B[m,s,r]) X[n,s,r]   (B: Dm->Dr, X: Dn->Dr)
void site3_times_site3H_right(Ctx* c, bool cplx, const SiteInfo* S, const DevSite3& B, const DevSite3& X, DevMat& F) {
  const int nq = S->nq;
  const Grading& ds = S->map.g;
  F.build(B.lay.X, X.lay.X);
  F.buf = Buf(c->be, (int64_t)esz(cplx) * std::max<int64_t>(F.total, 1), true);
  GB gb(cplx);
  for (int qm = 0; qm < nq; ++qm) {
    const int64_t M = B.lay.X.dim[qm], N = X.lay.X.dim[qm];
    if (M == 0 || N == 0) continue;
    // F'[m,n] = sum_k conj(B[m,k]) X[n,k]: A = B block as M x K (m contiguous) with CONJ_A, B-op = X^T (n contiguous)
    gb.begin(F.off[qm], M, M, N, GEMM_CONJ_A);
    for (int qs = 0; qs < nq; ++qs) {
      const int qr = qadd(qm, qs, nq);
      const int64_t K = ds.dim[qs] * B.lay.Y.dim[qr];
      gb.seg(B.lay.o(qm, qs), M, X.lay.o(qm, qs), N, K);
    }
    gb.end();
  }
  gb.run(c, false, true, B.buf.p, X.buf.p, F.buf.p);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ the sweep driver
std::unique_ptr<DmrgResult> run_dmrg(Ctx* c, const MpoIn& H, const std::vector<HostMPS>& prev, const std::vector<bool>& prev_cplx,
                                     const HostMPS& psi0, bool psi0_cplx, const std::vector<SweepParams>& sweeps, double weight) {
  Backend* be = c->be;
  const int n = H.n, nq = H.nq;
  if (n < 2) fail(-2, "dmrg needs at least 2 sites");
  if (psi0.n != n) fail(-2, "MPS / MPO length mismatch");
  bool cplx = H.cplx || psi0_cplx;
  for (bool b : prev_cplx) cplx = cplx || b;
  c->tm = Timers{};
  double t_other0 = now_s();

  SiteInfo S = SiteInfo::make(nq, H.d, H.site_charge);
  std::vector<HostW> W(n);
  for (int j = 0; j < n; ++j) W[j] = import_w(nq, H.d, S, H.link[j], H.link[j + 1], H.cplx, H.site[j]);

  // current state
  std::vector<IndexMap> link = psi0.link;
  std::vector<DevSite3> A(n);
  for (int j = 0; j < n; ++j) import_mps_site(c, cplx, psi0_cplx, S, psi0, j, A[j]);
  if (link[0].g.total() != 1 || link[n].g.total() != 1) fail(-2, "outer MPS links must have dimension 1");
  // previous states
  const int np = (int)prev.size();
  std::vector<std::vector<DevSite3>> P(np);
  for (int p = 0; p < np; ++p) {
    if (prev[p].n != n) fail(-2, "previous state length mismatch");
    P[p].resize(n);
    for (int j = 0; j < n; ++j) import_mps_site(c, cplx, prev_cplx[p], S, prev[p], j, P[p][j]);
  }

  auto phi_of = [&](int i, PhiLayout& pl, Buf& phi) {   // sites i, i+1 (0-based)
    pl.build(A[i].lay.X, A[i + 1].lay.Y, &S);
    phi = Buf(be, (int64_t)esz(cplx) * std::max<int64_t>(pl.total, 1), true);
    two_site(c, cplx, A[i], A[i + 1], pl, phi.p);
  };

  // psi.position(1) for a non-canonical start: exact (untruncated) splits from the right
  if (psi0.center != 1) {
    for (int i = n - 2; i >= 0; --i) {
      PhiLayout pl; Buf phi;
      phi_of(i, pl, phi);
      // tiny relative cutoff: drops only the numerically-zero directions, so a product state keeps bond dimension 1
      SplitResult sr = split_bond(c, cplx, &S, pl, phi.p, 1, INT64_MAX / 4, 1, 1e-20, true);
      A[i] = std::move(sr.A);
      A[i + 1] = std::move(sr.B);
    }
  }

  // environments: LE[b] covers sites 1..b, RE[b] covers sites b+1..n (cut index b = 0..n)
  std::vector<DevEnv> LE(n + 1), RE(n + 1);
  auto unit_env = [&](DevEnv& e, const Grading& x, const Grading& k) {
    e.lay.build(x, k);
    if (e.lay.total != 1) fail(-2, "boundary environment is not 1x1x1");
    std::vector<char> one(esz(cplx), 0);
    wr(one.data(), cplx, 0, cplx_t(1, 0));
    e.buf = Buf(be, (int64_t)esz(cplx));
    dev_h2d(be, e.buf.p, one.data(), one.size());
  };
  unit_env(LE[0], A[0].lay.X, W[0].KL);
  unit_env(RE[n], A[n - 1].lay.Y, W[n - 1].KR);
  std::vector<std::vector<DevMat>> FL(np), FR(np);
  for (int p = 0; p < np; ++p) { FL[p] = std::vector<DevMat>(n + 1); FR[p] = std::vector<DevMat>(n + 1); }
  for (int p = 0; p < np; ++p) {
    for (DevMat* f : {&FL[p][0], &FR[p][n]}) {
      const Grading& x = (f == &FL[p][0]) ? A[0].lay.X : A[n - 1].lay.Y;
      const Grading& j = (f == &FL[p][0]) ? P[p][0].lay.X : P[p][n - 1].lay.Y;
      f->build(x, j);
      if (f->total != 1) fail(-2, "previous state lives in a different charge sector");
      std::vector<char> one(esz(cplx), 0);
      wr(one.data(), cplx, 0, cplx_t(1, 0));
      f->buf = Buf(be, (int64_t)esz(cplx));
      dev_h2d(be, f->buf.p, one.data(), one.size());
    }
  }
  auto ovl_right = [&](int p, int b) {   // FR[p][b] from FR[p][b+1] and site b+1 (0-based b)
    DevSite3 X;
    site3_times_matT(c, cplx, &S, P[p][b], FR[p][b + 1], X);          // X[n,s,r] = P[n,s,p] F[r,p]
    site3_times_site3H_right(c, cplx, &S, A[b], X, FR[p][b]);
  };
  auto ovl_left = [&](int p, int b) {    // FL[p][b] from FL[p][b-1] and site b (1-based b) = index b-1
    DevSite3 X;
    mat_times_site3(c, cplx, &S, FL[p][b - 1], P[p][b - 1], X);       // X[l,s,n] = F[l,j] P[j,s,n]
    site3H_times_site3(c, cplx, &S, A[b - 1], X, FL[p][b]);
  };
  {
    double t = now_s();
    for (int b = n - 1; b >= 2; --b) {
      env_update_right(c, cplx, &S, RE[b + 1], A[b], W[b], RE[b]);
      for (int p = 0; p < np; ++p) ovl_right(p, b);
    }
    dev_sync(be);
    c->tm.env += now_s() - t;
  }
  c->tm.other += now_s() - t_other0 - c->tm.env - c->tm.trunc;

  auto res = std::make_unique<DmrgResult>();
  res->cplx = cplx;
  double energy = 0;
  std::vector<std::shared_ptr<OmegaMap>> omega(n - 1);   // W_i (x) W_{i+1} per site pair (context-level cache: reused across calls)
  for (int sw = 0; sw < (int)sweeps.size(); ++sw) {
    const SweepParams& sp = sweeps[sw];
    if (sp.noise != 0.0) fail(-2, "noise != 0 is not supported (Chain always runs with noise 0, src/dmrg.h:350)");
    for (int half = 0; half < 2; ++half)
      for (int k = 0; k < n - 1; ++k) {
        const int b = half == 0 ? k + 1 : n - 1 - k;     // 1-based bond (sites b, b+1)
        const int dir = half;                            // 0 Fromleft, 1 Fromright
        const int i = b - 1;
        double t0 = now_s();
        PhiLayout pl; Buf phi;
        phi_of(i, pl, phi);
        Bond bond;
        bond.ctx = c; bond.cplx = cplx; bond.nq = nq; bond.S = &S;
        bond.Dl = A[i].lay.X; bond.Dr = A[i + 1].lay.Y;
        bond.Kl = W[i].KL; bond.Km = W[i].KR; bond.Kr = W[i + 1].KR;
        bond.L = LE[b - 1].buf.p; bond.R = RE[b + 1].buf.p;
        bond.W1 = &W[i]; bond.W2 = &W[i + 1];
        if (!omega[i]) omega[i] = cached_omega(c, W[i], W[i + 1]);
        bond.omega_cache = omega[i].get();
        bond.plan();
        // penalty vectors o_j = FL . Psi_b . Psi_{b+1} . FR^T (built once per bond, LocalMPO_MPS)
        std::vector<Buf> pen(np);
        for (int p = 0; p < np; ++p) {
          DevSite3 X1, X2;
          mat_times_site3(c, cplx, &S, FL[p][b - 1], P[p][i], X1);
          site3_times_matT(c, cplx, &S, P[p][i + 1], FR[p][b + 1], X2);
          pen[p] = Buf(be, (int64_t)esz(cplx) * std::max<int64_t>(pl.total, 1), true);
          two_site(c, cplx, X1, X2, pl, pen[p].p);
          bond.pen.push_back(pen[p].p);
        }
        bond.weight = weight;
        dev_sync(be);
        c->tm.other += now_s() - t0;
        DavidsonResult dr = davidson(bond, phi.p, sp.niter);
        energy = dr.eigenvalue;
        t0 = now_s();
        SplitResult sr = split_bond(c, cplx, &S, pl, phi.p, dir, sp.maxdim, sp.mindim, sp.cutoff);
        c->tm.trunc += now_s() - t0;
        A[i] = std::move(sr.A);
        A[i + 1] = std::move(sr.B);
        t0 = now_s();
        if (dir == 0) {
          env_update_left(c, cplx, &S, LE[b - 1], A[i], W[i], LE[b]);
          for (int p = 0; p < np; ++p) ovl_left(p, b);
        } else {
          env_update_right(c, cplx, &S, RE[b + 1], A[i + 1], W[i + 1], RE[b]);
          for (int p = 0; p < np; ++p) ovl_right(p, b);
        }
        dev_sync(be);
        c->tm.env += now_s() - t0;
        res->stats.push_back({sw, b, dir, (int)sr.mid.total(), energy, sr.truncerr});
        check_dev(c);
      }
  }
  // psi.normalize(): the orthogonality centre is site 1
  {
    const void* xs[1] = {A[0].buf.p};
    double nn[2];
    multi_dot(be, cplx, 1, xs, A[0].buf.p, A[0].lay.total, nn);
    if (nn[0] > 0) {
      double co[2] = {1.0 / std::sqrt(nn[0]), 0.0};
      lincomb(be, cplx, 1, xs, co, 0, 0, A[0].buf.p, A[0].lay.total);
    }
  }
  // export
  res->energy = energy;
  res->psi.n = n;
  res->psi.center = 1;
  res->psi.link.resize(n + 1);
  for (int j = 0; j <= n; ++j) res->psi.link[j] = IndexMap::from_grading(j == 0 ? A[0].lay.X : A[j - 1].lay.Y);
  // the site tensors stay on the device; result_site_export packs one on request (site index in the caller's order, S.map;
  // links in internal, charge-sorted order)
  res->dev = std::move(A);
  res->S = S;
  for (auto& t : res->dev) t.lay.S = &res->S;
  res->ctx = c;
  res->tm = c->tm;
  check_dev(c);
  return res;
}

void result_site_export(const DmrgResult& r, int site, void* out) {
  export_site3(r.ctx, r.cplx, r.S, r.psi.link[site], r.psi.link[site + 1], r.dev[site], out);
  check_dev(r.ctx);
}

// ------------------------------------------------------------------------------------------------ CalcEE
// Entanglement entropy of every cut, as /root/reference/src/utility.cpp:5-27 computes it: move the orthogonality centre along
// the chain, take the Schmidt spectrum p = sigma^2 of each cut, S = -sum_{p > 1e-12} p ln p.  The spectrum comes from the same
// device truncation path as the sweep (rho eigenvalues; cutoff 1e-15 only drops numerically-zero directions).
std::vector<double> mps_entropies(Ctx* c, int nq, int d, const int32_t* site_charge, const HostMPS& psi, bool cplx) {
  const int n = psi.n;
  if (n < 2) fail(-2, "entropies need at least 2 sites");
  SiteInfo S = SiteInfo::make(nq, d, site_charge);
  std::vector<DevSite3> A(n);
  for (int j = 0; j < n; ++j) import_mps_site(c, cplx, cplx, S, psi, j, A[j]);
  auto phi_of = [&](int i, PhiLayout& pl, Buf& phi) {
    pl.build(A[i].lay.X, A[i + 1].lay.Y, &S);
    phi = Buf(c->be, (int64_t)esz(cplx) * std::max<int64_t>(pl.total, 1), true);
    two_site(c, cplx, A[i], A[i + 1], pl, phi.p);
  };
  if (psi.center != 1) {
    for (int i = n - 2; i >= 0; --i) {
      PhiLayout pl; Buf phi;
      phi_of(i, pl, phi);
      SplitResult sr = split_bond(c, cplx, &S, pl, phi.p, 1, INT64_MAX / 4, 1, 1e-20, true);
      A[i] = std::move(sr.A);
      A[i + 1] = std::move(sr.B);
    }
  }
  std::vector<double> out;
  for (int i = 0; i + 1 < n; ++i) {
    PhiLayout pl; Buf phi;
    phi_of(i, pl, phi);
    SplitResult sr = split_bond(c, cplx, &S, pl, phi.p, 0, INT64_MAX / 4, 1, 1e-15, false, true, 0);
    double sv = 0;
    for (double p : sr.spectrum)
      if (p > 1e-12) sv += -p * std::log(p);
    out.push_back(sv);
    A[i] = std::move(sr.A);
    A[i + 1] = std::move(sr.B);
  }
  check_dev(c);
  return out;
}

// ------------------------------------------------------------------------------------------------ translation / overlap
namespace {
void export_mps(Ctx* c, bool cplx, const SiteInfo& S, std::vector<DevSite3>& A, DmrgResult& res) {
  const int n = (int)A.size();
  res.cplx = cplx;
  res.psi.n = n;
  res.psi.link.resize(n + 1);
  for (int j = 0; j <= n; ++j) res.psi.link[j] = IndexMap::from_grading(j == 0 ? A[0].lay.X : A[j - 1].lay.Y);
  res.dev = std::move(A);
  res.S = S;
  for (auto& t : res.dev) t.lay.S = &res.S;
  res.ctx = c;
}
}  // namespace

std::unique_ptr<DmrgResult> mps_translate(Ctx* c, int nq, int d, const int32_t* site_charge, const HostMPS& psi, bool cplx, double cutoff) {
  const int n = psi.n;
  if (n < 2) fail(-2, "translation needs at least 2 sites");
  SiteInfo S = SiteInfo::make(nq, d, site_charge);
  std::vector<DevSite3> A(n);
  for (int j = 0; j < n; ++j) import_mps_site(c, cplx, cplx, S, psi, j, A[j]);
  auto phi_of = [&](int i, PhiLayout& pl, Buf& phi) {
    pl.build(A[i].lay.X, A[i + 1].lay.Y, &S);
    phi = Buf(c->be, (int64_t)esz(cplx) * std::max<int64_t>(pl.total, 1), true);
    two_site(c, cplx, A[i], A[i + 1], pl, phi.p);
  };
  if (psi.center != 1) {   // Swap starts with psi.position(b); bring an uncanonical input to site 1 once, exactly
    for (int i = n - 2; i >= 0; --i) {
      PhiLayout pl; Buf phi;
      phi_of(i, pl, phi);
      SplitResult sr = split_bond(c, cplx, &S, pl, phi.p, 1, INT64_MAX / 4, 1, 1e-20, true);
      A[i] = std::move(sr.A);
      A[i + 1] = std::move(sr.B);
    }
  }
  auto res = std::make_unique<DmrgResult>();
  for (int i = 0; i + 1 < n; ++i) {
    PhiLayout pl; Buf phi;
    phi_of(i, pl, phi);
    // wf[l, s1 = u, s2 = v, r] = phi[l, v, u, r]: a permutation of the fused site-pair index inside every charge block
    Buf sw(c->be, (int64_t)esz(cplx) * std::max<int64_t>(pl.total, 1), true);
    std::vector<CopyTask> tasks;
    for (int ql = 0; ql < nq; ++ql)
      for (int qr = 0; qr < nq; ++qr) {
        const int64_t Dl = pl.L.dim[ql], Dr = pl.R.dim[qr], ns = pl.ns(ql, qr);
        if (Dl == 0 || Dr == 0) continue;
        const int qd = qsub(qr, ql, nq);
        for (int64_t sr = 0; sr < ns; ++sr) {
          const int u = S.srow[qd][sr].first, v = S.srow[qd][sr].second;
          const int64_t src_sr = S.srow_index[(size_t)v * d + u];
          CopyTask t{};
          t.src_off = pl.o(ql, qr) + Dl * src_sr; t.dst_off = pl.o(ql, qr) + Dl * sr;
          t.n0 = (int32_t)Dl; t.n1 = (int32_t)Dr; t.n2 = 1;
          t.ss0 = 1; t.ds0 = 1; t.ss1 = Dl * ns; t.ds1 = Dl * ns; t.ss2 = 0; t.ds2 = 0;
          tasks.push_back(t);
        }
      }
    run_copy_tasks(c, cplx, tasks, phi.p, sw.p);
    // svd(wf, inds(psi(b)), {"Cutoff", cutoff}): U -> site b, S V -> site b+1 (not normalised)
    SplitResult sr = split_bond(c, cplx, &S, pl, sw.p, 0, INT64_MAX / 4, 1, cutoff, false, false);
    res->stats.push_back({0, i + 1, 0, (int)sr.mid.total(), 0.0, sr.truncerr});
    A[i] = std::move(sr.A);
    A[i + 1] = std::move(sr.B);
  }
  export_mps(c, cplx, S, A, *res);
  res->psi.center = n;
  check_dev(c);
  return res;
}

cplx_t mps_overlap(Ctx* c, int nq, int d, const int32_t* site_charge, const HostMPS& a, bool a_cplx, const HostMPS& b, bool b_cplx) {
  const int n = a.n;
  if (b.n != n || n < 1) fail(-2, "overlap: length mismatch");
  const bool cplx = a_cplx || b_cplx;
  SiteInfo S = SiteInfo::make(nq, d, site_charge);
  DevMat F;
  {
    DevSite3 A0, B0;
    import_mps_site(c, cplx, a_cplx, S, a, 0, A0);
    import_mps_site(c, cplx, b_cplx, S, b, 0, B0);
    if (A0.lay.X.total() != 1 || B0.lay.X.total() != 1) fail(-2, "overlap: outer links must have dimension 1");
    if (!(A0.lay.X == B0.lay.X)) return cplx_t(0, 0);          // different total charge
    site3H_times_site3(c, cplx, &S, A0, B0, F);                 // F[m, n] = sum_{l,s} conj(A[l,s,m]) B[l,s,n]
  }
  for (int j = 1; j < n; ++j) {
    DevSite3 Aj, Bj, X;
    import_mps_site(c, cplx, a_cplx, S, a, j, Aj);
    import_mps_site(c, cplx, b_cplx, S, b, j, Bj);
    // X[m, s, n'] = sum_n F[m, n] B[n, s, n'] ; F'[m', n'] = sum_{m,s} conj(A[m,s,m']) X[m,s,n']
    mat_times_site3(c, cplx, &S, F, Bj, X);
    DevMat Fn;
    site3H_times_site3(c, cplx, &S, Aj, X, Fn);
    F = std::move(Fn);
  }
  if (F.total == 0) return cplx_t(0, 0);                        // right boundary links carry different charges
  if (F.total != 1) fail(-2, "overlap: right boundary is not 1x1");
  std::vector<char> h(esz(cplx), 0);
  dev_d2h(c->be, h.data(), F.buf.p, h.size());
  check_dev(c);
  return rd(h.data(), cplx, 0);
}

// ------------------------------------------------------------------------------------------------ <bra| MPO |ket>
cplx_t mps_mpo_element(Ctx* c, int d, const HostMPS& bra, bool bra_cplx, const MpoIn& O, const HostMPS& ket, bool ket_cplx) {
  const int n = O.n;
  if (bra.n != n || ket.n != n || n < 1) fail(-2, "mpo_element: length mismatch");
  if (O.nq != 1) fail(-2, "mpo_element: dense (nq = 1) states only -- the reference removes the QNs before measuring rho (src/dmrg.h:781-784)");
  const bool cplx = bra_cplx || ket_cplx || O.cplx;
  Backend* be = c->be;
  const size_t es = esz(cplx);
  SiteInfo S = SiteInfo::make(1, d, nullptr);
  // environment E[x' (bra), a, x (ket)], element x' + Db*(a + k*x); starts as the 1x1x1 identity
  int64_t Db = 1, Dk = 1, k = 1;
  Buf E(be, (int64_t)es);
  {
    std::vector<char> one(es, 0);
    wr(one.data(), cplx, 0, cplx_t(1, 0));
    dev_h2d(be, E.p, one.data(), one.size());
  }
  for (int j = 0; j < n; ++j) {
    DevSite3 Ak, Ab;
    import_mps_site(c, cplx, ket_cplx, S, ket, j, Ak);
    import_mps_site(c, cplx, bra_cplx, S, bra, j, Ab);
    HostW W = import_w(1, d, S, O.link[j], O.link[j + 1], O.cplx, O.site[j]);
    const int64_t kl = W.KL.total(), kr = W.KR.total();
    const int64_t Dkr = Ak.lay.Y.total(), Dbr = Ab.lay.Y.total();
    if (Ak.lay.X.total() != Dk || Ab.lay.X.total() != Db || kl != k) fail(-2, "mpo_element: inconsistent bond dimensions at site " + std::to_string(j));
    // stage 1: X[(x',a), (s,m)] = sum_x E[(x',a), x] Aket[x, (s,m)]
    const int64_t M1 = Db * kl, N1 = (int64_t)d * Dkr;
    Buf X(be, (int64_t)es * std::max<int64_t>(M1 * N1, 1));
    {
      GB gb(cplx);
      gb.begin(0, M1, M1, N1, 0);
      gb.seg(0, M1, 0, Dk, Dk);
      gb.end();
      gb.run(c, false, false, E.p, Ak.buf.p, X.p);
    }
    // stage 2 (sparse W): Y[(x',t), (a',m)] = sum_{a,s} W[a,t,s,a'] X[(x',a), (s,m)]
    const int64_t MY = Db * d;
    Buf Y(be, (int64_t)es * std::max<int64_t>(MY * kr * Dkr, 1), true);
    {
      std::vector<MixPanel> panels(1);
      std::vector<MixIn> ins;
      std::vector<MixOut> outs;
      std::vector<MixEntry> ents;
      std::vector<int> slot((size_t)kl * d, -1);
      MixPanel& p = panels[0];
      p = MixPanel{};
      p.rows = (int32_t)Db; p.nr = (int32_t)Dkr;
      for (int t = 0; t < d; ++t)
        for (int64_t ap = 0; ap < kr; ++ap) {
          MixOut o{};
          o.off = Db * t + MY * ap; o.rstride = MY * kr; o.e_begin = (int32_t)ents.size();
          for (int s2 = 0; s2 < d; ++s2)
            for (int64_t a = 0; a < kl; ++a) {
              const cplx_t v = W.at(a, t, s2, ap);
              if (v == cplx_t(0, 0)) continue;
              int& sl = slot[(size_t)(a + kl * s2)];
              if (sl < 0) {
                sl = (int)ins.size();
                MixIn mi{};
                mi.off = Db * a + M1 * s2; mi.rstride = M1 * d;      // X column (s, m) = s + d*m
                ins.push_back(mi);
              }
              MixEntry e{};
              e.slot = sl; e.cre = v.real(); e.cim = v.imag();
              ents.push_back(e);
            }
          o.e_end = (int32_t)ents.size();
          outs.push_back(o);
        }
      p.out_begin = 0; p.out_end = (int32_t)outs.size();
      p.in_begin = 0; p.in_count = (int32_t)ins.size();
      if (Db > 0 && Dkr > 0 && !ins.empty()) {
        const int rpb = mix_rows_per_block(cplx, p.in_count);
        p.blk_begin = 0; p.row_chunks = (int32_t)((Db + rpb - 1) / rpb);
        const int blocks = p.row_chunks * (int)Dkr;
        DevList<MixPanel> dp; DevList<MixIn> di; DevList<MixOut> dox; DevList<MixEntry> de;
        dp.upload(be, panels); di.upload(be, ins); dox.upload(be, outs); de.upload(be, ents);
        launch_mix(be, cplx, X.p, Y.p, dp.ptr(), 1, di.ptr(), dox.ptr(), de.ptr(), blocks, p.in_count);
      }
    }
    // stage 3: E'[m', (a',m)] = sum_{(x',t)} conj(Abra[(x',t), m']) Y[(x',t), (a',m)]
    Buf En(be, (int64_t)es * std::max<int64_t>(Dbr * kr * Dkr, 1), true);
    {
      GB gb(cplx);
      gb.begin(0, Dbr, Dbr, kr * Dkr, GEMM_TRANS_A | GEMM_CONJ_A);
      gb.seg(0, MY, 0, MY, MY);
      gb.end();
      gb.run(c, true, false, Ab.buf.p, Y.p, En.p);
    }
    E = std::move(En);
    Db = Dbr; Dk = Dkr; k = kr;
  }
  if (Db != 1 || Dk != 1 || k != 1) fail(-2, "mpo_element: outer bonds must have dimension 1");
  std::vector<char> h(es, 0);
  dev_d2h(be, h.data(), E.p, h.size());
  check_dev(c);
  return rd(h.data(), cplx, 0);
}

}  // namespace cb200
