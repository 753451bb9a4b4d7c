"""Host-side mirror of the reference's interface for the hot path.

The reference calls `std::tie(en, psi) = dmrg(H, psis, psi0, sweeps, {"Quiet",true,"Weight=",1000.0})`
(/root/reference/src/dmrg.h:544,563,596) with ITensor's MPO / MPS / Sweeps types.  `dmrg()` below has the same
argument meaning and return value and forwards to the C ABI (cb200_dmrg); MPS, MPO and Sweeps are thin
containers of the dense tensors + charge labels that cross that boundary.  No arithmetic happens in Python.

Array convention on the Python side: NumPy arrays indexed A[l, s, r] / W[a, s_out, s_in, a'] (any memory order);
they are converted to column-major storage right before the call.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ChainB200Error, Index, Tensor, Sweep, MpsC, MpoC, BondStat, ptr, i32ptr


class Context:
    """One device context (cb200_ctx). Fails loudly when the CUDA library / device is unavailable."""

    def __init__(self, device=0, lib_path=None):
        self.lib = _lib.load(lib_path)
        h = C.c_void_p()
        rc = self.lib.cb200_create(C.byref(h), device)
        if rc != 0:
            raise ChainB200Error(rc, (self.lib.cb200_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.cb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise ChainB200Error(rc, (self.lib.cb200_last_error(self.h) or b"").decode())

    @property
    def backend(self):
        return self.lib.cb200_backend_name().decode()

    SHARD_MODES = {"none": 0, "nccl": 1, "p2p": 2}

    def shard(self, rank, world, mode="p2p", landing_bytes=0, min_dim=1600, group=None):
        """Row-shard every H_eff matvec of this context over `world` ranks (one process + context per GPU).

        mode "none": no exchange (the caller sums the ranks' partial vectors); "nccl": ncclAllReduce of the zero-padded
        partials inside the library; "p2p": all-gather fused into the R-stage GEMM epilogue over CUDA-IPC peer mappings.
        The rendezvous (NCCL unique id / IPC handles) travels through torch.distributed's `group` (any backend), which is
        plumbing only: the data path never touches torch."""
        m = self.SHARD_MODES[mode]
        uid, handles = None, None
        if world > 1 and m != 0:
            import torch.distributed as dist
            if m == 1:
                buf = C.create_string_buffer(128)
                if rank == 0:
                    self.check(self.lib.cb200_nccl_unique_id(self.h, buf))
                box = [buf.raw if rank == 0 else None]
                dist.broadcast_object_list(box, src=0, group=group)
                uid = C.create_string_buffer(box[0], 128)
            else:
                if landing_bytes <= 0:
                    raise ValueError("p2p sharding needs landing_bytes >= the largest phi in bytes")
                # every rank takes part in both exchanges even when its own step failed, so that a CUDA-IPC problem on one rank
                # surfaces as the same exception everywhere instead of a hang
                buf = C.create_string_buffer(64)
                rc = self.lib.cb200_shard_alloc(self.h, int(landing_bytes), buf)
                allh = [None] * world
                dist.all_gather_object(allh, buf.raw if rc == 0 else None, group=group)
                if any(h is None for h in allh):
                    raise ChainB200Error(-3, "p2p sharding unavailable: cb200_shard_alloc failed on rank(s) %s" % [i for i, h in enumerate(allh) if h is None])
                handles = C.create_string_buffer(b"".join(allh), 64 * world)
                rc = self.lib.cb200_shard_setup(self.h, rank, world, m, uid, handles, int(min_dim))
                rcs = [None] * world
                dist.all_gather_object(rcs, int(rc), group=group)
                if any(r != 0 for r in rcs):
                    self.lib.cb200_shard_setup(self.h, 0, 1, 0, None, None, int(min_dim))      # back to unsharded
                    raise ChainB200Error(-3, "p2p sharding unavailable: peer mapping failed on rank(s) %s" % [i for i, r in enumerate(rcs) if r != 0])
                self.shard_rank, self.shard_world, self.shard_mode = rank, world, mode
                return
        self.check(self.lib.cb200_shard_setup(self.h, rank, world, m, uid, handles, int(min_dim)))
        self.shard_rank, self.shard_world, self.shard_mode = rank, world, mode

    @property
    def launches(self):
        return int(self.lib.cb200_launch_count(self.h))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class Sweeps:
    """itensor::Sweeps: per-sweep maxdim / mindim / cutoff / niter / noise; short lists repeat their last entry."""

    def __init__(self, nsweep, maxdim, cutoff, niter=2, mindim=1, noise=0.0):
        def ext(x):
            x = list(x) if isinstance(x, (list, tuple)) else [x]
            return [x[min(i, len(x) - 1)] for i in range(nsweep)]
        self.nsweep = nsweep
        self.maxdim, self.cutoff, self.niter, self.mindim, self.noise = ext(maxdim), ext(cutoff), ext(niter), ext(mindim), ext(noise)

    def to_c(self):
        arr = (Sweep * self.nsweep)()
        for i in range(self.nsweep):
            arr[i].maxdim = int(min(self.maxdim[i], 2 ** 31 - 1))
            arr[i].mindim = int(self.mindim[i])
            arr[i].niter = int(self.niter[i])
            arr[i].cutoff = float(self.cutoff[i])
            arr[i].noise = float(self.noise[i])
        return arr


class MPS:
    def __init__(self, sites, link_charges=None, nq=1, center=0):
        self.sites = [np.asarray(a) for a in sites]
        n = len(self.sites)
        if link_charges is None:
            link_charges = [np.zeros(self.sites[0].shape[0], np.int32)] + [np.zeros(a.shape[2], np.int32) for a in self.sites]
        self.link_charges = [np.ascontiguousarray(q, dtype=np.int32) for q in link_charges]
        assert len(self.link_charges) == n + 1
        self.nq = nq
        self.center = center

    @property
    def n(self):
        return len(self.sites)

    def bond_dims(self):
        return [a.shape[2] for a in self.sites[:-1]]

    def is_complex(self):
        return any(np.iscomplexobj(a) for a in self.sites)

    def _to_c(self, keep):
        dt = np.complex128 if self.is_complex() else np.float64
        n = self.n
        ts = (Tensor * n)()
        ls = (Index * (n + 1))()
        for j, a in enumerate(self.sites):
            f = _lib.fcol(a, dt)
            keep.append(f)
            ts[j].dtype = 1 if dt is np.complex128 else 0
            ts[j].rank = 3
            for k in range(3):
                ts[j].dims[k] = f.shape[k]
            ts[j].data = f.ctypes.data
        for j, q in enumerate(self.link_charges):
            ls[j].dim = len(q)
            ls[j].charge = i32ptr(q)
        keep += [ts, ls]
        m = MpsC()
        m.n, m.nq, m.site, m.link, m.ortho_center = n, self.nq, ts, ls, int(self.center or 0)
        return m


def _site_block_slices(ql, sq, qr, nq):
    rl, rs, rr = _charge_runs(ql), _charge_runs(sq), _charge_runs(qr)
    for (r0, r1, cr) in rr:
        for (a0, a1, cs) in rs:
            for (l0, l1, cl) in rl:
                if (cl + cs - cr) % nq == 0:
                    yield (slice(l0, l1), slice(a0, a1), slice(r0, r1))


def pack_site_blocks(A, ql, sq, qr, nq):
    """Dense [l, s, r] -> the block-packed ("QDense order") layout of cb200_mps.layout = 1 (host-side helper)."""
    parts = [np.asarray(A[sl]).ravel(order="F") for sl in _site_block_slices(ql, sq, qr, nq)]
    return np.concatenate(parts) if parts else np.zeros(0, np.asarray(A).dtype)


def unpack_site_blocks(xb, ql, sq, qr, nq):
    out = np.zeros((len(ql), len(sq), len(qr)), dtype=np.asarray(xb).dtype, order="F")
    o = 0
    for sl in _site_block_slices(ql, sq, qr, nq):
        shp = tuple(x.stop - x.start for x in sl)
        k = int(np.prod(shp))
        out[sl] = np.asarray(xb[o:o + k]).reshape(shp, order="F")
        o += k
    return out


class PackedMPS:
    """An MPS whose site tensors cross the boundary block-packed (cb200_mps.layout = 1): what an adapter hands over when it passes
    ITensor's QDense storage as it is.  blocks[j] is the 1-D packed array of site j."""

    def __init__(self, blocks, link_charges, site_charges, nq=1, center=0):
        self.blocks = [np.ascontiguousarray(b) for b in blocks]
        self.link_charges = [np.ascontiguousarray(q, dtype=np.int32) for q in link_charges]
        self.site_charges = np.ascontiguousarray(site_charges, dtype=np.int32)
        assert len(self.link_charges) == len(self.blocks) + 1
        self.nq, self.center = nq, center

    @staticmethod
    def from_dense(psi, site_charges):
        sq = np.ascontiguousarray(site_charges, dtype=np.int32)
        return PackedMPS([pack_site_blocks(a, psi.link_charges[j], sq, psi.link_charges[j + 1], psi.nq) for j, a in enumerate(psi.sites)],
                         psi.link_charges, sq, nq=psi.nq, center=psi.center)

    def to_dense(self):
        return MPS([unpack_site_blocks(b, self.link_charges[j], self.site_charges, self.link_charges[j + 1], self.nq)
                    for j, b in enumerate(self.blocks)], self.link_charges, nq=self.nq, center=self.center)

    @property
    def n(self):
        return len(self.blocks)

    @property
    def d(self):
        return len(self.site_charges)

    def bond_dims(self):
        return [len(q) for q in self.link_charges[1:-1]]

    def is_complex(self):
        return any(np.iscomplexobj(b) for b in self.blocks)

    def _to_c(self, keep):
        dt = np.complex128 if self.is_complex() else np.float64
        n = self.n
        ts = (Tensor * n)()
        ls = (Index * (n + 1))()
        for j, b in enumerate(self.blocks):
            f = np.ascontiguousarray(b, dtype=dt)
            keep.append(f)
            ts[j].dtype = 1 if dt is np.complex128 else 0
            ts[j].rank = 3
            ts[j].dims[0], ts[j].dims[1], ts[j].dims[2] = len(self.link_charges[j]), self.d, len(self.link_charges[j + 1])
            ts[j].data = f.ctypes.data
        for j, q in enumerate(self.link_charges):
            ls[j].dim = len(q)
            ls[j].charge = i32ptr(q)
        keep += [ts, ls]
        m = MpsC()
        m.n, m.nq, m.site, m.link, m.ortho_center, m.layout = n, self.nq, ts, ls, int(self.center or 0), 1
        return m


class MPO:
    def __init__(self, sites, link_charges=None, site_charges=None, nq=1):
        self.sites = [np.asarray(w) for w in sites]
        n = len(self.sites)
        d = self.sites[0].shape[1]
        if link_charges is None:
            link_charges = [np.zeros(self.sites[0].shape[0], np.int32)] + [np.zeros(w.shape[3], np.int32) for w in self.sites]
        self.link_charges = [np.ascontiguousarray(q, dtype=np.int32) for q in link_charges]
        self.site_charges = np.ascontiguousarray(site_charges if site_charges is not None else np.zeros(d), dtype=np.int32)
        self.nq = nq
        self.d = d
        assert len(self.link_charges) == n + 1

    @property
    def n(self):
        return len(self.sites)

    def is_complex(self):
        return any(np.iscomplexobj(w) and np.abs(np.imag(w)).max() > 0 for w in self.sites)

    def _to_c(self, keep):
        dt = np.complex128 if self.is_complex() else np.float64
        n = self.n
        ts = (Tensor * n)()
        ls = (Index * (n + 1))()
        for j, w in enumerate(self.sites):
            f = _lib.fcol(np.real(w) if dt is np.float64 else w, dt)
            keep.append(f)
            ts[j].dtype = 1 if dt is np.complex128 else 0
            ts[j].rank = 4
            for k in range(4):
                ts[j].dims[k] = f.shape[k]
            ts[j].data = f.ctypes.data
        for j, q in enumerate(self.link_charges):
            ls[j].dim = len(q)
            ls[j].charge = i32ptr(q)
        keep += [ts, ls]
        m = MpoC()
        m.n, m.nq, m.site, m.link, m.site_charge, m.d = n, self.nq, ts, ls, i32ptr(self.site_charges), self.d
        return m


class DmrgInfo:
    """What ITensor prints / returns besides (energy, psi): per-bond stats and phase timers."""

    def __init__(self):
        self.stats = []
        self.timers = {}


def _unpack_result_blocks(ctx, res):
    """The result's site tensors in the block-packed layout (cb200_result_site_blocks): (blocks, link charges)."""
    lib = ctx.lib
    n = lib.cb200_result_nsites(res)
    dt = np.complex128 if lib.cb200_result_dtype(res) == 1 else np.float64
    charges = []
    for j in range(n + 1):
        q = np.zeros(int(lib.cb200_result_link_dim(res, j)), np.int32)
        ctx.check(lib.cb200_result_link_charges(res, j, i32ptr(q)))
        charges.append(q)
    blocks = []
    for j in range(n):
        b = np.empty(max(int(lib.cb200_result_site_blocks_elems(res, j)), 0), dtype=dt)
        if b.size:
            ctx.check(lib.cb200_result_site_blocks(res, j, ptr(b)))
        blocks.append(b)
    return blocks, charges


def _unpack_result(ctx, res, d):
    lib = ctx.lib
    n = lib.cb200_result_nsites(res)
    dt = np.complex128 if lib.cb200_result_dtype(res) == 1 else np.float64
    dims = [int(lib.cb200_result_link_dim(res, j)) for j in range(n + 1)]
    charges = []
    for j in range(n + 1):
        q = np.zeros(dims[j], np.int32)
        ctx.check(lib.cb200_result_link_charges(res, j, i32ptr(q)))
        charges.append(q)
    sites = []
    for j in range(n):
        a = np.empty((dims[j], d, dims[j + 1]), dtype=dt, order="F")      # the library writes every element (zeros included)
        ctx.check(lib.cb200_result_site(res, j, ptr(a)))
        sites.append(a)
    return sites, charges


def translate(psi, site_charges=None, cutoff=1e-5, ctx=None):
    """One-site translation by L-1 swap gates + truncated SVDs, as Measure("translation") applies it (src/dmrg.h:855-858)."""
    ctx = ctx or default_context()
    keep = []
    pc = psi._to_c(keep)
    d = psi.sites[0].shape[1]
    sq = np.ascontiguousarray(site_charges if site_charges is not None else np.zeros(d), dtype=np.int32)
    res = C.c_void_p()
    ctx.check(ctx.lib.cb200_mps_translate(ctx.h, C.byref(pc), d, i32ptr(sq), float(cutoff), C.byref(res)))
    try:
        sites, charges = _unpack_result(ctx, res, d)
    finally:
        ctx.lib.cb200_result_free(res)
    return MPS(sites, charges, nq=psi.nq, center=psi.n)


def overlap(a, b, site_charges=None, ctx=None):
    """innerC(a, b) = <a|b>."""
    ctx = ctx or default_context()
    keep = []
    ac, bc = a._to_c(keep), b._to_c(keep)
    d = a.d if isinstance(a, PackedMPS) else a.sites[0].shape[1]
    if site_charges is None and isinstance(a, PackedMPS):
        site_charges = a.site_charges
    sq = np.ascontiguousarray(site_charges if site_charges is not None else np.zeros(d), dtype=np.int32)
    z = (C.c_double * 2)()
    ctx.check(ctx.lib.cb200_mps_overlap(ctx.h, C.byref(ac), C.byref(bc), d, i32ptr(sq), z))
    return complex(z[0], z[1])


def mpo_element(bra, O, ket, ctx=None):
    """<bra| O |ket> for an MPO and two (different) dense MPS (nq = 1)."""
    ctx = ctx or default_context()
    keep = []
    bc, oc, kc = bra._to_c(keep), O._to_c(keep), ket._to_c(keep)
    z = (C.c_double * 2)()
    ctx.check(ctx.lib.cb200_mps_mpo_element(ctx.h, C.byref(bc), C.byref(oc), C.byref(kc), z))
    return complex(z[0], z[1])


def dmrg(H, psis, psi0, sweeps, weight=1.0, ctx=None, info=None):
    """(energy, psi) = dmrg(H, psis, psi0, sweeps, {"Weight": weight}) -- same meaning as itensor::dmrg at
    src/dmrg.h:544: `psis` are the previously found states to be penalised with `weight`."""
    ctx = ctx or default_context()
    lib = ctx.lib
    keep = []
    Hc = H._to_c(keep)
    p0 = psi0._to_c(keep)
    prev = (MpsC * max(len(psis), 1))()
    for i, p in enumerate(psis):
        prev[i] = p._to_c(keep)
    sw = sweeps.to_c()
    energy = C.c_double()
    res = C.c_void_p()
    ctx.check(lib.cb200_dmrg(ctx.h, C.byref(Hc), prev, len(psis), C.byref(p0), sw, sweeps.nsweep, float(weight),
                             C.byref(energy), C.byref(res)))
    try:
        if isinstance(psi0, PackedMPS):      # the answer comes back in the layout the start state arrived in
            blocks, charges = _unpack_result_blocks(ctx, res)
            out = PackedMPS(blocks, charges, H.site_charges, nq=H.nq, center=1)
        else:
            sites, charges = _unpack_result(ctx, res, H.d)
            out = MPS(sites, charges, nq=H.nq, center=1)
        if info is not None:
            ns = lib.cb200_result_nstats(res)
            arr = (BondStat * max(ns, 1))()
            ctx.check(lib.cb200_result_stats(res, arr))
            info.stats = [dict(sweep=arr[i].sweep, bond=arr[i].bond, dir=arr[i].dir, m=arr[i].m, energy=arr[i].energy,
                               truncerr=arr[i].truncerr) for i in range(ns)]
            t = (C.c_double * 7)()
            ctx.check(lib.cb200_result_timers(res, t))
            info.timers = dict(env=t[0], matvec=t[1], level1=t[2], trunc=t[3], other=t[4], nmatvec=t[5], matvec_flops=t[6])
    finally:
        lib.cb200_result_free(res)
    return energy.value, out


def calc_ee(psi, site_charges=None, ctx=None):
    """CalcEE of /root/reference/src/utility.cpp:5-27 on the device: entropy of every cut (p > 1e-12 rule, natural log)."""
    ctx = ctx or default_context()
    keep = []
    pc = psi._to_c(keep)
    d = psi.d if isinstance(psi, PackedMPS) else psi.sites[0].shape[1]
    if site_charges is None and isinstance(psi, PackedMPS):
        site_charges = psi.site_charges
    sq = np.ascontiguousarray(site_charges if site_charges is not None else np.zeros(d), dtype=np.int32)
    out = np.zeros(psi.n - 1)
    ctx.check(ctx.lib.cb200_mps_entropies(ctx.h, C.byref(pc), d, i32ptr(sq), out.ctypes.data_as(C.POINTER(C.c_double))))
    return out.tolist()


class Bond:
    """One two-site problem resident on the device (cb200_bond): LocalMPO::product, davidson, svdBond, makeL/makeR."""

    def __init__(self, L, W1, W2, R, ql=None, qr=None, qkl=None, qkm=None, qkr=None, site_charges=None, nq=1, ctx=None):
        self.ctx = ctx or default_context()
        lib = self.ctx.lib
        cplx = any(np.iscomplexobj(x) and np.abs(np.imag(x)).max() > 0 for x in (L, W1, W2, R))
        self.dt = np.complex128 if cplx else np.float64
        self.cplx = cplx
        self.Dl, self.kl = L.shape[0], L.shape[1]
        self.Dr, self.kr = R.shape[0], R.shape[1]
        self.km, self.d = W1.shape[3], W1.shape[1]
        self.nq = nq
        z = lambda n: np.zeros(n, np.int32)
        self.q = [np.ascontiguousarray(x if x is not None else z(n), dtype=np.int32)
                  for x, n in ((ql, self.Dl), (qr, self.Dr), (qkl, self.kl), (qkm, self.km), (qkr, self.kr))]
        self.sq = np.ascontiguousarray(site_charges if site_charges is not None else z(self.d), dtype=np.int32)
        idx = [Index(len(q), i32ptr(q)) for q in self.q]
        conv = lambda a: _lib.fcol(a if cplx else np.real(a), self.dt)
        arrs = [conv(L), conv(W1), conv(W2), conv(R)]
        h = C.c_void_p()
        self.ctx.check(lib.cb200_bond_create(self.ctx.h, 1 if cplx else 0, nq, self.d, i32ptr(self.sq), C.byref(idx[0]), C.byref(idx[1]),
                                             C.byref(idx[2]), C.byref(idx[3]), C.byref(idx[4]), ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]),
                                             ptr(arrs[3]), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.cb200_bond_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def matvec_flops(self):
        return float(self.ctx.lib.cb200_bond_matvec_flops(self.h))

    def _phi(self, phi):
        return _lib.fcol(phi, self.dt)

    def set_penalty(self, vecs, weight):
        arrs = [self._phi(o) for o in vecs]
        ps = (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
        self.ctx.check(self.ctx.lib.cb200_bond_set_penalty(self.h, len(arrs), ps, float(weight)))

    def apply(self, phi, out=None):
        x = self._phi(phi)
        y = out if out is not None else np.zeros_like(x, order="F")
        self.ctx.check(self.ctx.lib.cb200_bond_apply(self.h, ptr(x), ptr(y)))
        return y

    # ---- ITensor's QDense storage at the boundary: only the charge-allowed blocks cross PCIe (include/chain_b200.h)
    @property
    def phi_blocks_elems(self):
        return int(self.ctx.lib.cb200_bond_phi_blocks_elems(self.h))

    def phi_blocks(self):
        """[(offset, (len_l, len_s1, len_s2, len_r))] of the block-packed layout, in storage order."""
        n = C.c_int32()
        self.ctx.check(self.ctx.lib.cb200_bond_phi_blocks(self.h, C.byref(n), None, None))
        off = (C.c_int64 * max(n.value, 1))()
        dims = (C.c_int64 * max(4 * n.value, 1))()
        self.ctx.check(self.ctx.lib.cb200_bond_phi_blocks(self.h, C.byref(n), off, dims))
        return [(off[i], tuple(dims[4 * i:4 * i + 4])) for i in range(n.value)]

    def pack(self, phi):
        return pack_phi_blocks(self._phi(phi), self.q[0], self.sq, self.q[1], self.nq)

    def unpack(self, xb):
        return unpack_phi_blocks(xb, self.q[0], self.sq, self.q[1], self.nq)

    def apply_blocks(self, xb, out=None):
        """LocalMPO::product with block-packed host vectors (1-D arrays of phi_blocks_elems elements)."""
        x = np.ascontiguousarray(xb, dtype=self.dt)
        if x.size != self.phi_blocks_elems:
            raise ValueError("block-packed phi has %d elements, expected %d" % (x.size, self.phi_blocks_elems))
        y = out if out is not None else np.zeros(x.size, self.dt)
        self.ctx.check(self.ctx.lib.cb200_bond_apply_blocks(self.h, ptr(x), ptr(y)))
        return y

    def stage_bytes(self):
        """Algorithmic HBM bytes of [L-stage GEMM, W(x)W mix, R-stage GEMM] of one product."""
        by = (C.c_double * 3)()
        self.ctx.check(self.ctx.lib.cb200_bond_stage_bytes(self.h, by))
        return list(by)

    def phi_blocks_slice(self):
        """(begin, count): the contiguous slice of the block-packed vector this rank carries in apply_blocks_sharded."""
        b, n = C.c_int64(), C.c_int64()
        self.ctx.check(self.ctx.lib.cb200_bond_phi_blocks_slice(self.h, C.byref(b), C.byref(n)))
        return b.value, n.value

    def apply_blocks_sharded(self, x_slice, out=None):
        """The product on a sharded context with only this rank's slice of the block-packed vectors crossing the host boundary."""
        b, n = self.phi_blocks_slice()
        x = np.ascontiguousarray(x_slice, dtype=self.dt)
        if x.size != n:
            raise ValueError("slice has %d elements, expected %d" % (x.size, n))
        y = out if out is not None else np.zeros(max(n, 1), self.dt)[:n]
        self.ctx.check(self.ctx.lib.cb200_bond_apply_blocks_sharded(self.h, ptr(x), ptr(y)))
        return y

    def apply_timed(self, phi, reps=5, warmup=3):
        x = self._phi(phi)
        ms = C.c_double()
        self.ctx.check(self.ctx.lib.cb200_bond_apply_timed(self.h, ptr(x), reps, warmup, C.byref(ms)))
        return ms.value

    def stage_times(self, phi, reps=3):
        """(ms[3], flops[3]) for [L-stage GEMM, W(x)W mix, R-stage GEMM], CUDA events on the library's stream."""
        x = self._phi(phi)
        ms = (C.c_double * 3)()
        fl = (C.c_double * 3)()
        self.ctx.check(self.ctx.lib.cb200_bond_stage_times(self.h, ptr(x), reps, ms, fl))
        return list(ms), list(fl)

    def update_timed(self, phi, direction="left", maxdim=2 ** 30, cutoff=1e-8, niter=2):
        """One full bond update on the device; returns dict(energy, m, jacobi_sweeps, matvec, level1, trunc, env) [seconds]."""
        x = self._phi(phi)
        en, m, js = C.c_double(), C.c_int32(), C.c_int32()
        t = (C.c_double * 4)()
        self.ctx.check(self.ctx.lib.cb200_bond_update_timed(self.h, ptr(x), 0 if direction == "left" else 1, int(min(maxdim, 2 ** 31 - 1)),
                                                            float(cutoff), niter, C.byref(en), C.byref(m), C.byref(js), t))
        return dict(energy=en.value, m=m.value, jacobi_sweeps=js.value, matvec=t[0], level1=t[1], trunc=t[2], env=t[3])

    def davidson(self, phi, niter=2):
        x = self._phi(phi).copy(order="F")
        ev, nmv = C.c_double(), C.c_int32()
        self.ctx.check(self.ctx.lib.cb200_bond_davidson(self.h, ptr(x), niter, C.byref(ev), C.byref(nmv)))
        return ev.value, x, nmv.value

    def truncate(self, phi, direction, maxdim, cutoff, mindim=1):
        """direction 'left' (Fromleft) or 'right'. Returns A, B, charges_mid, spectrum, truncerr."""
        x = self._phi(phi)
        d = 0 if direction == "left" else 1
        cap = min(self.Dl * self.d, self.d * self.Dr)
        m = C.c_int32()
        qm = np.zeros(cap, np.int32)
        terr = C.c_double()
        spec = np.zeros(cap)
        lib = self.ctx.lib
        self.ctx.check(lib.cb200_bond_truncate(self.h, ptr(x), d, int(min(maxdim, 2 ** 31 - 1)), mindim, float(cutoff), C.byref(m), i32ptr(qm),
                                               C.byref(terr), spec.ctypes.data_as(C.POINTER(C.c_double)), None, None))
        mm = m.value
        A = np.zeros((self.Dl, self.d, mm), dtype=self.dt, order="F")
        B = np.zeros((mm, self.d, self.Dr), dtype=self.dt, order="F")
        self.ctx.check(lib.cb200_bond_truncate(self.h, ptr(x), d, int(min(maxdim, 2 ** 31 - 1)), mindim, float(cutoff), C.byref(m), i32ptr(qm),
                                               C.byref(terr), spec.ctypes.data_as(C.POINTER(C.c_double)), ptr(A), ptr(B)))
        return A, B, qm[:mm].copy(), spec[:mm].copy(), terr.value

    def env_update(self, direction, T, charges_mid=None):
        """makeL ('left', T = A[Dl,d,m]) or makeR ('right', T = B[m,d,Dr]); returns the new environment [m, km, m]."""
        d = 0 if direction == "left" else 1
        m = T.shape[2] if d == 0 else T.shape[0]
        qm = np.ascontiguousarray(charges_mid if charges_mid is not None else np.zeros(m), dtype=np.int32)
        t = _lib.fcol(T, self.dt)
        out = np.zeros((m, self.km, m), dtype=self.dt, order="F")
        self.ctx.check(self.ctx.lib.cb200_bond_env_update(self.h, d, m, i32ptr(qm), ptr(t), ptr(out)))
        return out


# ---- kernel-level helpers (unit tests / calibration) -------------------------------------------------------------
def _charge_runs(q):
    """Sectors of an index = maximal runs of equal charge label: [(start, stop, charge)]."""
    q = np.asarray(q)
    if q.size == 0:
        return []
    cut = np.flatnonzero(np.diff(q)) + 1
    starts = np.concatenate([[0], cut])
    stops = np.concatenate([cut, [q.size]])
    return [(int(a), int(b), int(q[a])) for a, b in zip(starts, stops)]


def _phi_block_slices(ql, sq, qr, nq):
    rl, rs, rr = _charge_runs(ql), _charge_runs(sq), _charge_runs(qr)
    for (r0, r1, cr) in rr:
        for (b0, b1, c2) in rs:
            for (a0, a1, c1) in rs:
                for (l0, l1, cl) in rl:
                    if (cl + c1 + c2 - cr) % nq == 0:
                        yield (slice(l0, l1), slice(a0, a1), slice(b0, b1), slice(r0, r1))


def pack_phi_blocks(phi, ql, sq, qr, nq):
    """Dense [l, s1, s2, r] -> the block-packed ("QDense order") 1-D layout of include/chain_b200.h (host-side helper)."""
    parts = [np.asarray(phi[sl]).ravel(order="F") for sl in _phi_block_slices(ql, sq, qr, nq)]
    return np.concatenate(parts) if parts else np.zeros(0, phi.dtype)


def unpack_phi_blocks(xb, ql, sq, qr, nq):
    out = np.zeros((len(ql), len(sq), len(sq), len(qr)), dtype=np.asarray(xb).dtype, order="F")
    o = 0
    for sl in _phi_block_slices(ql, sq, qr, nq):
        shp = tuple(s.stop - s.start for s in sl)
        n = int(np.prod(shp))
        out[sl] = np.asarray(xb[o:o + n]).reshape(shp, order="F")
        o += n
    return out


def k_gemm(A, B, C0, M, N, Ks, a_offs, b_offs, lda, ldb, c_off, ldc, trans_a=False, trans_b=False, conj_a=False,
           accumulate=False, reps=0, ctx=None):
    ctx = ctx or default_context()
    cplx = np.iscomplexobj(A) or np.iscomplexobj(B)
    dt = np.complex128 if cplx else np.float64
    a = np.ascontiguousarray(A, dtype=dt).ravel()
    b = np.ascontiguousarray(B, dtype=dt).ravel()
    c = np.ascontiguousarray(C0, dtype=dt).ravel().copy()
    K = np.ascontiguousarray(Ks, dtype=np.int32)
    ao = np.ascontiguousarray(a_offs, dtype=np.int64)
    bo = np.ascontiguousarray(b_offs, dtype=np.int64)
    ms = C.c_double()
    ctx.check(ctx.lib.cb200_k_gemm(ctx.h, 1 if cplx else 0, int(trans_a), int(trans_b), int(conj_a), int(accumulate), M, N, len(K),
                                   i32ptr(K), ptr(a), a.size, ao.ctypes.data_as(C.POINTER(C.c_int64)), lda, ptr(b), b.size,
                                   bo.ctypes.data_as(C.POINTER(C.c_int64)), ldb, ptr(c), c.size, c_off, ldc, reps, C.byref(ms)))
    return c, ms.value


def k_jacobi_eig(G, tol=1e-14, ctx=None):
    ctx = ctx or default_context()
    cplx = np.iscomplexobj(G)
    dt = np.complex128 if cplx else np.float64
    g = np.ascontiguousarray(np.asarray(G, dtype=dt).transpose(0, 2, 1))   # each matrix column-major
    q = np.zeros_like(g)
    rot = C.c_int64()
    ctx.check(ctx.lib.cb200_k_jacobi_eig(ctx.h, 1 if cplx else 0, g.shape[0], ptr(g), ptr(q), tol, C.byref(rot)))
    return q.transpose(0, 2, 1), rot.value


def k_svd(X, ctx=None):
    ctx = ctx or default_context()
    cplx = np.iscomplexobj(X)
    dt = np.complex128 if cplx else np.float64
    rows, cols = X.shape
    x = _lib.fcol(X, dt).copy(order="F")
    v = np.zeros((cols, cols), dtype=dt, order="F")
    w = np.zeros(cols)
    sw = C.c_int32()
    ctx.check(ctx.lib.cb200_k_svd(ctx.h, 1 if cplx else 0, rows, cols, ptr(x), ptr(v), w.ctypes.data_as(C.POINTER(C.c_double)), C.byref(sw)))
    return x, v, w, sw.value


def k_heig(A, m, ctx=None):
    """Eigenvalues (descending) of the Hermitian matrix A and an orthonormal basis of its top-m invariant subspace."""
    ctx = ctx or default_context()
    cplx = np.iscomplexobj(A)
    dt = np.complex128 if cplx else np.float64
    n = A.shape[0]
    a = _lib.fcol(A, dt)
    w = np.zeros(n)
    v = np.zeros((n, max(m, 1)), dtype=dt, order="F")
    it = C.c_int32()
    ctx.check(ctx.lib.cb200_k_heig(ctx.h, 1 if cplx else 0, n, ptr(a), m, w.ctypes.data_as(C.POINTER(C.c_double)), ptr(v), C.byref(it)))
    return w, v[:, :m], it.value


def k_heig_batched(As, ms, ctx=None):
    """k_heig for several Hermitian matrices in one batched call (the charge blocks of one bond): [(w, V)]."""
    ctx = ctx or default_context()
    cplx = any(np.iscomplexobj(A) for A in As)
    dt = np.complex128 if cplx else np.float64
    k = len(As)
    a = [_lib.fcol(A, dt) for A in As]
    ns = (C.c_int64 * k)(*[x.shape[0] for x in a])
    mm = (C.c_int64 * k)(*[int(m) for m in ms])
    w = [np.zeros(x.shape[0]) for x in a]
    v = [np.zeros((x.shape[0], max(int(m), 1)), dtype=dt, order="F") for x, m in zip(a, ms)]
    ap = (C.c_void_p * k)(*[x.ctypes.data for x in a])
    wp = (C.POINTER(C.c_double) * k)(*[x.ctypes.data_as(C.POINTER(C.c_double)) for x in w])
    vp_ = (C.c_void_p * k)(*[x.ctypes.data for x in v])
    ctx.check(ctx.lib.cb200_k_heig_batched(ctx.h, 1 if cplx else 0, k, ns, ap, mm, wp, vp_))
    return [(wi, vi[:, :int(m)]) for wi, vi, m in zip(w, v, ms)]


def k_dots(xs, y, ctx=None):
    ctx = ctx or default_context()
    cplx = np.iscomplexobj(xs) or np.iscomplexobj(y)
    dt = np.complex128 if cplx else np.float64
    x = np.ascontiguousarray(xs, dtype=dt)
    yy = np.ascontiguousarray(y, dtype=dt)
    out = np.zeros(2 * x.shape[0])
    ctx.check(ctx.lib.cb200_k_dots(ctx.h, 1 if cplx else 0, x.shape[0], x.shape[1], ptr(x), ptr(yy), out.ctypes.data_as(C.POINTER(C.c_double))))
    return out[0::2] + 1j * out[1::2] if cplx else out[0::2]


def k_lincomb(xs, coefs, dst, beta=0.0, ctx=None):
    ctx = ctx or default_context()
    cplx = np.iscomplexobj(xs) or np.iscomplexobj(dst)
    dt = np.complex128 if cplx else np.float64
    x = np.ascontiguousarray(xs, dtype=dt)
    d = np.ascontiguousarray(dst, dtype=dt).copy()
    c = np.zeros(2 * x.shape[0])
    c[0::2] = np.real(coefs)
    c[1::2] = np.imag(coefs)
    ctx.check(ctx.lib.cb200_k_lincomb(ctx.h, 1 if cplx else 0, x.shape[0], x.shape[1], ptr(x), c.ctypes.data_as(C.POINTER(C.c_double)),
                                      float(np.real(beta)), float(np.imag(beta)), ptr(d)))
    return d


def k_gemm_peak(n=8192, reps=5, cplx=False, ctx=None):
    ctx = ctx or default_context()
    tf = C.c_double()
    ctx.check(ctx.lib.cb200_k_gemm_peak(ctx.h, 1 if cplx else 0, n, reps, C.byref(tf)))
    return tf.value


def k_level1_timed(n, nv=3, cplx=False, reps=5, ctx=None):
    """Mean device milliseconds of (multi_dot with nv vectors, lincomb with nv vectors, device copy) on n-element vectors."""
    ctx = ctx or default_context()
    ms = (C.c_double * 3)()
    ctx.check(ctx.lib.cb200_k_level1_timed(ctx.h, 1 if cplx else 0, nv, n, reps, ms))
    return list(ms)


def k_fp64_peak(which=0, iters=20000, ctx=None):
    """Issue-rate peak of the FP64 pipes (0 = DMMA.8x8x4, 1 = DFMA) in TFLOP/s."""
    ctx = ctx or default_context()
    tf = C.c_double()
    ctx.check(ctx.lib.cb200_k_fp64_peak(ctx.h, which, iters, C.byref(tf)))
    return tf.value
