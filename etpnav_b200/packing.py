"""Caller-side packing on B200: the tensors ``ETPTrainer`` builds in Python every step.

``pack_gmap`` replaces ``ETPTrainer._nav_gmap_variable`` (vlnce_baselines/ss_trainer_ETP.py:344-417) and
``pack_vp_features`` replaces ``_vp_feature_variable`` (:308-342); both return dictionaries with the reference's keys,
dtypes and shapes, already on the device.  The reference walks string-keyed dictionaries in O(B.N^2) Python and issues one
``.cuda()`` per tensor; here the host only FLATTENS the map state (positions, step ids, fronts, the node-to-node
shortest-distance table) into one pinned blob — no arithmetic — one H2D copy moves it, and ``etp_gmap_pack`` computes
``gmap_pos_fts`` (``GraphMap.get_pos_fts``, models/graph_utils.py:278-322), ``gmap_pair_dists`` (:371-387), step ids
and both masks for the whole batch in one launch.  The image features of the map (``get_node_embeds`` stack / [stop]
row / padding, :362-366,399) and the view features are row gathers: ``etp_segment_gather`` over a pool of the tensors
the policy already holds on the device (differentiable, so the averaged panorama embeddings keep their graph).
No CPU / PyTorch fallback: without the library or a B200 the calls raise.

``GmapPacker`` is ``pack_gmap`` with memory: it mirrors every GraphMap of the batch in flat arrays between calls and
re-reads only what ``update_graph`` / ``delete_ghost`` changed (pure-Python ``_EnvMirror`` or, when built, the CPython-API
helper ``csrc_py/gmap_mirror.c``), stages one pinned blob, one H2D copy, two launches — same tensors, bit for bit.
"""
import ctypes as C
from itertools import accumulate, chain
from operator import itemgetter

import numpy as np
import torch

from . import lib as _L
from .pretrain import Csr, segment_gather

p_void, i32 = C.c_void_p, C.c_int32
_declared = False


def _declare():
    global _declared
    if not _declared:
        _L.lib().etp_gmap_pack.argtypes = [p_void, p_void, p_void, i32, i32, i32, p_void, p_void, p_void, p_void, p_void,
                                           p_void]
        _L.lib().etp_segment_gather_rows.argtypes = [p_void, p_void, p_void, p_void, i32, i32, p_void, p_void]
        _declared = True


def heading_from_quaternion(coeffs):
    """``heading_from_quaternion`` (models/graph_utils.py:53-58) without habitat: coefficients are [x, y, z, w]
    (``quaternion_from_coeff``); rotate (0, 0, -1) by the inverse quaternion and take the polar angle
    ``arctan2(hv.x, -hv.z)`` (``cartesian_to_polar``), modulo 2 pi.  Scalar bookkeeping of one pose per environment."""
    x, y, z, w = (float(c) for c in coeffs)
    n = w * w + x * x + y * y + z * z
    iw, ix, iy, iz = w / n, -x / n, -y / n, -z / n          # q^-1
    # (q^-1) * (0, v) * q   with v = (0, 0, -1)
    aw, ax, ay, az = iz, -iy, ix, -iw                         # q^-1 * (0, 0, 0, -1)
    hx = aw * x + ax * w + ay * z - az * y
    hz = aw * z + ax * y - ay * x + az * w
    return float(np.arctan2(hx, -hz) % (2 * np.pi))


def flatten_gmaps(gmaps, cur_vp, cur_pos, cur_ori):
    """GraphMap objects (models/graph_utils.py:133; anything with ``node_pos, ghost_pos, ghost_aug_pos, node_stepId,
    ghost_fronts, shortest_dist, shortest_path``) -> (meta int32 [B,8], f64 blob, i32 blob, vp id lists, n_max,
    max_ghosts) in the layout ``etp_gmap_pack`` documents (include/etpnav_b200.h).  Pure re-layout: the values of the whole
    batch are appended to two Python lists by C-level row getters (``itemgetter`` / ``map`` / ``chain``) and converted to
    arrays ONCE, instead of one numpy conversion per table row."""
    B = len(gmaps)
    meta = np.zeros((B, 8), dtype=np.int32)
    fvals, ivals, vp_ids = [], [], []
    fext, iext = fvals.extend, ivals.extend
    n_max, max_g = 1, 0
    for e, gm in enumerate(gmaps):
        nid, gid = list(gm.node_pos), list(gm.ghost_pos)
        n, g = len(nid), len(gid)
        ix = dict(zip(nid, range(n)))
        # row getters run the per-row dictionary reads in C (itemgetter with one key returns a scalar: wrap it)
        row = itemgetter(*nid) if n > 1 else (lambda d, k=nid[0]: (d[k],))
        off_f, off_i = len(fvals), len(ivals)
        cp = cur_pos[e]
        fext((float(cp[0]), float(cp[1]), float(cp[2]), heading_from_quaternion(cur_ori[e])))
        fext(np.asarray(row(gm.node_pos), dtype=np.float64).ravel().tolist())
        if g:
            fext(np.asarray(itemgetter(*gid)(gm.ghost_aug_pos) if g > 1 else (gm.ghost_aug_pos[gid[0]],),
                            dtype=np.float64).ravel().tolist())
        sd, sp = gm.shortest_dist, gm.shortest_path
        for a in nid:
            fext(row(sd[a]))
        iext(row(gm.node_stepId))
        ivals.append(0)
        nnz = 0
        if g:
            fronts = list(map(gm.ghost_fronts.__getitem__, gid))         # lists of node ids, one per ghost
            lens = list(accumulate(map(len, fronts)))
            nnz = lens[-1]
            iext(lens)
            iext(map(ix.__getitem__, chain.from_iterable(fronts)))
        for a in nid:
            iext(map(len, row(sp[a])))
        meta[e, :6] = (n, g, ix[cur_vp[e]], off_f, off_i, nnz)
        n_max, max_g = max(n_max, 1 + n + g), max(max_g, g)
        vp_ids.append([None] + nid + gid)
    return meta, np.array(fvals, dtype=np.float64), np.array(ivals, dtype=np.int32), vp_ids, n_max, max_g


def _to_device(arr, dtype, device):
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(dtype)
    if device.type == "cuda":
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


def pack_gmap_geometry(meta, f64, i32b, n_max, max_ghosts, device):
    """One H2D copy per blob + one ``etp_gmap_pack`` launch -> the five geometry tensors, padded to n_max."""
    _L.require_device()
    _declare()
    device = torch.device(device)
    B = meta.shape[0]
    d_meta, d_f64, d_i32 = _to_device(meta, torch.int32, device), _to_device(f64, torch.float64, device), \
        _to_device(i32b, torch.int32, device)
    step_ids = torch.empty(B, n_max, dtype=torch.int64, device=device)
    visited = torch.empty(B, n_max, dtype=torch.uint8, device=device)
    masks = torch.empty(B, n_max, dtype=torch.uint8, device=device)
    pos = torch.empty(B, n_max, 7, dtype=torch.float32, device=device)
    pd = torch.empty(B, n_max, n_max, dtype=torch.float32, device=device)
    _L._check(_L.lib().etp_gmap_pack(_L.ptr(d_meta), _L.ptr(d_f64), _L.ptr(d_i32), B, n_max, max_ghosts, _L.ptr(step_ids),
                                     _L.ptr(visited), _L.ptr(masks), _L.ptr(pos), _L.ptr(pd), _L.stream_ptr()),
              "etp_gmap_pack")
    return dict(gmap_step_ids=step_ids, gmap_visited_masks=visited.view(torch.bool), gmap_masks=masks.view(torch.bool),
                gmap_pos_fts=pos, gmap_pair_dists=pd)


def pack_gmap_img_fts(gmaps, n_max, device):
    """``gmap_img_fts`` (ss_trainer_ETP.py:362-366,399): [stop] row of zeros, node embeddings, ghost embeddings
    (running sum / count, ``get_node_embeds`` graph_utils.py:272-276), zero padding — one gather with weights 1 or
    1/count over the pool of tensors the maps already hold."""
    rows, ptr, idx, wt = [], [0], [], []
    for gm in gmaps:
        k = 1
        ptr.append(len(idx))  # [stop]
        for vp in gm.node_pos.keys():
            idx.append(len(rows)); wt.append(1.0); rows.append(gm.node_embeds[vp]); ptr.append(len(idx)); k += 1
        for vp in gm.ghost_pos.keys():
            idx.append(len(rows)); wt.append(1.0 / gm.ghost_embeds[vp][1]); rows.append(gm.ghost_embeds[vp][0])
            ptr.append(len(idx)); k += 1
        ptr.extend([len(idx)] * (n_max - k))
    csr = Csr(np.asarray(ptr, dtype=np.int32), np.asarray(idx, dtype=np.int32), np.asarray(wt, dtype=np.float32), len(rows))
    device = torch.device(device)
    no_grad = not torch.is_grad_enabled() or not any(r.requires_grad for r in rows)
    if no_grad and device.type == "cuda" and all(r.is_cuda and r.dtype == torch.float32 and r.is_contiguous() for r in rows):
        # inference: read every node / ghost tensor where it lives (table of row pointers), no stacking copy
        _declare()
        table = _to_device(np.fromiter((r.data_ptr() for r in rows), dtype=np.int64, count=len(rows)), torch.int64, device)
        d_ptr, d_idx, d_wt = csr.to(device)
        out = torch.empty(csr.num_segments, rows[0].shape[0], device=device, dtype=torch.float32)
        _L._check(_L.lib().etp_segment_gather_rows(_L.ptr(table), _L.ptr(d_ptr), _L.ptr(d_idx), _L.ptr(d_wt),
                                                   csr.num_segments, rows[0].shape[0], _L.ptr(out), _L.stream_ptr()),
                  "etp_segment_gather_rows")
        return out.view(len(gmaps), n_max, rows[0].shape[0])
    pool = torch.stack(rows, 0).to(device)
    return segment_gather(pool, csr).view(len(gmaps), n_max, pool.shape[1])


def pack_gmap(gmaps, cur_vp, cur_pos, cur_ori, device="cuda"):
    """Drop-in for ``ETPTrainer._nav_gmap_variable(cur_vp, cur_pos, cur_ori)`` given ``self.gmaps``: same keys."""
    meta, f64, i32b, vp_ids, n_max, max_g = flatten_gmaps(gmaps, cur_vp, cur_pos, cur_ori)
    out = pack_gmap_geometry(meta, f64, i32b, n_max, max_g, device)
    out["gmap_vp_ids"] = vp_ids
    out["gmap_img_fts"] = pack_gmap_img_fts(gmaps, n_max, device)
    out["no_vp_left"] = [len(gm.ghost_pos) == 0 for gm in gmaps]
    return out


def pack_vp_features(obs, device="cuda"):
    """Drop-in for ``ETPTrainer._vp_feature_variable(obs)`` (ss_trainer_ETP.py:308-342): candidate views first, then the
    panorama views that are not candidates; zero-padded to the longest row.  Three gathers (rgb 512, depth 128, angle 4)
    over pools made of the candidate tensors followed by the panorama tensor."""
    device = torch.device(device)
    B = len(obs["cand_rgb"])
    n_cand = [len(x) for x in obs["cand_angles"]]
    keep = []
    for i in range(B):
        is_cand = np.zeros(12, dtype=bool)
        is_cand[np.asarray(obs["cand_img_idxes"][i], dtype=np.int64)] = True
        keep.append(np.nonzero(~is_cand)[0])
    lens = [n_cand[i] + len(keep[i]) for i in range(B)]
    V = max(lens)
    c0 = np.concatenate([[0], np.cumsum(n_cand)]).astype(np.int64)
    total_c = int(c0[-1])
    nav = np.zeros((B, V), dtype=np.int64)

    def index(pano_shared):
        idx = []
        for i in range(B):
            base = total_c + (0 if pano_shared else 12 * i)
            idx.append(np.concatenate([np.arange(c0[i], c0[i + 1]), base + keep[i],
                                       np.full(V - lens[i], -1, dtype=np.int64)]))
        idx = np.concatenate(idx)
        valid = idx >= 0
        ptr = np.concatenate([[0], np.cumsum(valid)]).astype(np.int32)
        return ptr, idx[valid].astype(np.int32)

    def gather(cands, pano, pano_shared):
        pool = torch.cat([c.to(device).float() for c in cands] + [pano.to(device).float().reshape(-1, pano.shape[-1])], 0)
        ptr, idx = index(pano_shared)
        csr = Csr(ptr, idx, np.ones(len(idx), dtype=np.float32), pool.shape[0])
        return segment_gather(pool, csr).view(B, V, pool.shape[1])

    for i in range(B):
        nav[i, :n_cand[i]] = 1
    return dict(rgb_fts=gather(obs["cand_rgb"], obs["pano_rgb"], False),
                dep_fts=gather(obs["cand_depth"], obs["pano_depth"], False),
                loc_fts=gather(obs["cand_angle_fts"], obs["pano_angle_fts"], True),
                nav_types=torch.from_numpy(nav).to(device), view_lens=torch.tensor(lens, dtype=torch.long, device=device))


class _EnvMirror:
    """Array-backed mirror of ONE GraphMap (models/graph_utils.py:133), brought up to date by ``sync``.

    What ``GraphMap.update_graph`` (:185-250) can change between two packs, and what is re-read for it:
      * one node is appended (ids are never reused, positions / step ids of older nodes never change): the new row of
        ``node_pos`` / ``node_stepId`` / ``node_embeds`` only;
      * ``shortest_dist`` / ``shortest_path`` are rebuilt as NEW dictionaries (so an unchanged object = no update since
        the last pack = nothing to read).  If the new node has degree 1 in ``graph_nx`` no path between two older nodes
        can run through it, so only its row and column are read; a loop closure (degree > 1) re-reads the n x n tables;
      * ghosts are created, gain a front (``ghost_fronts`` / ``ghost_embeds`` / ``ghost_pos`` grow together) or are
        deleted (``delete_ghost`` :178-183): detected from the id list and the per-ghost front counts, both compared at C
        speed; only the ghosts whose count moved have their embedding pointer / weight refreshed;
      * ``ghost_aug_pos`` is redrawn for every ghost inside ``update_graph`` (:238-244): re-read after an update."""
    __slots__ = ("gm", "sd_obj", "nid", "ix", "n", "cap", "npos", "steps", "sd", "spl", "gids", "flens", "fcum", "fidx",
                 "gpos", "nptr", "nkeep", "gl", "gp", "gw", "gkeep", "gptr", "gwt", "img_ok", "img_gids", "img_lens", "vp_ids")

    def __init__(self, gm, cap=16):
        self.gm, self.sd_obj, self.nid, self.ix, self.n = gm, None, [], {}, 0
        self._alloc(cap)
        self.gids, self.flens = None, None
        self.fcum = self.fidx = np.zeros(0, dtype=np.int32)
        self.gpos = np.zeros((0, 3), dtype=np.float64)
        self.nptr, self.nkeep, self.gl, self.gp, self.gw, self.gkeep, self.gptr, self.gwt = [], [], {}, {}, {}, {}, [], []
        self.img_ok, self.img_gids, self.img_lens, self.vp_ids = True, None, None, None

    def _alloc(self, cap):
        old = (self.npos, self.steps, self.sd, self.spl) if self.n else None
        self.cap = cap
        self.npos = np.zeros((cap, 3), dtype=np.float64)
        self.steps = np.zeros(cap + 1, dtype=np.int32)          # [n] stays 0: front_ptr[0] of the i32 layout
        self.sd = np.zeros((cap, cap), dtype=np.float64)
        self.spl = np.zeros((cap, cap), dtype=np.int32)
        if old is not None:
            n = self.n
            self.npos[:n], self.steps[:n], self.sd[:n, :n], self.spl[:n, :n] = old[0][:n], old[1][:n], old[2][:n, :n], old[3][:n, :n]

    def sync(self, gm, want_img, width, device):
        sd = gm.shortest_dist
        updated = sd is not self.sd_obj
        if updated:
            nid = list(gm.node_pos)
            n0, n = self.n, len(nid)
            if n0 and nid[:n0] != self.nid:     # not the append-only history update_graph produces: start over
                self.__init__(gm, self.cap)
                n0 = 0
            if n > self.cap:
                self._alloc(max(n, 2 * self.cap))
            npos, nstep, ix = gm.node_pos, gm.node_stepId, self.ix
            for k in range(n0, n):
                v = nid[k]
                ix[v] = k
                self.npos[k] = npos[v]
                self.steps[k] = nstep[v]
            self.steps[n] = 0
            sp = gm.shortest_path
            G = getattr(gm, "graph_nx", None)
            if n0 and n == n0 + 1 and G is not None and len(G[nid[n0]]) == 1:
                v, old = nid[n0], self.nid
                row = itemgetter(*nid)
                self.sd[n0, :n] = row(sd[v])
                self.sd[:n0, n0] = [sd[a][v] for a in old]
                self.spl[n0, :n] = list(map(len, row(sp[v])))
                self.spl[:n0, n0] = [len(sp[a][v]) for a in old]
            else:
                row = itemgetter(*nid) if n > 1 else (lambda d, k=nid[0]: (d[k],))
                self.sd[:n, :n] = [row(sd[a]) for a in nid]
                self.spl[:n, :n] = [list(map(len, row(sp[a]))) for a in nid]
            self.nid, self.n, self.sd_obj = nid, n, sd
        n = self.n
        gids = list(gm.ghost_pos)
        g = len(gids)
        if g:
            fronts = itemgetter(*gids)(gm.ghost_fronts) if g > 1 else (gm.ghost_fronts[gids[0]],)
            lens = list(map(len, fronts))
        else:
            fronts, lens = (), []
        same_ids = gids == self.gids
        ghosts_changed = not same_ids or lens != self.flens
        if ghosts_changed:
            self.fcum = np.fromiter(accumulate(lens), dtype=np.int32, count=g)
            self.fidx = np.fromiter(map(self.ix.__getitem__, chain.from_iterable(fronts)), dtype=np.int32)
        if updated or not same_ids:
            self.gpos = np.array(itemgetter(*gids)(gm.ghost_aug_pos) if g > 1 else
                                 [gm.ghost_aug_pos[v] for v in gids], dtype=np.float64).reshape(g, 3)
        if updated or not same_ids or self.vp_ids is None:
            self.vp_ids = [None] + self.nid + gids
        if want_img and self.img_ok:
            grad = torch.is_grad_enabled()

            def usable(t):
                return (t.device == device and t.dtype == torch.float32 and t.dim() == 1 and t.shape[0] == width
                        and t.is_contiguous() and not (grad and t.requires_grad))
            ne = gm.node_embeds
            for k in range(len(self.nptr), n):
                t = ne[self.nid[k]]
                if not usable(t):
                    self.img_ok = False
                    break
                self.nptr.append(t.data_ptr())
                self.nkeep.append(t)
            if self.img_ok and (gids != self.img_gids or lens != self.img_lens):
                ge, gl, gp, gw, keep = gm.ghost_embeds, self.gl, self.gp, self.gw, self.gkeep
                for v, l in zip(gids, lens):
                    if gl.get(v) != l:
                        t, c = ge[v]
                        if not usable(t):
                            self.img_ok = False
                            break
                        gp[v], gw[v], gl[v], keep[v] = t.data_ptr(), 1.0 / c, l, t
                if self.img_ok:
                    self.gptr = list(map(gp.__getitem__, gids)) if g else []
                    self.gwt = list(map(gw.__getitem__, gids)) if g else []
                    self.img_gids, self.img_lens = gids, lens
        self.gids, self.flens = gids, lens
        return g


class GmapPacker:
    """``pack_gmap`` with memory: drop-in for ``ETPTrainer._nav_gmap_variable`` (ss_trainer_ETP.py:344-417) that keeps an
    array-backed mirror of every environment's GraphMap between calls (``_EnvMirror``), so a step re-reads only what
    ``update_graph`` / ``delete_ghost`` changed instead of walking every dictionary of every map again.  Same outputs,
    bit for bit, as the stateless ``pack_gmap`` (tests/test_packing_incremental_cpu.py, tests/test_packing_gpu.py).

    Per call: the mirrors are synced, the blobs ``etp_gmap_pack`` documents are concatenated straight into ONE pinned
    staging buffer (two buffers alternate, so the host never waits for the previous copy), one H2D copy moves it, and two
    launches (``etp_gmap_pack``, ``etp_segment_gather_rows`` over the cached row-pointer table) write the six tensors.
    Mirrors follow the GraphMap OBJECTS (environments finish and drop out of the batch; a new episode builds new maps)."""

    def __init__(self, device="cuda", width=768, impl=None):
        """``impl``: "c" = the CPython-API host half (etpnav_b200/_gmap_mirror.so, csrc_py/gmap_mirror.c: the same mirror,
        dictionary walks in C, one call per pack), "py" = the pure-Python twin (``_EnvMirror``); None picks "c" when the
        helper is built.  Both produce the same bytes (tests/test_packing_incremental_cpu.py)."""
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())      # tensors report cuda:N, never bare "cuda"
        self.device, self.width = dev, width
        self._c = _load_mirror_helper() if impl in (None, "c") else None
        if impl == "c" and self._c is None:
            raise RuntimeError("etpnav_b200/_gmap_mirror.so is not built (python -m etpnav_b200.build)")
        self.impl = "c" if self._c is not None else "py"
        self._mirrors = {}
        self._pin, self._pin_np, self._pin_ev, self._turn = [None, None], [None, None], [None, None], 0

    def reset(self):
        self._mirrors = {}

    def _sync_c(self, gmaps, want_img):
        old, new, caps = self._mirrors, {}, []
        for gm in gmaps:
            ent = old.get(id(gm))
            if ent is None or ent[0] is not gm:
                ent = (gm, self._c.etp_pm_new())
            new[id(gm)] = ent
            caps.append(ent[1])
        self._mirrors = new
        ctx = (self.device, torch.float32, self.width, torch.is_grad_enabled())      # rows the image gather may read in place
        return caps, self._c.etp_pm_sync(caps, list(gmaps), int(want_img), ctx)

    @staticmethod
    def _poses(cur_pos, cur_ori):
        out = np.empty((len(cur_pos), 4), dtype=np.float64)
        for e, cp in enumerate(cur_pos):
            out[e] = (float(cp[0]), float(cp[1]), float(cp[2]), heading_from_quaternion(cur_ori[e]))
        return out

    # ---- host side -------------------------------------------------------------------------------------------------
    def _sync(self, gmaps, want_img):
        old, new, sts = self._mirrors, {}, []
        for gm in gmaps:
            st = old.get(id(gm))
            if st is None or st.gm is not gm:
                st = _EnvMirror(gm)
            st.sync(gm, want_img, self.width, self.device)
            new[id(gm)] = st
            sts.append(st)
        self._mirrors = new            # maps that left the batch are forgotten (and their kept tensors released)
        return sts

    @staticmethod
    def _layout(sts, cur_vp):
        B = len(sts)
        ns = np.fromiter((st.n for st in sts), dtype=np.int64, count=B)
        gs = np.fromiter((len(st.gids) for st in sts), dtype=np.int64, count=B)
        nnz = np.fromiter((len(st.fidx) for st in sts), dtype=np.int64, count=B)
        f_len = 4 + 3 * ns + 3 * gs + ns * ns
        i_len = ns + 1 + gs + nnz + ns * ns
        meta = np.zeros((B, 8), dtype=np.int32)
        meta[:, 0], meta[:, 1], meta[:, 5] = ns, gs, nnz
        meta[:, 2] = [st.ix[v] for st, v in zip(sts, cur_vp)]
        meta[1:, 3], meta[1:, 4] = np.cumsum(f_len)[:-1], np.cumsum(i_len)[:-1]
        return meta, int(f_len.sum()), int(i_len.sum()), int(max(1, (1 + ns + gs).max())), int(gs.max())

    @staticmethod
    def _fill(sts, meta, fout, iout, cur_pos, cur_ori):
        """Write every environment's slice of the two blobs (layout: include/etpnav_b200.h, etp_gmap_pack) in place."""
        offs_f, offs_i = meta[:, 3].tolist(), meta[:, 4].tolist()
        for e, st in enumerate(sts):
            n, g, cp = st.n, len(st.gids), cur_pos[e]
            a = offs_f[e]
            fout[a:a + 4] = (float(cp[0]), float(cp[1]), float(cp[2]), heading_from_quaternion(cur_ori[e]))
            a += 4
            fout[a:a + 3 * n] = st.npos[:n].reshape(-1)
            a += 3 * n
            if g:
                fout[a:a + 3 * g] = st.gpos.reshape(-1)
                a += 3 * g
            fout[a:a + n * n].reshape(n, n)[...] = st.sd[:n, :n]
            a = offs_i[e]
            iout[a:a + n + 1] = st.steps[:n + 1]
            a += n + 1
            if g:
                k = len(st.fidx)
                iout[a:a + g] = st.fcum
                iout[a + g:a + g + k] = st.fidx
                a += g + k
            iout[a:a + n * n].reshape(n, n)[...] = st.spl[:n, :n]

    def flatten(self, gmaps, cur_vp, cur_pos, cur_ori):
        """Host half only: the tuple ``flatten_gmaps`` returns (used by the CPU tests to compare the two)."""
        if self._c is not None:
            caps, (n_max, max_g, nf, ni, _rows, _ok) = self._sync_c(gmaps, False)
            meta = np.zeros((len(gmaps), 8), dtype=np.int32)
            fout, iout = np.empty(nf, dtype=np.float64), np.empty(ni, dtype=np.int32)
            poses = self._poses(cur_pos, cur_ori)
            self._c.etp_pm_fill(caps, list(cur_vp), poses.ctypes.data, meta.ctypes.data, fout.ctypes.data, iout.ctypes.data,
                                None, None, None, None, n_max)
            return meta, fout, iout, self._c.etp_pm_vp_ids(caps)[0], n_max, max_g
        sts = self._sync(gmaps, False)
        meta, nf, ni, n_max, max_g = self._layout(sts, cur_vp)
        fout, iout = np.empty(nf, dtype=np.float64), np.empty(ni, dtype=np.int32)
        self._fill(sts, meta, fout, iout, cur_pos, cur_ori)
        return meta, fout, iout, [list(st.vp_ids) for st in sts], n_max, max_g

    def host_tables(self, gmaps, cur_vp, cur_pos, cur_ori):
        """(meta, f64 blob, i32 blob, row-pointer table, csr ptr, csr idx, csr weights, all-fast flag) as numpy arrays: what
        ``pack`` stages for the device, without a device (tests)."""
        B = len(gmaps)
        if self._c is not None:
            caps, (n_max, max_g, nf, ni, rows, ok) = self._sync_c(gmaps, True)
            meta, fout, iout = np.zeros((B, 8), dtype=np.int32), np.empty(nf, dtype=np.float64), np.empty(ni, dtype=np.int32)
            table, cptr = np.empty(rows, dtype=np.int64), np.empty(B * n_max + 1, dtype=np.int32)
            cidx, cwt = np.empty(rows, dtype=np.int32), np.empty(rows, dtype=np.float32)
            poses = self._poses(cur_pos, cur_ori)
            tb = (table.ctypes.data, cptr.ctypes.data, cidx.ctypes.data, cwt.ctypes.data) if ok else (None,) * 4
            self._c.etp_pm_fill(caps, list(cur_vp), poses.ctypes.data, meta.ctypes.data, fout.ctypes.data, iout.ctypes.data,
                                *tb, n_max)
            return meta, fout, iout, table, cptr, cidx, cwt, bool(ok)
        sts = self._sync(gmaps, True)
        meta, nf, ni, n_max, max_g = self._layout(sts, cur_vp)
        fout, iout = np.empty(nf, dtype=np.float64), np.empty(ni, dtype=np.int32)
        self._fill(sts, meta, fout, iout, cur_pos, cur_ori)
        ok = all(st.img_ok for st in sts)
        return (meta, fout, iout) + (self.img_tables(sts, n_max) if ok else (None,) * 4) + (ok,)

    def img_tables(self, sts, n_max):
        """Row-pointer table + CSR (``etp_segment_gather_rows``) of ``gmap_img_fts``: [stop] (empty segment), one row per
        node (weight 1) and ghost (weight 1 / count, ``get_node_embeds`` graph_utils.py:272-276), empty padding."""
        B = len(sts)
        k = np.fromiter((len(st.nptr) + len(st.gptr) for st in sts), dtype=np.int64, count=B)
        base = np.concatenate([[0], np.cumsum(k)])
        total = int(base[-1])
        table = np.fromiter(chain.from_iterable(chain(st.nptr, st.gptr) for st in sts), dtype=np.int64, count=total)
        wt = np.ones(total, dtype=np.float32)
        for e, st in enumerate(sts):
            if st.gwt:
                wt[base[e] + len(st.nptr):base[e + 1]] = st.gwt
        j = np.arange(n_max, dtype=np.int64)[None, :] - 1
        ptr = np.empty(B * n_max + 1, dtype=np.int32)
        ptr[:-1] = (base[:-1, None] + np.clip(j, 0, k[:, None])).ravel()
        ptr[-1] = total
        return table, ptr, np.arange(total, dtype=np.int32), wt

    # ---- device side -----------------------------------------------------------------------------------------------
    def _staging(self, nbytes):
        t = self._turn = self._turn ^ 1
        if self._pin[t] is None or self._pin[t].numel() < nbytes:
            self._pin[t] = torch.empty(max(nbytes * 2, 1 << 20), dtype=torch.uint8).pin_memory()
            self._pin_np[t] = self._pin[t].numpy()
            self._pin_ev[t] = torch.cuda.Event()
        else:
            self._pin_ev[t].synchronize()      # the copy issued two packs ago has long finished
        return self._pin[t], self._pin_np[t], self._pin_ev[t]

    def pack(self, gmaps, cur_vp, cur_pos, cur_ori):
        _L.require_device()
        _declare()
        dev, B, W = self.device, len(gmaps), self.width
        if self._c is not None:
            caps, (n_max, max_g, nf, ni, rows, img_fast) = self._sync_c(gmaps, True)
            nrow = rows if img_fast else 0
            sizes = [nf * 8, nrow * 8, B * 32, ni * 4, (B * n_max + 1) * 4 if img_fast else 0, nrow * 4, nrow * 4]
        else:
            sts = self._sync(gmaps, True)
            meta, nf, ni, n_max, max_g = self._layout(sts, cur_vp)
            img_fast = all(st.img_ok for st in sts)
            if img_fast:
                table, ptr, idx, wt = self.img_tables(sts, n_max)
            else:
                table, ptr, idx, wt = (np.zeros(0, dtype=t) for t in (np.int64, np.int32, np.int32, np.float32))
            sizes = [nf * 8, table.nbytes, meta.nbytes, ni * 4, ptr.nbytes, idx.nbytes, wt.nbytes]
        # one blob (8-byte sections first): f64 | row pointers | meta | i32 | csr ptr | csr idx | csr weights
        offs = [0]
        for sz in sizes:
            offs.append((offs[-1] + sz + 15) // 16 * 16)
        pin, pnp, ev = self._staging(offs[-1])
        if self._c is not None:
            base = pin.data_ptr()
            poses = self._poses(cur_pos, cur_ori)
            tb = [base + offs[k] for k in (1, 4, 5, 6)] if img_fast else [None] * 4
            self._c.etp_pm_fill(caps, list(cur_vp), poses.ctypes.data, base + offs[2], base + offs[0], base + offs[3], *tb, n_max)
            vp_ids, no_vp_left = self._c.etp_pm_vp_ids(caps)
        else:
            self._fill(sts, meta, pnp[offs[0]:offs[0] + nf * 8].view(np.float64), pnp[offs[3]:offs[3] + ni * 4].view(np.int32),
                       cur_pos, cur_ori)
            for o, a in ((offs[1], table), (offs[2], meta), (offs[4], ptr), (offs[5], idx), (offs[6], wt)):
                pnp[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
            vp_ids, no_vp_left = [list(st.vp_ids) for st in sts], [not st.gids for st in sts]
        blob = torch.empty(offs[-1], dtype=torch.uint8, device=dev)
        blob.copy_(pin[:offs[-1]], non_blocking=True)
        ev.record()
        base = blob.data_ptr()
        P = lambda k: C.c_void_p(base + offs[k])
        step_ids = torch.empty(B, n_max, dtype=torch.int64, device=dev)
        visited = torch.empty(B, n_max, dtype=torch.uint8, device=dev)
        masks = torch.empty(B, n_max, dtype=torch.uint8, device=dev)
        pos = torch.empty(B, n_max, 7, dtype=torch.float32, device=dev)
        pd = torch.empty(B, n_max, n_max, dtype=torch.float32, device=dev)
        lib, sp = _L.lib(), _L.stream_ptr()
        _L._check(lib.etp_gmap_pack(P(2), P(0), P(3), B, n_max, max_g, _L.ptr(step_ids), _L.ptr(visited), _L.ptr(masks),
                                    _L.ptr(pos), _L.ptr(pd), sp), "etp_gmap_pack")
        if img_fast:
            img = torch.empty(B, n_max, W, dtype=torch.float32, device=dev)
            _L._check(lib.etp_segment_gather_rows(P(1), P(4), P(5), P(6), B * n_max, W, _L.ptr(img), sp),
                      "etp_segment_gather_rows")
        else:   # some embedding needs a gradient (training) or lives elsewhere: the differentiable gather of pack_gmap
            img = pack_gmap_img_fts(gmaps, n_max, dev)
        self._last_blob = blob     # stays referenced until the next pack: the launches above read it asynchronously
        return dict(gmap_step_ids=step_ids, gmap_visited_masks=visited.view(torch.bool), gmap_masks=masks.view(torch.bool),
                    gmap_pos_fts=pos, gmap_pair_dists=pd, gmap_vp_ids=vp_ids, gmap_img_fts=img, no_vp_left=no_vp_left)


_mirror_helper = None


def _load_mirror_helper():
    """ctypes.PyDLL handle of etpnav_b200/_gmap_mirror.so (the functions take and return Python objects and run under the
    GIL), or None when it has not been built."""
    global _mirror_helper
    if _mirror_helper is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_gmap_mirror.so")
        if not os.path.exists(path):
            return None
        h = C.PyDLL(path)
        h.etp_pm_new.restype, h.etp_pm_new.argtypes = C.py_object, []
        h.etp_pm_sync.restype, h.etp_pm_sync.argtypes = C.py_object, [C.py_object, C.py_object, C.c_int, C.py_object]
        h.etp_pm_fill.restype = C.py_object
        h.etp_pm_fill.argtypes = [C.py_object, C.py_object] + [C.c_void_p] * 8 + [C.c_int]
        h.etp_pm_vp_ids.restype, h.etp_pm_vp_ids.argtypes = C.py_object, [C.py_object]
        _mirror_helper = h
    return _mirror_helper
