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
