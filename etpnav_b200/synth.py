"""Seeded synthetic weights and inputs for the planner hot path.

There is no dataset or checkpoint on the GPU box, so parity tests, ``smoke()`` and ``bench.py``
all draw from these generators.  Everything is produced on the CPU from a ``torch.Generator``
with a fixed seed, so the same tensors can be rebuilt bit-for-bit here, on the GPU box and
by ``oracle/make_golden.py`` (the golden fixtures store outputs only).

Shapes and value laws follow what the reference's caller builds:
``ss_trainer_ETP.py:308-342`` (_vp_feature_variable), ``:344-417`` (_nav_gmap_variable),
``graph_utils.py:278-322`` (get_pos_fts) and ``models/utils.py:49-57`` (angle_feature_torch).
"""
import math
from collections import OrderedDict

import torch

from .config import PlannerConfig
from .spec import param_shapes


def make_weights(cfg: PlannerConfig, seed: int = 0, skip_text: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Explicit fp32 weights for every key of the reference state_dict.

    Linear / embedding weights ~ N(0, 0.02) (HF ``_init_weights``); pano
    ``in_proj_weight`` ~ N(0, 0.0255) (measured xavier std); LayerNorm gamma = 1 + N(0, 0.05),
    beta and all biases ~ N(0, 0.02) so that every affine / bias path is exercised.
    ``sprel_linear`` gets w=0.7, b=-0.1 plus noise so the graph bias matters.
    With ``skip_text`` the (large) ``embeddings.word/position`` and ``lang_encoder`` tensors
    are left out (the per-step hot path does not read them).
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        if skip_text and (name.startswith("lang_encoder.") or name.startswith("embeddings.word")
                          or name.startswith("embeddings.position") or name.startswith("embeddings.LayerNorm")):
            continue
        is_ln = ("LayerNorm" in name or "layer_norm" in name or ".norm" in name
                 or name.startswith("global_encoder.gmap_pos_embeddings.1")
                 or name.startswith("global_sap_head.net.2"))
        if name.startswith("global_encoder.sprel_linear"):
            base = 0.7 if name.endswith("weight") else -0.1
            t = torch.full(shape, base) + 0.01 * torch.randn(shape, generator=g)
        elif is_ln and name.endswith("weight"):
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("in_proj_weight"):
            t = 0.0255 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[name] = t.float().contiguous()
    if "embeddings.word_embeddings.weight" in sd:
        sd["embeddings.word_embeddings.weight"][0].zero_()  # padding_idx=0 (vilmodel_cmt.py:52)
    return sd


def _lens(g, B, lo, hi):
    lens = torch.randint(lo, hi + 1, (B,), generator=g)
    lens[0] = hi  # at least one full row: the padded width equals max(lens) (common/ops.py:38)
    return lens


def r2r_text_lengths(g, B: int, L: int):
    """R2R-CE-like instruction lengths in BERT tokens (BASELINE.json configs[3]): the dataset is not in the
    reference repository, so the law is synthetic and stated here — normal(mean 32, sd 12) clipped to [8, min(80, L)]
    (``IL.max_text_len`` = 80, run_r2r/iter_train.yaml:42); one row takes the maximum so the padded width is reached."""
    hi = min(80, L)
    lens = (torch.randn(B, generator=g) * 12.0 + 32.0).round().long().clamp_(8, hi)
    lens[0] = hi
    return lens


def make_inputs(cfg: PlannerConfig, B: int, V: int, N: int, L: int, seed: int = 0,
                ragged: bool = True, txt_from: str = "normal", pad_id: int = 0, txt_law: str = None) -> dict:
    """One planner step's inputs (CPU tensors).

    ``ragged=False`` gives fixed V/N/L (peak-throughput shape); ``ragged=True`` draws
    view_lens in [min(12,V),V], gmap_lens in [N//3,N], text lengths in [L//2,L].
    ``txt_embeds`` ~ N(0,1) unless produced by ``forward_txt`` by the caller (``txt_ids`` given).
    """
    g = torch.Generator().manual_seed(1000 + seed)
    H = cfg.hidden_size
    if ragged:
        view_lens = _lens(g, B, min(12, V), V)
        gmap_lens = _lens(g, B, max(2, N // 3), N)
        txt_lens = _lens(g, B, max(2, L // 2), L)
    else:
        view_lens = torch.full((B,), V, dtype=torch.long)
        gmap_lens = torch.full((B,), N, dtype=torch.long)
        txt_lens = torch.full((B,), L, dtype=torch.long)
    if txt_law == "r2r":
        txt_lens = r2r_text_lengths(g, B, L)
    ar = torch.arange
    rgb_fts = torch.randn(B, V, cfg.image_feat_size, generator=g)
    dep_fts = torch.randn(B, V, cfg.depth_feat_size, generator=g)
    theta = torch.rand(B, V, generator=g) * (2 * math.pi)
    loc_fts = torch.stack([theta.sin(), theta.cos(), torch.zeros_like(theta), torch.ones_like(theta)], -1)
    nav_types = (torch.rand(B, V, generator=g) < 0.3).long()
    view_mask = ar(V)[None] < view_lens[:, None]
    rgb_fts = rgb_fts * view_mask[..., None]
    dep_fts = dep_fts * view_mask[..., None]
    loc_fts = loc_fts * view_mask[..., None]
    nav_types = nav_types * view_mask

    txt_ids = torch.randint(1000, cfg.vocab_size, (B, L), generator=g)
    txt_ids[:, 0] = 101
    txt_masks = ar(L)[None] < txt_lens[:, None]
    txt_ids = torch.where(txt_masks, txt_ids, torch.full_like(txt_ids, pad_id))
    txt_embeds = torch.randn(B, L, H, generator=g)

    gmap_masks = ar(N)[None] < gmap_lens[:, None]
    n_vis = (0.3 * gmap_lens.float()).floor().long()
    idx = ar(N)[None].expand(B, N)
    gmap_visited_masks = (idx >= 1) & (idx <= n_vis[:, None])
    gmap_step_ids = torch.where(gmap_visited_masks, idx, torch.zeros_like(idx))
    gmap_step_ids = gmap_step_ids.clamp_(max=cfg.max_action_steps - 1)
    gmap_img_fts = torch.randn(B, N, H, generator=g) * gmap_masks[..., None]
    gmap_img_fts[:, 0] = 0  # [stop] node (ss_trainer_ETP.py:364-366)
    a1 = torch.rand(B, N, generator=g) * (2 * math.pi)
    a2 = torch.rand(B, N, generator=g) * (2 * math.pi)
    d1 = torch.rand(B, N, generator=g)
    d2 = torch.rand(B, N, generator=g)
    k = torch.randint(0, 10, (B, N), generator=g).float() / 10.0
    gmap_pos_fts = torch.stack([a1.sin(), a1.cos(), a2.sin(), a2.cos(), d1, d2, k], -1)
    gmap_pos_fts = gmap_pos_fts * gmap_masks[..., None]
    gmap_pos_fts[:, 0] = 0
    pd = torch.rand(B, N, N, generator=g)
    pd = 0.5 * (pd + pd.transpose(1, 2))
    pd = pd * gmap_masks[:, :, None] * gmap_masks[:, None, :]
    pd[:, 0, :] = 0
    pd[:, :, 0] = 0
    pd = pd * (1 - torch.eye(N))[None]
    # teacher labels ~ U over unvisited valid nodes (ss_trainer_ETP.py:890-892)
    ok = gmap_masks & ~gmap_visited_masks
    score = torch.rand(B, N, generator=g).masked_fill(~ok, -1.0)
    labels = score.argmax(1)
    return dict(
        rgb_fts=rgb_fts.contiguous(), dep_fts=dep_fts.contiguous(), loc_fts=loc_fts.contiguous(),
        nav_types=nav_types.contiguous(), view_lens=view_lens,
        txt_ids=txt_ids, txt_masks=txt_masks, txt_embeds=txt_embeds,
        gmap_vpids=None, gmap_step_ids=gmap_step_ids.contiguous(),
        gmap_img_fts=gmap_img_fts.contiguous(), gmap_pos_fts=gmap_pos_fts.contiguous(),
        gmap_masks=gmap_masks, gmap_visited_masks=gmap_visited_masks,
        gmap_pair_dists=pd.contiguous(), labels=labels,
    )


# FLOP model of SURVEY.md §8d (reference formulation, forward; multiply-add = 2).
def step_flops(cfg: PlannerConfig, B: int, V: int, N: int, L: int) -> dict:
    H, I, X = cfg.hidden_size, cfg.intermediate_size, cfg.num_x_layers
    fin = cfg.image_feat_size + (cfg.depth_feat_size if cfg.use_depth_embedding else 0) + cfg.angle_feat_size
    pano = 2 * V * fin * H + cfg.num_pano_layers * (8 * V * H * H + 4 * V * V * H + 4 * V * H * I)
    xlayer = 28 * N * H * H + 4 * L * H * H + 4 * N * L * H + 4 * N * N * H
    nav = X * xlayer + 2 * N * 7 * H + 2 * N * H * H + 2 * N * H
    txt = cfg.num_l_layers * (8 * L * H * H + 4 * L * L * H + 4 * L * H * I)
    return dict(pano=B * pano, nav=B * nav, step_fwd=B * (pano + nav), txt=B * txt,
                txt_kv=B * X * 4 * L * H * H)


def make_traj_batch(cfg: PlannerConfig, B: int, T: int, V: int, L: int, seed: int = 0, mlm_prob: float = 0.15,
                    ghosts: int = 3) -> dict:
    """One PRE-TRAINING batch (CPU tensors / Python lists) with the keys ``mlm_collate`` / ``sap_collate`` produce
    (pretrain_src/pretrain_src/data/tasks.py:98-138): a text, a trajectory of 1..T visited viewpoints per episode with up
    to V views each (the first ones are the navigable candidates, as ``_aggregate_gmap_features`` assumes, vilmodel.py:600)
    and the topological map built from it.  Viewpoint ids are strings; some episodes revisit their first viewpoint (the
    reference keeps the LAST visit's feature), unvisited candidates are shared between steps (their features average),
    and a candidate may be visited later (then it is a visited node)."""
    g = torch.Generator().manual_seed(5000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    steps = [ri(1, T) for _ in range(B)]
    steps[0] = T
    traj_vpids, traj_cand_vpids, gmap_vpids, view_lens = [], [], [], []
    for i in range(B):
        path = [f"v{i}_{t}" for t in range(steps[i])]
        if steps[i] >= 3 and i % 2 == 0:
            path[2] = path[0]                      # revisit
        ghost_ids = [f"u{i}_{k}" for k in range(ghosts)]
        cands, visited_order, unvisited_order = [], [], []
        for t in range(steps[i]):
            vl = V if (i == 0 and t == 0) else ri(max(4, V // 2), V)
            view_lens.append(vl)
            c = []
            if t + 1 < steps[i]:
                c.append(path[t + 1])              # the next viewpoint is a candidate now, visited later
            if t > 0:
                c.append(path[t - 1])              # already visited: skipped by the aggregation
            for k in range(ri(1, max(2, ghosts // 2))):
                c.append(ghost_ids[ri(0, ghosts - 1)])
            c = list(dict.fromkeys(c))[:vl]
            cands.append(c)
            if path[t] not in visited_order:
                visited_order.append(path[t])
        for c in cands:
            for vp in c:
                if vp not in visited_order and vp not in unvisited_order:
                    unvisited_order.append(vp)
        traj_vpids.append(path)
        traj_cand_vpids.append(cands)
        gmap_vpids.append([None] + visited_order + unvisited_order)
    S = sum(steps)
    vlens = torch.tensor(view_lens, dtype=torch.long)
    vm = (torch.arange(V)[None] < vlens[:, None])
    img = torch.randn(S, V, cfg.image_feat_size, generator=g) * vm[..., None]
    dep = torch.randn(S, V, cfg.depth_feat_size, generator=g) * vm[..., None]
    th = torch.rand(S, V, generator=g) * (2 * math.pi)
    loc = torch.stack([th.sin(), th.cos(), torch.zeros_like(th), torch.ones_like(th)], -1) * vm[..., None]
    nav_types = torch.zeros(S, V, dtype=torch.long)
    k = 0
    for i in range(B):
        for t in range(steps[i]):
            nav_types[k, :len(traj_cand_vpids[i][t])] = 1
            k += 1
    gmap_lens = torch.tensor([len(x) for x in gmap_vpids], dtype=torch.long)
    N = int(gmap_lens.max())
    gm = torch.arange(N)[None] < gmap_lens[:, None]
    step_ids = torch.zeros(B, N, dtype=torch.long)
    visited = torch.zeros(B, N, dtype=torch.bool)
    labels = torch.zeros(B, dtype=torch.long)
    for i in range(B):
        nv = len(dict.fromkeys(traj_vpids[i]))
        for n in range(1, 1 + nv):
            step_ids[i, n] = max(t for t in range(steps[i]) if traj_vpids[i][t] == gmap_vpids[i][n]) + 1
        visited[i, 1:1 + nv] = True
        glen = int(gmap_lens[i])
        labels[i] = ri(1 + nv, glen - 1) if (glen > 1 + nv and i % 3 != 2) else 0
    a1 = torch.rand(B, N, generator=g) * (2 * math.pi)
    a2 = torch.rand(B, N, generator=g) * (2 * math.pi)
    pos = torch.stack([a1.sin(), a1.cos(), a2.sin(), a2.cos(), torch.rand(B, N, generator=g), torch.rand(B, N, generator=g),
                       torch.randint(0, 10, (B, N), generator=g).float() / 10.0], -1) * gm[..., None]
    pos[:, 0] = 0
    pd = torch.rand(B, N, N, generator=g)
    pd = 0.5 * (pd + pd.transpose(1, 2)) * gm[:, :, None] * gm[:, None, :] * (1 - torch.eye(N))[None]
    pd[:, 0, :] = 0
    pd[:, :, 0] = 0
    txt_lens = _lens(g, B, max(4, L // 2), L)
    tm = torch.arange(L)[None] < txt_lens[:, None]
    ids = torch.randint(1000, cfg.vocab_size, (B, L), generator=g)
    ids[:, 0] = 101
    sel = (torch.rand(B, L, generator=g) < mlm_prob) & tm
    sel[:, 0] = False
    sel[0, 1] = True                                # at least one masked token
    txt_labels = torch.where(sel, ids, torch.full_like(ids, -1))
    ids = torch.where(sel, torch.full_like(ids, 103), ids)
    ids = torch.where(tm, ids, torch.zeros_like(ids))
    return dict(
        txt_ids=ids, txt_lens=txt_lens, txt_labels=txt_labels,
        traj_view_img_fts=img.contiguous(), traj_view_dep_fts=dep.contiguous(), traj_obj_img_fts=None,
        traj_loc_fts=loc.contiguous(), traj_nav_types=nav_types, traj_step_lens=steps, traj_vp_view_lens=vlens,
        traj_vp_obj_lens=None, traj_vpids=traj_vpids, traj_cand_vpids=traj_cand_vpids,
        gmap_lens=gmap_lens, gmap_step_ids=step_ids, gmap_pos_fts=pos.contiguous(), gmap_pair_dists=pd.contiguous(),
        gmap_vpids=gmap_vpids, gmap_visited_masks=visited, global_act_labels=labels, local_act_labels=None,
    )
