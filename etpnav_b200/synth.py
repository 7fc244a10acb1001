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


def make_inputs(cfg: PlannerConfig, B: int, V: int, N: int, L: int, seed: int = 0,
                ragged: bool = True, txt_from: str = "normal", pad_id: int = 0) -> dict:
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
