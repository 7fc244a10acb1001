"""CPU restatement of the reference planner hot path — TEST INFRASTRUCTURE, not product code.

Plain tensor algebra on the raw ``state_dict`` (no nn.Module, no reference code), written from
the numerical contract in SURVEY.md Appendix A.  Pinned against the reference modules imported
from /root/reference by ``oracle/make_golden.py`` / ``tests/test_oracle.py`` (the reference has
no golden vectors of its own, so the pin is "reference run here").  Works in fp32 or fp64 and is
autograd-transparent, so the same code provides the backward oracle.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
import math

import torch
import torch.nn.functional as F


def _lin(sd, name, x):
    return x @ sd[name + ".weight"].t() + sd[name + ".bias"]


def _ln(sd, name, x, eps):
    # torch.nn.LayerNorm / BertLayerNorm (vilmodel_cmt.py:24-28): biased variance over last dim
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * sd[name + ".weight"] + sd[name + ".bias"]


def gelu_erf(x):
    # vilmodel_cmt.py:31-37 ; common/transformer.py:469-474 (F.gelu) — both the exact erf form
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


# ---- dropout hooks (train() mode).  ``drop`` is None (eval: identity) or a callable drop(x, kind, site) returning
# x * mask / (1 - p) for the nn.Dropout instance ``site``; kind in {"hidden", "attn", "head"} picks the probability.
# Site ids mirror etpnav_b200/csrc/planner.h (base + 16 * layer + k) so a test can feed the kernels' own masks.
SITE_NAV, SITE_PANO, SITE_TXT, SITE_EMBED, SITE_HEAD = 1000, 2000, 3000, 900, 901
K_XATTN, K_XOUT, K_SATTN, K_SOUT, K_FFNOUT = 0, 1, 2, 3, 4
K_PATTN, K_POUT, K_PFFN, K_PFFNOUT = 0, 1, 2, 3


def _site(base, layer, k):
    return base + 16 * layer + k


def _drop(drop, x, kind, site):
    return x if drop is None else drop(x, kind, site)


def _heads(x, h):
    B, S, H = x.shape
    return x.view(B, S, h, H // h).permute(0, 2, 1, 3)  # vilmodel_cmt.py:98-101


def _merge(x):
    B, h, S, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, S, h * d)  # vilmodel_cmt.py:135-137


def _bert_ctx(sd, pfx, q_src, kv_src, bias, h, drop=None, site=0):
    """BertSelfAttention.forward (vilmodel_cmt.py:103-141) / BertOutAttention.forward (:325-352):
    scale applied AFTER QK^T, additive mask, softmax, dropout on the probabilities (:133, :349), P.V."""
    q = _heads(_lin(sd, pfx + "query", q_src), h)
    k = _heads(_lin(sd, pfx + "key", kv_src), h)
    v = _heads(_lin(sd, pfx + "value", kv_src), h)
    s = q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1])
    s = s + bias
    p = _drop(drop, torch.softmax(s, dim=-1), "attn", site)
    return _merge(p @ v)


def gen_seq_masks(seq_lens, max_len=None):
    # common/ops.py:36-44
    if max_len is None:
        max_len = int(seq_lens.max())
    return torch.arange(max_len, device=seq_lens.device)[None] < seq_lens[:, None]


def extend_neg_masks(masks, dtype):
    # common/ops.py:25-34
    return (1.0 - masks[:, None, None, :].to(dtype)) * -10000.0


def forward_txt(sd, cfg, txt_ids, txt_masks, drop=None):
    """GlocalTextPathNavCMT.forward_txt (vilmodel_cmt.py:684-688): BertEmbeddings.forward (:62-77),
    LanguageEncoder.forward (:426-433), BertLayer.forward (:202-208)."""
    dt = sd["embeddings.LayerNorm.weight"].dtype
    B, L = txt_ids.shape
    e = (sd["embeddings.word_embeddings.weight"][txt_ids]
         + sd["embeddings.position_embeddings.weight"][:L][None]
         + sd["embeddings.token_type_embeddings.weight"][0])
    e = _drop(drop, _ln(sd, "embeddings.LayerNorm", e, cfg.layer_norm_eps), "hidden", SITE_TXT + SITE_EMBED)  # :76
    tb = extend_neg_masks(txt_masks, dt)
    h = cfg.num_attention_heads
    for i in range(cfg.num_l_layers):
        p = f"lang_encoder.layer.{i}."
        ctx = _bert_ctx(sd, p + "attention.self.", e, e, tb, h, drop, _site(SITE_TXT, i, K_SATTN))
        o = _drop(drop, _lin(sd, p + "attention.output.dense", ctx), "hidden", _site(SITE_TXT, i, K_SOUT))  # :153
        a = _ln(sd, p + "attention.output.LayerNorm", o + e, cfg.layer_norm_eps)
        f = _lin(sd, p + "output.dense", gelu_erf(_lin(sd, p + "intermediate.dense", a)))
        f = _drop(drop, f, "hidden", _site(SITE_TXT, i, K_FFNOUT))  # :192
        e = _ln(sd, p + "output.LayerNorm", f + a, cfg.layer_norm_eps)
    if not cfg.update_lang_bert:
        e = e.detach()  # vilmodel_cmt.py:431-432
    return e


def forward_panorama(sd, cfg, rgb_fts, dep_fts, loc_fts, nav_types, view_lens, drop=None):
    """GlocalTextPathNavCMT.forward_panorama (vilmodel_cmt.py:690-719) with the pano encoder of
    common/ops.py:11-23 -> common/transformer.py:71-89 (TransformerEncoder.forward) and :170-182
    (TransformerEncoderLayer.forward_pre, nn.MultiheadAttention: q pre-scaled, key padding -> -inf)."""
    H, h = cfg.hidden_size, cfg.num_attention_heads
    x = _ln(sd, "img_embeddings.img_layer_norm", _lin(sd, "img_embeddings.img_linear", rgb_fts), 1e-12)
    if cfg.use_depth_embedding:
        x = x + _ln(sd, "img_embeddings.dep_layer_norm", _lin(sd, "img_embeddings.dep_linear", dep_fts), 1e-12)
    x = (x + _ln(sd, "img_embeddings.loc_layer_norm", _lin(sd, "img_embeddings.loc_linear", loc_fts), 1e-12)
         + sd["img_embeddings.nav_type_embedding.weight"][nav_types]
         + sd["embeddings.token_type_embeddings.weight"][1])
    x = _drop(drop, _ln(sd, "img_embeddings.layer_norm", x, 1e-12), "hidden", SITE_PANO + SITE_EMBED)  # :710-711
    V = rgb_fts.shape[1]
    m = gen_seq_masks(view_lens, V)
    kbias = torch.zeros(m.shape, dtype=x.dtype, device=x.device).masked_fill(~m, float("-inf"))[:, None, None, :]
    for i in range(cfg.num_pano_layers):
        p = f"img_embeddings.pano_encoder.layers.{i}."
        y = _ln(sd, p + "norm1", x, cfg.pano_layer_norm_eps)
        qkv = y @ sd[p + "self_attn.in_proj_weight"].t() + sd[p + "self_attn.in_proj_bias"]
        q, k, v = qkv.split(H, dim=-1)
        q = _heads(q, h) / math.sqrt(H // h)
        s = q @ _heads(k, h).transpose(-1, -2) + kbias
        pr = _drop(drop, torch.softmax(s, -1), "hidden", _site(SITE_PANO, i, K_PATTN))  # MHA dropout = hidden_dropout_prob (ops.py:15)
        ctx = _merge(pr @ _heads(v, h))
        x = x + _drop(drop, _lin(sd, p + "self_attn.out_proj", ctx), "hidden", _site(SITE_PANO, i, K_POUT))  # dropout1
        y = _ln(sd, p + "norm2", x, cfg.pano_layer_norm_eps)
        hid = _drop(drop, gelu_erf(_lin(sd, p + "linear1", y)), "hidden", _site(SITE_PANO, i, K_PFFN))  # transformer.py:180
        x = x + _drop(drop, _lin(sd, p + "linear2", hid), "hidden", _site(SITE_PANO, i, K_PFFNOUT))  # dropout2
    if cfg.num_pano_layers > 0:
        x = _ln(sd, "img_embeddings.pano_encoder.norm", x, 1e-12)
    return x, m


def forward_navigation(sd, cfg, txt_embeds, txt_masks, gmap_vpids, gmap_step_ids, gmap_img_fts,
                       gmap_pos_fts, gmap_masks, gmap_visited_masks, gmap_pair_dists, drop=None):
    """GlocalTextPathNavCMT.forward_navigation (vilmodel_cmt.py:721-750): node packing (:728-730),
    sprel bias (:732-734), CrossmodalEncoder.forward (:443-452), GraphLXRTXLayer.forward (:383-398),
    NextActionPrediction (:651-661) and the two masked_fill_ (:743-744)."""
    h, eps = cfg.num_attention_heads, cfg.layer_norm_eps
    dt = gmap_img_fts.dtype
    x = (gmap_img_fts + sd["global_encoder.gmap_step_embeddings.weight"][gmap_step_ids]
         + _ln(sd, "global_encoder.gmap_pos_embeddings.1",
               _lin(sd, "global_encoder.gmap_pos_embeddings.0", gmap_pos_fts), 1e-12))
    tbias = extend_neg_masks(txt_masks, dt)
    nbias = extend_neg_masks(gmap_masks, dt)
    if cfg.graph_sprels:
        w = sd["global_encoder.sprel_linear.weight"].reshape(())
        b0 = sd["global_encoder.sprel_linear.bias"].reshape(())
        nbias = nbias + (gmap_pair_dists * w + b0)[:, None]
    for i in range(cfg.num_x_layers):
        p = f"global_encoder.encoder.x_layers.{i}."
        ctx = _bert_ctx(sd, p + "visual_attention.att.", x, txt_embeds, tbias, h, drop, _site(SITE_NAV, i, K_XATTN))
        o = _drop(drop, _lin(sd, p + "visual_attention.output.dense", ctx), "hidden", _site(SITE_NAV, i, K_XOUT))
        a = _ln(sd, p + "visual_attention.output.LayerNorm", o + x, eps)
        ctx = _bert_ctx(sd, p + "visn_self_att.self.", a, a, nbias, h, drop, _site(SITE_NAV, i, K_SATTN))
        o = _drop(drop, _lin(sd, p + "visn_self_att.output.dense", ctx), "hidden", _site(SITE_NAV, i, K_SOUT))
        c = _ln(sd, p + "visn_self_att.output.LayerNorm", o + a, eps)
        f = _lin(sd, p + "visn_output.dense", gelu_erf(_lin(sd, p + "visn_inter.dense", c)))
        f = _drop(drop, f, "hidden", _site(SITE_NAV, i, K_FFNOUT))
        x = _ln(sd, p + "visn_output.LayerNorm", f + c, eps)
    hh = _ln(sd, "global_sap_head.net.2", F.relu(_lin(sd, "global_sap_head.net.0", x)), 1e-12)
    hh = _drop(drop, hh, "head", SITE_NAV + SITE_HEAD)  # NextActionPrediction's Dropout (:656)
    logits = _lin(sd, "global_sap_head.net.4", hh)[..., 0]
    logits = logits.masked_fill(gmap_visited_masks, float("-inf"))
    logits = logits.masked_fill(~gmap_masks, float("-inf"))
    return {"gmap_embeds": x, "global_logits": logits}


def step_loss(logits, labels):
    """Caller-side loss of one step: F.cross_entropy(reduction='sum', ignore_index=-100)
    (ss_trainer_ETP.py:890-892)."""
    return F.cross_entropy(logits, labels, reduction="sum", ignore_index=-100)
