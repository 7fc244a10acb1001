"""Flat parameter layout of the planner.

All parameters of the module live in ONE fp32 buffer (and their gradients in one more, and the bf16
GEMM copies in a third), in the order produced here.  Weights that one kernel consumes as a fused
operand are adjacent, so the fused operand is a contiguous slice and needs no concatenation:
``key|value`` of the cross-attention ([1536,768]) and ``query|key|value`` of the self-attention
([2304,768]) — the reference keeps them as separate ``nn.Linear`` (vilmodel_cmt.py:87-89,310-312) and so
does our ``state_dict()``.

Groups follow the three reference methods so the step-level C calls and the gradient all-reduce can
address them as slices: ``txt`` (embeddings + lang_encoder), ``pano`` (img_embeddings), ``nav``
(global_encoder + global_sap_head).
"""
from collections import OrderedDict

from .config import PlannerConfig
from .spec import param_shapes


def ordered_names(cfg: PlannerConfig):
    """Returns OrderedDict group -> [param names] with fused operands adjacent."""
    g = OrderedDict(txt=[], pano=[], nav=[], pre=[])
    t = g["txt"]
    # token_type_embeddings first: forward_panorama reads row 1 of it (vilmodel_cmt.py:709)
    t += ["embeddings.token_type_embeddings.weight", "embeddings.word_embeddings.weight",
          "embeddings.position_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"]
    for i in range(cfg.num_l_layers):
        p = f"lang_encoder.layer.{i}."
        t += [p + f"attention.self.{n}.weight" for n in ("query", "key", "value")]
        t += [p + f"attention.self.{n}.bias" for n in ("query", "key", "value")]
        t += [p + "attention.output.dense.weight", p + "attention.output.dense.bias",
              p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias",
              p + "intermediate.dense.weight", p + "intermediate.dense.bias",
              p + "output.dense.weight", p + "output.dense.bias",
              p + "output.LayerNorm.weight", p + "output.LayerNorm.bias"]
    pn = g["pano"]
    pn += ["img_embeddings.img_linear.weight", "img_embeddings.img_linear.bias",
           "img_embeddings.img_layer_norm.weight", "img_embeddings.img_layer_norm.bias"]
    if cfg.use_depth_embedding:
        pn += ["img_embeddings.dep_linear.weight", "img_embeddings.dep_linear.bias",
               "img_embeddings.dep_layer_norm.weight", "img_embeddings.dep_layer_norm.bias"]
    pn += ["img_embeddings.loc_linear.weight", "img_embeddings.loc_linear.bias",
           "img_embeddings.loc_layer_norm.weight", "img_embeddings.loc_layer_norm.bias",
           "img_embeddings.nav_type_embedding.weight",
           "img_embeddings.layer_norm.weight", "img_embeddings.layer_norm.bias"]
    for i in range(cfg.num_pano_layers):
        p = f"img_embeddings.pano_encoder.layers.{i}."
        pn += [p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias",
               p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias",
               p + "linear1.weight", p + "linear1.bias", p + "linear2.weight", p + "linear2.bias",
               p + "norm1.weight", p + "norm1.bias", p + "norm2.weight", p + "norm2.bias"]
    if cfg.num_pano_layers > 0:
        pn += ["img_embeddings.pano_encoder.norm.weight", "img_embeddings.pano_encoder.norm.bias"]
    nv = g["nav"]
    nv += ["global_encoder.gmap_pos_embeddings.0.weight", "global_encoder.gmap_pos_embeddings.0.bias",
           "global_encoder.gmap_pos_embeddings.1.weight", "global_encoder.gmap_pos_embeddings.1.bias",
           "global_encoder.gmap_step_embeddings.weight"]
    if cfg.graph_sprels:
        nv += ["global_encoder.sprel_linear.weight", "global_encoder.sprel_linear.bias"]
    # text-side key|value projections of ALL cross-modal layers are adjacent: K/V of every layer come out of ONE
    # [B*L,768] x [X*1536,768]^T GEMM (they depend only on txt_embeds), and their dgrad / wgrad are one GEMM each
    for i in range(cfg.num_x_layers):
        p = f"global_encoder.encoder.x_layers.{i}."
        nv += [p + f"visual_attention.att.{n}.weight" for n in ("key", "value")]
    for i in range(cfg.num_x_layers):
        p = f"global_encoder.encoder.x_layers.{i}."
        nv += [p + f"visual_attention.att.{n}.bias" for n in ("key", "value")]
    for i in range(cfg.num_x_layers):
        p = f"global_encoder.encoder.x_layers.{i}."
        nv += [p + "visual_attention.att.query.weight", p + "visual_attention.att.query.bias"]
        nv += [p + "visual_attention.output.dense.weight", p + "visual_attention.output.dense.bias",
               p + "visual_attention.output.LayerNorm.weight", p + "visual_attention.output.LayerNorm.bias"]
        nv += [p + f"visn_self_att.self.{n}.weight" for n in ("query", "key", "value")]
        nv += [p + f"visn_self_att.self.{n}.bias" for n in ("query", "key", "value")]
        nv += [p + "visn_self_att.output.dense.weight", p + "visn_self_att.output.dense.bias",
               p + "visn_self_att.output.LayerNorm.weight", p + "visn_self_att.output.LayerNorm.bias",
               p + "visn_inter.dense.weight", p + "visn_inter.dense.bias",
               p + "visn_output.dense.weight", p + "visn_output.dense.bias",
               p + "visn_output.LayerNorm.weight", p + "visn_output.LayerNorm.bias"]
    nv += ["global_sap_head.net.0.weight", "global_sap_head.net.0.bias",
           "global_sap_head.net.2.weight", "global_sap_head.net.2.bias",
           "global_sap_head.net.4.weight", "global_sap_head.net.4.bias"]
    # pre-training twin only: appended as a fourth group so the offsets of the three navigation groups never move
    pr = g["pre"]
    if cfg.use_lang2visn_attn:
        for i in range(cfg.num_x_layers):
            p = f"global_encoder.encoder.x_layers.{i}."
            pr += [p + f"lang_self_att.self.{n}.weight" for n in ("query", "key", "value")]
            pr += [p + f"lang_self_att.self.{n}.bias" for n in ("query", "key", "value")]
            pr += [p + "lang_self_att.output.dense.weight", p + "lang_self_att.output.dense.bias",
                   p + "lang_self_att.output.LayerNorm.weight", p + "lang_self_att.output.LayerNorm.bias",
                   p + "lang_inter.dense.weight", p + "lang_inter.dense.bias",
                   p + "lang_output.dense.weight", p + "lang_output.dense.bias",
                   p + "lang_output.LayerNorm.weight", p + "lang_output.LayerNorm.bias"]
    if cfg.mlm_head:
        pr += ["mlm_head.predictions.transform.dense.weight", "mlm_head.predictions.transform.dense.bias",
               "mlm_head.predictions.transform.LayerNorm.weight", "mlm_head.predictions.transform.LayerNorm.bias",
               "mlm_head.predictions.bias"]
    return g


class FlatLayout:
    """name -> (offset, numel, shape) into the flat buffers; every slice starts on a 64-element boundary
    (256 B in fp32, 128 B in bf16) so TMA / 128-bit accesses are always aligned."""

    ALIGN = 64

    def __init__(self, cfg: PlannerConfig):
        shapes = param_shapes(cfg)
        groups = ordered_names(cfg)
        flat = [n for names in groups.values() for n in names]
        assert sorted(flat) == sorted(shapes.keys()), set(flat) ^ set(shapes.keys())
        self.entries = OrderedDict()
        self.group_ranges = OrderedDict()
        off = 0
        for gname, names in groups.items():
            start = off
            for n in names:
                numel = 1
                for d in shapes[n]:
                    numel *= d
                self.entries[n] = (off, numel, tuple(shapes[n]))
                off += numel
                # fused operands must stay contiguous: their sizes are multiples of ALIGN already
                off = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            self.group_ranges[gname] = (start, off)
        self.total = off

    def offset(self, name):
        return self.entries[name][0]
