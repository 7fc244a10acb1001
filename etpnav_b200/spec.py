"""Parameter table of the planner: names and shapes of the reference ``state_dict()``.

Follows SURVEY.md Appendix B (measured from the oracle's ``state_dict()``: 307 tensors at
X=4, BERT vocab). Reference modules: ``vlnce_baselines/models/etp/vilmodel_cmt.py``
``BertEmbeddings`` :48, ``LanguageEncoder`` :413, ``ImageEmbeddings`` :454,
``GlobalMapEncoder`` :566, ``GraphLXRTXLayer`` :365, ``NextActionPrediction`` :651 and
``common/transformer.py`` ``TransformerEncoderLayer`` :127.
"""
from collections import OrderedDict
from .config import PlannerConfig


def _lin(d, name, out_f, in_f):
    d[name + ".weight"] = (out_f, in_f)
    d[name + ".bias"] = (out_f,)


def _ln(d, name, h):
    d[name + ".weight"] = (h,)
    d[name + ".bias"] = (h,)


def param_shapes(cfg: PlannerConfig) -> "OrderedDict[str, tuple]":
    H, I = cfg.hidden_size, cfg.intermediate_size
    d = OrderedDict()
    # BertEmbeddings (vilmodel_cmt.py:48-60)
    d["embeddings.word_embeddings.weight"] = (cfg.vocab_size, H)
    d["embeddings.position_embeddings.weight"] = (cfg.max_position_embeddings, H)
    d["embeddings.token_type_embeddings.weight"] = (cfg.type_vocab_size, H)
    _ln(d, "embeddings.LayerNorm", H)
    # LanguageEncoder: num_l_layers x BertLayer (vilmodel_cmt.py:195-208, 413-424)
    for i in range(cfg.num_l_layers):
        p = f"lang_encoder.layer.{i}."
        for n in ("query", "key", "value"):
            _lin(d, p + "attention.self." + n, H, H)
        _lin(d, p + "attention.output.dense", H, H)
        _ln(d, p + "attention.output.LayerNorm", H)
        _lin(d, p + "intermediate.dense", I, H)
        _lin(d, p + "output.dense", H, I)
        _ln(d, p + "output.LayerNorm", H)
    # ImageEmbeddings (vilmodel_cmt.py:454-486)
    _lin(d, "img_embeddings.img_linear", H, cfg.image_feat_size)
    _ln(d, "img_embeddings.img_layer_norm", H)
    _lin(d, "img_embeddings.loc_linear", H, cfg.angle_feat_size)
    _ln(d, "img_embeddings.loc_layer_norm", H)
    if cfg.use_depth_embedding:
        _lin(d, "img_embeddings.dep_linear", H, cfg.depth_feat_size)
        _ln(d, "img_embeddings.dep_layer_norm", H)
    d["img_embeddings.nav_type_embedding.weight"] = (2, H)
    _ln(d, "img_embeddings.layer_norm", H)
    for i in range(cfg.num_pano_layers):
        p = f"img_embeddings.pano_encoder.layers.{i}."
        d[p + "self_attn.in_proj_weight"] = (3 * H, H)
        d[p + "self_attn.in_proj_bias"] = (3 * H,)
        _lin(d, p + "self_attn.out_proj", H, H)
        _lin(d, p + "linear1", I, H)
        _lin(d, p + "linear2", H, I)
        _ln(d, p + "norm1", H)
        _ln(d, p + "norm2", H)
    if cfg.num_pano_layers > 0:
        _ln(d, "img_embeddings.pano_encoder.norm", H)
    # GlobalMapEncoder (vilmodel_cmt.py:566-579)
    _lin(d, "global_encoder.gmap_pos_embeddings.0", H, cfg.angle_feat_size + 3)
    _ln(d, "global_encoder.gmap_pos_embeddings.1", H)
    d["global_encoder.gmap_step_embeddings.weight"] = (cfg.max_action_steps, H)
    for i in range(cfg.num_x_layers):
        p = f"global_encoder.encoder.x_layers.{i}."
        for n in ("query", "key", "value"):
            _lin(d, p + "visual_attention.att." + n, H, H)
        _lin(d, p + "visual_attention.output.dense", H, H)
        _ln(d, p + "visual_attention.output.LayerNorm", H)
        for n in ("query", "key", "value"):
            _lin(d, p + "visn_self_att.self." + n, H, H)
        _lin(d, p + "visn_self_att.output.dense", H, H)
        _ln(d, p + "visn_self_att.output.LayerNorm", H)
        _lin(d, p + "visn_inter.dense", I, H)
        _lin(d, p + "visn_output.dense", H, I)
        _ln(d, p + "visn_output.LayerNorm", H)
    if cfg.graph_sprels:
        _lin(d, "global_encoder.sprel_linear", 1, 1)
    # NextActionPrediction (vilmodel_cmt.py:651-661)
    _lin(d, "global_sap_head.net.0", H, H)
    _ln(d, "global_sap_head.net.2", H)
    _lin(d, "global_sap_head.net.4", 1, H)
    # pre-training twin only (pretrain_src/pretrain_src/model/vilmodel.py:370-374; BertOnlyMLMHead :258-299)
    if cfg.use_lang2visn_attn:
        for i in range(cfg.num_x_layers):
            p = f"global_encoder.encoder.x_layers.{i}."
            for n in ("query", "key", "value"):
                _lin(d, p + "lang_self_att.self." + n, H, H)
            _lin(d, p + "lang_self_att.output.dense", H, H)
            _ln(d, p + "lang_self_att.output.LayerNorm", H)
            _lin(d, p + "lang_inter.dense", I, H)
            _lin(d, p + "lang_output.dense", H, I)
            _ln(d, p + "lang_output.LayerNorm", H)
    if cfg.mlm_head:
        _lin(d, "mlm_head.predictions.transform.dense", H, H)
        _ln(d, "mlm_head.predictions.transform.LayerNorm", H)
        d["mlm_head.predictions.bias"] = (cfg.vocab_size,)
    return d
