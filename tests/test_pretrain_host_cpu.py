"""CPU: host logic of the pre-training twin (etpnav_b200/pretrain.py) — reference key layout incl. the tied decoder
alias, flat-layout adjacency of the lang_* fused operands, ctypes struct construction, loud failure without a GPU."""
import ctypes as C

import pytest
import torch

from etpnav_b200.config import PlannerConfig
from etpnav_b200.layout import FlatLayout
from etpnav_b200.spec import param_shapes
from etpnav_b200.synth import make_traj_batch, make_weights


def _cfg(**kw):
    return PlannerConfig(vocab_size=2048, num_l_layers=1, num_x_layers=2, **kw)


def test_navigation_layout_is_unchanged_by_the_pretrain_group():
    base, pre = FlatLayout(_cfg()), FlatLayout(_cfg(use_lang2visn_attn=True, mlm_head=True))
    for k, e in base.entries.items():
        assert pre.entries[k] == e, k          # the three navigation groups keep their offsets
    assert base.group_ranges["pre"][0] == base.group_ranges["pre"][1] == base.total
    assert pre.group_ranges["pre"][0] == pre.group_ranges["nav"][1]  # adjacent: one gradient slice for lang2visn


def test_pretraining_state_dict_layout_and_roundtrip():
    from etpnav_b200.pretrain import B200PreTraining
    cfg = _cfg()
    m = B200PreTraining(cfg, device="cpu")
    assert cfg.use_lang2visn_attn and cfg.mlm_head
    sd = m.state_dict()
    want = {("bert." + k if not k.startswith(("mlm_head.", "global_sap_head.")) else k) for k in param_shapes(cfg)}
    want.add("mlm_head.predictions.decoder.weight")
    assert set(sd.keys()) == want
    assert sd["mlm_head.predictions.decoder.weight"].data_ptr() == sd["bert.embeddings.word_embeddings.weight"].data_ptr()
    w = make_weights(cfg, seed=9)
    ref_sd = {("bert." + k if not k.startswith(("mlm_head.", "global_sap_head.")) else k): v for k, v in w.items()}
    ref_sd["mlm_head.predictions.decoder.weight"] = ref_sd["bert.embeddings.word_embeddings.weight"]
    ref_sd["bert.embeddings.position_ids"] = torch.arange(4)[None]   # buffer present in old HF checkpoints
    m.load_state_dict(ref_sd, strict=True)
    for k, v in w.items():
        assert torch.equal(m.bert._pmap[k], v), k
    with pytest.raises(KeyError):
        m.load_state_dict({"bert.nonexistent.weight": torch.zeros(1)}, strict=True)


def test_l2v_structs_point_at_the_lang_blocks():
    from etpnav_b200.pretrain import B200TextPathCMT
    cfg = _cfg(use_lang2visn_attn=True, mlm_head=True)
    m = B200TextPathCMT(cfg, device="cpu")
    b32, b16 = 1 << 30, 1 << 32
    s = m._build_structs(b32, b16, 2)
    lay = m.layout
    nav, l2v = s["nav"], s["l2v"]
    assert l2v.num_x_layers == 2 and l2v.xkv_all_w == nav.xkv_all_w and l2v.pos_w == nav.pos_w
    for i in range(2):
        p = f"global_encoder.encoder.x_layers.{i}."
        lw, nw = s["l2v_layers"][i], s["nav_layers"][i]
        assert lw.xq_w == nw.xq_w and lw.xo_w == nw.xo_w and lw.xkv_w == nw.xkv_w     # shared visual_attention
        assert lw.sqkv_w == b16 + 2 * lay.offset(p + "lang_self_att.self.query.weight")
        assert lw.sqkv_b == b32 + 4 * lay.offset(p + "lang_self_att.self.query.bias")
        assert lw.f1_w == b16 + 2 * lay.offset(p + "lang_inter.dense.weight")
        assert lw.fln_g == b32 + 4 * lay.offset(p + "lang_output.LayerNorm.weight")
        assert nw.sqkv_w == b16 + 2 * lay.offset(p + "visn_self_att.self.query.weight")
    names = m._l2v_param_names()
    assert any(".lang_self_att." in n for n in names) and not any(n.startswith("mlm_head") for n in names)
    lo, hi = lay.group_ranges["nav"][0], lay.group_ranges["pre"][1]
    assert all(lo <= lay.offset(n) < hi for n in names)


def test_pretraining_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from etpnav_b200 import lib
    from etpnav_b200.pretrain import B200PreTraining
    cfg = _cfg()
    m = B200PreTraining(cfg, device="cpu")
    b = make_traj_batch(cfg, 2, 2, 6, 10, seed=0)
    with pytest.raises(lib.EtpError):
        m(b, "mlm")
    with pytest.raises(lib.EtpError):
        m(b, "sap")
    with pytest.raises(ValueError):
        m(b, "mrc")


def test_from_pretrained_mirrors_the_reference_driver_call():
    """train_r2r.py:146-148: from_pretrained(None, config=<object with the JSON's attributes>, state_dict=<checkpoint>):
    foreign keys (e.g. a checkpoint that still carries heads this model does not own) are ignored, own keys load."""
    import json
    import types
    from etpnav_b200.pretrain import B200PreTraining
    js = dict(pred_head_dropout_prob=0.1, attention_probs_dropout_prob=0.1, hidden_dropout_prob=0.1, hidden_size=768,
              image_feat_size=512, depth_feat_size=128, angle_feat_size=4, obj_feat_size=0, intermediate_size=3072,
              num_l_layers=1, num_x_layers=2, num_pano_layers=2, layer_norm_eps=1e-12, max_position_embeddings=512,
              max_action_steps=100, num_attention_heads=12, type_vocab_size=2, update_lang_bert=True, vocab_size=2048,
              use_lang2visn_attn=True, graph_sprels=True, glocal_fuse=True)
    hf = types.SimpleNamespace(**json.loads(json.dumps(js)))
    w = make_weights(PlannerConfig(vocab_size=2048, num_l_layers=1, num_x_layers=2, use_lang2visn_attn=True, mlm_head=True), seed=4)
    ck = {("bert." + k if not k.startswith(("mlm_head.", "global_sap_head.")) else k): v for k, v in w.items()}
    ck["image_classifier.net.0.weight"] = torch.zeros(3, 3)        # an 'mrc' head of another run: ignored
    m = B200PreTraining.from_pretrained(pretrained_model_name_or_path=None, config=hf, state_dict=ck, device="cpu")
    assert m.config.num_x_layers == 2 and m.config.use_lang2visn_attn and m.config.mlm_head and m.config.use_depth_embedding
    for k, v in w.items():
        assert torch.equal(m.bert._pmap[k], v), k
    with pytest.raises(ValueError):
        B200PreTraining.from_pretrained("bert-base-uncased", config=hf)


def test_set_dropout_reaches_the_kernel_arguments():
    from etpnav_b200.pretrain import B200PreTraining
    m = B200PreTraining(_cfg(), device="cpu").train()
    m.set_dropout(0.25)
    d = m.bert._next_dropout()
    assert abs(d.p_hidden - 0.25) < 1e-7 and abs(d.p_attn - 0.25) < 1e-7 and abs(d.p_head - 0.25) < 1e-7
    m.set_dropout(0.0)
    assert m.bert._next_dropout() is None       # p = 0: the eval() code path
    m.eval()
    m.set_dropout(0.1)
    assert m.bert._next_dropout() is None
