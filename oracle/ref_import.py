"""Import the UNMODIFIED reference planner from /root/reference (authoring container only).

TEST INFRASTRUCTURE.  Recipe from SURVEY.md Appendix C: (1) bypass
``vlnce_baselines/__init__.py`` (imports Habitat) by pre-registering namespace packages,
(2) patch the transformers-4.x ``init_weights`` idiom used at ``vilmodel_cmt.py:673`` for
transformers 5.x.  No reference file is modified or copied.
"""
import os
import sys
import types

REF = os.environ.get("ETPNAV_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "vlnce_baselines/models/etp/vilmodel_cmt.py"))


def load_vilmodel():
    for name, sub in [("vlnce_baselines", ""), ("vlnce_baselines.common", "/common"),
                      ("vlnce_baselines.models", "/models"), ("vlnce_baselines.models.etp", "/models/etp")]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [REF + "/vlnce_baselines" + sub]
            sys.modules[name] = m
    sys.dont_write_bytecode = True  # /root/reference is read-only
    from transformers import PreTrainedModel
    from vlnce_baselines.models.etp import vilmodel_cmt as V
    _orig = PreTrainedModel.init_weights

    def _init_weights_compat(self):
        return self.post_init() if not hasattr(self, "all_tied_weights_keys") else _orig(self)

    V.GlocalTextPathNavCMT.init_weights = _init_weights_compat
    return V


def build_reference(cfg, state_dict=None):
    """Instantiate the reference ``GlocalTextPathNavCMT`` (vilmodel_cmt.py:663) for a PlannerConfig."""
    from transformers import PretrainedConfig
    V = load_vilmodel()
    name = "bert-base-uncased" if cfg.layer_norm_eps == 1e-12 else "xlm-roberta-base"
    hf = PretrainedConfig.from_pretrained(REF + "/bert_config/" + name)
    hf.type_vocab_size = cfg.type_vocab_size
    hf.layer_norm_eps = cfg.layer_norm_eps
    hf.vocab_size = cfg.vocab_size                      # tests shrink the vocab to keep fixtures small
    hf.max_position_embeddings = cfg.max_position_embeddings
    hf.hidden_dropout_prob = cfg.hidden_dropout_prob
    hf.attention_probs_dropout_prob = cfg.attention_probs_dropout_prob
    for k, v in dict(max_action_steps=cfg.max_action_steps, image_feat_size=cfg.image_feat_size,
                     use_depth_embedding=cfg.use_depth_embedding, depth_feat_size=cfg.depth_feat_size,
                     angle_feat_size=cfg.angle_feat_size, num_l_layers=cfg.num_l_layers,
                     num_pano_layers=cfg.num_pano_layers, num_x_layers=cfg.num_x_layers,
                     graph_sprels=cfg.graph_sprels, glocal_fuse="global",
                     fix_lang_embedding=cfg.fix_lang_embedding, fix_pano_embedding=cfg.fix_pano_embedding,
                     update_lang_bert=cfg.update_lang_bert, output_attentions=True,
                     pred_head_dropout_prob=cfg.pred_head_dropout_prob, use_lang2visn_attn=False).items():
        setattr(hf, k, v)
    model = V.GlocalTextPathNavCMT(hf)
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
        assert all("position_ids" in m for m in missing), missing
    return model


# ----------------------------------------------------------------------------------------------------
# pre-training twin (pretrain_src/pretrain_src/model): SURVEY.md §8c, second half of the import recipe
# ----------------------------------------------------------------------------------------------------
def load_pretrain():
    """Import ``model.vilmodel`` / ``model.pretrain_cmt`` of the reference's pre-training tree with two class-attribute
    shims for transformers 5.x: the ``init_weights`` idiom (as above) and ``tie_weights`` (pretrain_cmt.py:79-82 calls
    the removed ``_tie_or_clone_weights``; the shim shares the Parameter, which is what that helper did)."""
    root = REF + "/pretrain_src/pretrain_src"
    if root not in sys.path:
        sys.path.insert(0, root)
    sys.dont_write_bytecode = True
    from transformers import PreTrainedModel
    from model import vilmodel, pretrain_cmt
    _orig = PreTrainedModel.init_weights

    def _init_weights_compat(self):
        return self.post_init() if not hasattr(self, "all_tied_weights_keys") else _orig(self)

    def _tie(self, *a, **k):
        if "mlm" in self.config.pretrain_tasks:
            self.mlm_head.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight

    vilmodel.GlocalTextPathCMT.init_weights = _init_weights_compat
    pretrain_cmt.GlocalTextPathCMTPreTraining.init_weights = _init_weights_compat
    pretrain_cmt.GlocalTextPathCMTPreTraining.tie_weights = _tie
    return vilmodel, pretrain_cmt


def ref_key(k):
    """planner-namespace key -> key of ``GlocalTextPathCMTPreTraining.state_dict()``."""
    return k if k.startswith(("mlm_head.", "global_sap_head.")) else "bert." + k


def build_pretrain_reference(cfg, state_dict=None):
    """Instantiate the reference ``GlocalTextPathCMTPreTraining`` (pretrain_cmt.py:50) for a PlannerConfig with
    ``use_lang2visn_attn`` / ``mlm_head`` set, tasks mlm + sap (run_pt/r2r_pretrain_habitat.json)."""
    from transformers import PretrainedConfig
    _, pretrain_cmt = load_pretrain()
    hf = PretrainedConfig.from_json_file(REF + "/pretrain_src/run_pt/r2r_model_config_dep.json")
    hf.pretrain_tasks = ["mlm", "sap"]
    for k in ("vocab_size", "max_position_embeddings", "type_vocab_size", "layer_norm_eps", "hidden_dropout_prob",
              "attention_probs_dropout_prob", "pred_head_dropout_prob", "max_action_steps", "image_feat_size",
              "depth_feat_size", "angle_feat_size", "num_l_layers", "num_pano_layers", "num_x_layers", "graph_sprels",
              "update_lang_bert", "use_lang2visn_attn"):
        setattr(hf, k, getattr(cfg, k))
    if not cfg.use_depth_embedding:
        hf.depth_feat_size = 0
    model = pretrain_cmt.GlocalTextPathCMTPreTraining(hf)
    if state_dict is not None:
        sd = {ref_key(k): v for k, v in state_dict.items()}
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all("position_ids" in m or m == "mlm_head.predictions.decoder.weight" for m in missing), missing
        model.tie_weights()
    return model


# ----------------------------------------------------------------------------------------------------
# caller-side packing: graph_utils.GraphMap and two ETPTrainer methods
# ----------------------------------------------------------------------------------------------------
def load_graph_utils():
    """Import ``vlnce_baselines/models/graph_utils.py`` unmodified.  Its module-level imports of matplotlib (unused) and
    habitat-lab 0.1.7 (three small geometry helpers used by ``heading_from_quaternion``) are absent here and are
    provided as stub modules restating the published helpers (oracle/packing_port.py)."""
    from . import packing_port as PK
    import numpy as np
    load_vilmodel()  # registers the namespace packages

    def mod(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(sys.modules[name], k, v)

    class _Quat:  # the slice of numpy-quaternion the reference touches: inverse() and being passed back to the helpers
        def __init__(self, q):
            self.q = np.asarray(q, dtype=np.float64)  # (w, x, y, z)

        def inverse(self):
            return _Quat(PK.quat_inverse(self.q))

    mod("matplotlib")
    mod("matplotlib.pyplot")
    mod("habitat")
    mod("habitat.tasks")
    mod("habitat.utils")
    mod("habitat.tasks.utils", cartesian_to_polar=lambda x, y: (np.sqrt(x ** 2 + y ** 2), np.arctan2(y, x)))
    mod("habitat.utils.geometry_utils",
        quaternion_from_coeff=lambda c: _Quat([c[3], c[0], c[1], c[2]]),
        quaternion_rotate_vector=lambda quat, v: PK.quaternion_rotate_vector(quat.q, v))
    from vlnce_baselines.models import graph_utils
    return graph_utils


def load_trainer_packers():
    """The UNMODIFIED bodies of ``ETPTrainer._vp_feature_variable`` / ``_nav_gmap_variable``
    (vlnce_baselines/ss_trainer_ETP.py:308-417), compiled from the reference source (the module itself imports Habitat
    and cannot be imported).  Returned as plain functions taking ``self`` first."""
    import ast
    import numpy as np
    import torch
    from torch.nn.utils.rnn import pad_sequence
    gu = load_graph_utils()
    from vlnce_baselines.common.ops import gen_seq_masks, pad_tensors_wgrad
    path = REF + "/vlnce_baselines/ss_trainer_ETP.py"
    tree = ast.parse(open(path).read())
    ns = dict(torch=torch, np=np, pad_sequence=pad_sequence, pad_tensors_wgrad=pad_tensors_wgrad,
              gen_seq_masks=gen_seq_masks, MAX_DIST=gu.MAX_DIST)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("_vp_feature_variable", "_nav_gmap_variable"):
            m = ast.Module(body=[node], type_ignores=[])
            exec(compile(m, path, "exec"), ns)
            out[node.name] = ns[node.name]
    assert len(out) == 2
    return out
