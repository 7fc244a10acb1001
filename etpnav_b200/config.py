"""Planner configuration.

Mirrors the attributes the reference puts on its HF ``PretrainedConfig`` in
``vlnce_baselines/models/etp/vlnbert_init.py:32-59`` plus the BERT / XLM-R JSON
(``bert_config/bert-base-uncased/config.json``, ``bert_config/xlm-roberta-base/config.json``).
Only what the planner hot path reads is kept.
"""
from dataclasses import dataclass, asdict


@dataclass
class PlannerConfig:
    # bert_config/*/config.json
    hidden_size: int = 768
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    vocab_size: int = 30522
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    # vlnbert_init.py:41-59
    max_action_steps: int = 100
    image_feat_size: int = 512
    use_depth_embedding: bool = True
    depth_feat_size: int = 128
    angle_feat_size: int = 4
    num_l_layers: int = 9
    num_pano_layers: int = 2
    num_x_layers: int = 4
    graph_sprels: bool = True
    fix_lang_embedding: bool = False
    fix_pano_embedding: bool = False
    update_lang_bert: bool = True
    pred_head_dropout_prob: float = 0.1
    # pano encoder layers use nn.LayerNorm's default eps (common/transformer.py:144-145)
    pano_layer_norm_eps: float = 1e-5
    # pre-training twin (pretrain_src/run_pt/r2r_model_config_dep.json): extra lang_self_att / lang_inter / lang_output
    # parameters in every x-layer (pretrain_src/pretrain_src/model/vilmodel.py:370-374) and the MLM head
    # (pretrain_cmt.py:56-57; decoder tied to the word embeddings, :79-82).  Both off for the navigation model.
    use_lang2visn_attn: bool = False
    mlm_head: bool = False

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @classmethod
    def for_task(cls, task_type: str = "r2r", **kw) -> "PlannerConfig":
        """task_type 'r2r' -> BERT vocab / eps 1e-12; 'rxr' -> XLM-R (vlnbert_init.py:32-39)."""
        if task_type == "r2r":
            base = dict(vocab_size=30522, max_position_embeddings=512, layer_norm_eps=1e-12)
        elif task_type == "rxr":
            base = dict(vocab_size=250002, max_position_embeddings=514, layer_norm_eps=1e-5,
                        type_vocab_size=2)
        else:
            raise ValueError(f"unknown task_type {task_type!r}")
        base.update(kw)
        return cls(**base)

    def to_dict(self):
        return asdict(self)
