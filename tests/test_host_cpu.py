"""CPU: host-side logic — state_dict parity with the reference's key set, flat layout, checkpoint key
remapping, and that the C-ABI library loads and exports every symbol include/etpnav_b200.h declares."""
import ctypes
import os
import re

import pytest
import torch

from etpnav_b200.config import PlannerConfig
from etpnav_b200.layout import FlatLayout
from etpnav_b200.spec import param_shapes
from etpnav_b200.synth import make_weights, step_flops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_param_table_matches_survey_counts():
    cfg = PlannerConfig()
    shapes = param_shapes(cfg)
    assert len(shapes) == 307
    n = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    assert abs(n - 140.79e6) < 0.01e6


def test_state_dict_keys_and_flat_storage():
    from etpnav_b200.planner import B200Planner
    cfg = PlannerConfig(vocab_size=512, num_l_layers=2)
    m = B200Planner(cfg, device="cpu")
    sd = m.state_dict()
    assert list(sorted(sd.keys())) == sorted(param_shapes(cfg).keys())
    for k, shape in param_shapes(cfg).items():
        assert tuple(sd[k].shape) == tuple(shape), k
    w = make_weights(cfg, seed=3)
    m.load_state_dict(w, strict=True)
    lay = FlatLayout(cfg)
    for k, (off, numel, shape) in lay.entries.items():
        assert torch.equal(m._flat[off:off + numel].view(shape), w[k]), k  # parameters are views of the flat buffer
    # fused operands are contiguous slices
    p = "global_encoder.encoder.x_layers.0."
    assert lay.offset(p + "visual_attention.att.value.weight") == lay.offset(p + "visual_attention.att.key.weight") + 768 * 768
    assert lay.offset(p + "visn_self_att.self.key.weight") == lay.offset(p + "visn_self_att.self.query.weight") + 768 * 768
    assert all(off % 64 == 0 for off, _, _ in lay.entries.values())


def test_freezing_flags():
    from etpnav_b200.planner import B200Planner
    m = B200Planner(PlannerConfig(vocab_size=512, num_l_layers=1, fix_lang_embedding=True, fix_pano_embedding=True), device="cpu")
    for n, p in m.named_parameters():
        frozen = n.startswith(("embeddings.", "lang_encoder.", "img_embeddings."))
        assert p.requires_grad == (not frozen), n


def test_checkpoint_key_remap():
    from etpnav_b200.planner import remap_checkpoint_keys
    own = ["global_sap_head.net.0.weight", "embeddings.LayerNorm.weight"]
    ck = {"module.bert.global_sap_head.net.0.weight": 1, "net.module.vln_bert.embeddings.LayerNorm.weight": 2,
          "mlm_head.x": 3}
    out = remap_checkpoint_keys(ck, own)
    assert out == {"global_sap_head.net.0.weight": 1, "embeddings.LayerNorm.weight": 2}


def test_library_exports_every_declared_symbol():
    from etpnav_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "etpnav_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(etp_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 15
    L = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/etpnav_b200.h but not exported"
    assert L.etp_version() >= 100


def test_product_path_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from etpnav_b200 import lib
    from etpnav_b200.planner import B200Planner
    m = B200Planner(PlannerConfig(vocab_size=512, num_l_layers=1), device="cpu")
    with pytest.raises(lib.EtpError):
        m.forward_txt(torch.zeros(1, 4, dtype=torch.long), torch.ones(1, 4, dtype=torch.bool))


def test_flop_model_matches_survey():
    f = step_flops(PlannerConfig(), 64, 12, 80, 200)
    assert abs(f["step_fwd"] / 1e9 - 505.3) < 1.0
    assert abs(f["txt"] / 1e9 - 1701.5) < 2.0
    f1 = step_flops(PlannerConfig(), 2, 12, 16, 80)
    assert abs(f1["step_fwd"] / 1e9 - 4.40) < 0.05
