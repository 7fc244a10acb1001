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


def test_checkpoint_interop_finetune_and_pretrain_layouts():
    """N4: the key layouts of the reference's checkpoints load unchanged.
    (1) fine-tune ``ckpt.iter*.pth``: ``{"state_dict": policy.state_dict()}`` with keys ``net.vln_bert.<name>``
        (``net.module.vln_bert.<name>`` when saved from DDP, ss_trainer_ETP.py:74-83,223-236) — loaded by the PARENT
        module's ``load_state_dict``, so only our key set has to equal the reference's;
    (2) pre-train ``model_step_*.pt``: ``bert.<name>`` for the backbone, the SAP head un-prefixed, plus heads we do not
        own (vlnbert_init.py:20-30 prefixes the head, HF strips ``bert.``) — loaded through ``remap_checkpoint_keys``."""
    import torch.nn as nn
    from etpnav_b200.planner import B200Planner, remap_checkpoint_keys
    cfg = PlannerConfig(vocab_size=256, num_l_layers=1, num_x_layers=2)
    src = make_weights(cfg, seed=5)

    class Net(nn.Module):       # stands in for ETP(Net) (Policy_ViewSelection_ETP.py:78-92)
        def __init__(self):
            super().__init__()
            self.vln_bert = B200Planner(cfg, device="cpu")

    class Policy(nn.Module):    # stands in for PolicyViewSelectionETP: policy.net
        def __init__(self):
            super().__init__()
            self.net = Net()

    pol = Policy()
    ckpt = {"state_dict": {"net.vln_bert." + k: v.clone() for k, v in src.items()}, "iteration": 7}
    missing, unexpected = pol.load_state_dict(ckpt["state_dict"], strict=True)
    assert not missing and not unexpected
    for k, v in src.items():
        assert torch.equal(pol.net.vln_bert.state_dict()[k], v), k
    # saving gives back the reference's key layout
    assert sorted(pol.state_dict().keys()) == sorted("net.vln_bert." + k for k in src)

    # DDP-saved variant and the pre-training layout go through the remap
    m = B200Planner(cfg, device="cpu")
    ddp = {"net.module.vln_bert." + k: v for k, v in src.items()}
    out = remap_checkpoint_keys(ddp, m.state_dict().keys())
    assert sorted(out) == sorted(src)
    pre = {}
    for k, v in src.items():
        pre[(k if k.startswith("global_sap_head") else "bert." + k)] = v
    pre["mlm_head.predictions.bias"] = torch.zeros(3)           # heads the planner does not own are ignored
    pre["module.sap_fuse_linear.weight"] = torch.zeros(1, 3)
    ref_style = {}
    for k, v in pre.items():                                     # vlnbert_init.py:24-30, restated
        if k.startswith("module"):
            ref_style[k[7:]] = v
        ref_style[("bert." + k) if "sap_head" in k else k] = v
    out = remap_checkpoint_keys(ref_style, m.state_dict().keys())
    assert sorted(out) == sorted(src)
    res = m.load_state_dict(out, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in src.items():
        assert torch.equal(m.state_dict()[k], v), k


def test_trainer_steps_only_trainable_runs():
    """torch.optim.AdamW skips parameters without a gradient: with fix_pano_embedding (vilmodel_cmt.py:680-682) the
    panorama group is neither updated nor decayed — PlannerTrainer's stepped runs exclude it."""
    from etpnav_b200.planner import B200Planner, PlannerTrainer
    cfg = PlannerConfig(vocab_size=512, num_l_layers=1, num_x_layers=2)
    m = B200Planner(cfg, device="cpu")
    lo = min(m.layout.group_ranges[g][0] for g in ("pano", "nav"))
    hi = max(m.layout.group_ranges[g][1] for g in ("pano", "nav"))
    assert PlannerTrainer._active_ranges(m, lo, hi) == [(lo, hi)]                       # nothing frozen: one run
    mf = B200Planner(PlannerConfig(vocab_size=512, num_l_layers=1, num_x_layers=2, fix_pano_embedding=True), device="cpu")
    assert PlannerTrainer._active_ranges(mf, lo, hi) == [mf.layout.group_ranges["nav"]]
    # one frozen tensor in the middle splits the run around it
    name = "global_encoder.encoder.x_layers.0.visn_inter.dense.weight"
    m._pmap[name].requires_grad = False
    off, numel, _ = m.layout.entries[name]
    runs = PlannerTrainer._active_ranges(m, lo, hi)
    assert len(runs) == 2 and runs[0][1] == off and runs[1][0] >= off + numel and runs[0][0] == lo and runs[1][1] == hi


def test_pretraining_state_dict_survives_wrapping():
    """B200PreTraining keeps the reference's key layout (bert.*, mlm_head.*, global_sap_head.* at top level, tied decoder
    alias) when it is a child module (DDP / DataParallel: train_r2r.py's ModelSaver strips ``module.``) — both ways."""
    import torch.nn as nn
    from etpnav_b200.pretrain import B200PreTraining, PretrainTrainer

    def make():
        return B200PreTraining(PlannerConfig(vocab_size=512, num_l_layers=1, num_x_layers=1, num_pano_layers=1), device="cpu")

    class Wrap(nn.Module):
        def __init__(self, mod):
            super().__init__()
            self.module = mod
    m = make()
    sd = m.state_dict()
    assert "mlm_head.predictions.decoder.weight" in sd and "global_sap_head.net.0.weight" in sd
    assert "bert.embeddings.word_embeddings.weight" in sd and not any(k.startswith("bert.mlm_head") for k in sd)
    wsd = Wrap(m).state_dict()
    assert len(wsd) == len(sd) > 100 and set(wsd) == {"module." + k for k in sd}        # not empty, reference keys
    stripped = {k[7:]: v.clone() + 1.0 for k, v in wsd.items()}                          # what ModelSaver writes
    m2 = make()
    m2.load_state_dict(stripped, strict=True)
    assert all(torch.equal(m2.state_dict()[k], stripped[k]) for k in stripped)
    w3 = Wrap(make())
    w3.load_state_dict(wsd, strict=True)                                                 # parent recursion reaches the hooks
    assert all(torch.equal(w3.state_dict()[k], wsd[k]) for k in wsd)
    # optimizer block flags: biases / LayerNorm carry no weight decay (optim/misc.py:14), frozen tensors are inactive
    m.bert._pmap["embeddings.word_embeddings.weight"].requires_grad = False
    fl = PretrainTrainer.block_flags(m.bert)
    ent = m.bert.layout.entries
    at = lambda n: int(fl[ent[n][0] // 64])
    assert at("embeddings.word_embeddings.weight") == 0
    assert at("lang_encoder.layer.0.attention.self.query.weight") == 3
    assert at("lang_encoder.layer.0.attention.self.query.bias") == 1
    assert at("lang_encoder.layer.0.attention.output.LayerNorm.weight") == 1


def test_ctypes_struct_mirrors_match_the_header():
    """The ctypes mirrors in etpnav_b200/lib.py and planner.py have the sizes the compiled header has (a field added on
    one side only would shift every later pointer)."""
    import ctypes as C
    from etpnav_b200 import lib as L
    from etpnav_b200 import planner as P
    lib = C.CDLL(L.LIB_PATH)
    lib.etp_struct_sizes.argtypes = [C.POINTER(C.c_int32), C.c_int32]
    out = (C.c_int32 * 16)()
    n = lib.etp_struct_sizes(out, 16)
    mirrors = [L.GemmArgs, L.AttnArgs, L.AttnBwdArgs, L.PanoPackArgs, L.NodePackArgs, P.Dropout, P.LayerWeights, P.NavWeights,
               P.NavInputs, P.PanoLayerWeights, P.PanoWeights, P.PanoInputs, P.TxtWeights]
    assert n == len(mirrors)
    for i, m in enumerate(mirrors):
        assert C.sizeof(m) == out[i], (m.__name__, C.sizeof(m), out[i])


def test_peer_partition_covers_every_run_with_aligned_disjoint_sub_slices():
    """Owner sub-slices of the peer-memory update (planner.peer_partition): disjoint, in rank order, 64-element aligned
    starts, covering the trainable run exactly; trailing ranks of a short run own nothing."""
    from etpnav_b200.planner import peer_partition
    for x, y in [(0, 64), (128, 128 + 64 * 3), (4096, 4096 + 64 * 1001), (64, 64 + 7_077_888), (0, 0)]:
        for world in (2, 4, 8):
            parts = [peer_partition(x, y, world, r) for r in range(world)]
            assert parts[0][0] == x and parts[-1][1] == y
            assert all(a <= b and (a - x) % 64 == 0 for a, b in parts)
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert sum(b - a for a, b in parts) == y - x


def test_peer_entry_points_validate_their_arguments_without_a_device():
    """etp_peer_* / etp_ipc_* (include/etpnav_b200.h) reject malformed groups and ranges before anything is launched."""
    import ctypes as C
    from etpnav_b200 import lib as L
    from etpnav_b200.planner import PeerGroup
    lib = L.lib()
    lib.etp_last_error.restype = C.c_char_p
    lib.etp_peer_signal.argtypes = [C.POINTER(PeerGroup), C.c_int32, C.c_int32, C.c_uint32, C.c_void_p]
    lib.etp_peer_wait.argtypes = [C.POINTER(PeerGroup), C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_double, C.c_void_p]
    lib.etp_peer_reduce_adamw.argtypes = [C.POINTER(PeerGroup), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p] + [C.c_float] * 5 + \
        [C.c_int32] * 3 + [C.c_void_p]
    g = PeerGroup()
    g.world, g.rank = 9, 0                                  # more ranks than a flag row holds
    assert lib.etp_peer_signal(C.byref(g), 0, 0, 1, None) == -1 and b"world" in lib.etp_last_error()
    g.world, g.rank = 2, 2                                  # rank outside the group
    assert lib.etp_peer_wait(C.byref(g), 0, 0, 1, 1, 1.0, None) == -1
    g.world, g.rank = 2, 0                                  # buffers never mapped
    assert lib.etp_peer_signal(C.byref(g), 0, 0, 1, None) == -1 and b"null buffer" in lib.etp_last_error()
    assert lib.etp_peer_reduce_adamw(C.byref(g), 0, 64, None, None, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 0, 0, None) == -1
    assert lib.etp_ipc_export(None, None, None) == -1 and lib.etp_ipc_open(None, None) == -1
    assert C.sizeof(PeerGroup) == 8 + 4 * 8 * 8             # the struct of the header: 2 x int32 + 4 arrays of 8 pointers
