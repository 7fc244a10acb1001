"""CPU: dry run of the pre-training twin's HOST plumbing with the C library stubbed out (every entry point returns 0 and
writes nothing): exercises the autograd graph wiring, argument marshalling, CSR construction, gradient-buffer views and
the returned gradient arity of every custom Function, which cannot otherwise run without a GPU.  Numbers are garbage by
construction; only structure is asserted.  (The real kernels are checked by tests/test_pretrain_gpu.py.)"""
import ctypes as C

import pytest
import torch

from etpnav_b200.config import PlannerConfig
from etpnav_b200.synth import make_traj_batch


class _StubLib:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        class Fn:
            restype = None
            argtypes = None

            def __call__(fn, *a):
                self.calls.append(name)
                if name.endswith("_bytes"):
                    return 4096
                return 0
        f = Fn()
        setattr(self, name, f)
        return f


@pytest.fixture()
def stub(monkeypatch):
    from etpnav_b200 import lib as L
    from etpnav_b200 import planner, pretrain
    s = _StubLib()
    monkeypatch.setattr(L, "lib", lambda: s)
    monkeypatch.setattr(L, "require_device", lambda: None)
    monkeypatch.setattr(L, "stream_ptr", lambda: C.c_void_p(0))
    monkeypatch.setattr(L, "ptr", lambda t: None if t is None else C.c_void_p(t.data_ptr()))
    monkeypatch.setattr(planner, "_declared", True)
    monkeypatch.setattr(pretrain, "_declared", True)
    monkeypatch.setattr(planner.B200Planner, "_refresh_cache", _fake_refresh)
    return s


def _fake_refresh(self):
    if self._structs is None:
        self._structs = self._build_structs(self._flat.data_ptr(), self._flat_bf16.data_ptr(), 2)


def test_mlm_and_sap_graph_wiring(stub):
    from etpnav_b200.pretrain import B200PreTraining
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=1, num_x_layers=2, hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, pred_head_dropout_prob=0.0)
    m = B200PreTraining(cfg, device="cpu").train()
    b = make_traj_batch(cfg, 3, 3, 6, 12, seed=0)
    b["traj_view_img_fts"] = b["traj_view_img_fts"].clone().requires_grad_(True)
    mlm = m(b, "mlm")
    sap = m(b, "sap")
    n_masked = int((b["txt_labels"] != -1).sum())
    assert mlm.shape == (n_masked,) and sap.shape == (3,)
    torch.nan_to_num(mlm, nan=0.0, posinf=0.0, neginf=0.0).sum().backward()
    torch.nan_to_num(sap, nan=0.0, posinf=0.0, neginf=0.0).sum().backward()
    assert b["traj_view_img_fts"].grad is not None and b["traj_view_img_fts"].grad.shape == b["traj_view_img_fts"].shape
    got = {k for k, p in m.bert._pmap.items() if p.grad is not None}
    assert {k for k in m.bert._pmap if ".lang_" in k} <= got
    assert {k for k in m.bert._pmap if k.startswith("mlm_head.")} <= got
    assert "embeddings.word_embeddings.weight" in got and "global_sap_head.net.0.weight" in got
    for k in got:
        assert m.bert._pmap[k].grad.shape == m.bert._pmap[k].shape, k
    for fn in ("etp_forward_txt", "etp_forward_panorama", "etp_segment_gather", "etp_forward_lang2visn",
               "etp_backward_lang2visn", "etp_forward_navigation", "etp_backward_navigation", "etp_backward_panorama",
               "etp_backward_txt", "etp_gemm", "etp_layernorm_fwd", "etp_layernorm_bwd"):
        assert fn in stub.calls, fn
    # eval / no_grad path
    m.eval()
    with torch.no_grad():
        scores = m(b, "mlm", compute_loss=False)
        logits, labels = m(b, "sap", compute_loss=False)
        emb = m.bert(*[b[k] for k in ("txt_ids", "txt_lens", "traj_view_img_fts", "traj_view_dep_fts", "traj_obj_img_fts",
                                      "traj_loc_fts", "traj_nav_types", "traj_step_lens", "traj_vp_view_lens",
                                      "traj_vp_obj_lens", "traj_vpids", "traj_cand_vpids", "gmap_lens", "gmap_step_ids",
                                      "gmap_pos_fts", "gmap_pair_dists", "gmap_vpids")])
    assert scores.shape == (n_masked, cfg.vocab_size) and logits.shape == b["gmap_step_ids"].shape
    assert emb.shape == (*b["gmap_step_ids"].shape, 768)


def test_trainer_wiring(stub):
    from etpnav_b200.pretrain import B200PreTraining, PretrainTrainer
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=1, num_x_layers=2)
    model = B200PreTraining(cfg, device="cpu").train()
    b = make_traj_batch(cfg, 2, 3, 6, 12, seed=1)
    tr = PretrainTrainer(model, world_size=1)
    for task in ("mlm", "sap"):
        loss = tr.step(b, task)
        assert loss.dim() == 0
    assert tr.t == 2 and "etp_adamw_step_ex" in stub.calls and "etp_grad_sumsq" in stub.calls
    assert all(p.grad is None for p in model.bert._pmap.values())   # gradients live in the flat buffer only
