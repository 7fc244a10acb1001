"""GPU: the episode-level text K|V cache (``B200Planner.encode_text_kv`` -> ``TextKV`` -> ``forward_navigation``).
The reference recomputes the instruction's key / value projections at every step of every cross-modal layer although
``txt_embeds`` is constant over the episode (vilmodel_cmt.py:326-328) and re-indexes ``all_txt_embeds[not_done_index]`` as
episodes finish (ss_trainer_ETP.py:819-821).  The cached path must give the SAME logits as the un-cached one (same GEMM,
same attention kernel, only the batch coordinate of the K / V tiles goes through a row map) and stay within the bf16
envelope of the fp32 oracle."""
import pytest
import torch

from tests.common import BF16_LOGIT_TOL

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _setup(B=6, N=30, L=150, X=3):
    from etpnav_b200.config import PlannerConfig
    from etpnav_b200.planner import B200Planner
    from etpnav_b200.synth import make_inputs, make_weights
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=0, num_x_layers=X)
    sd = make_weights(cfg, seed=13)
    m = B200Planner(cfg, device="cuda")
    m.load_state_dict(sd, strict=True)
    inp = make_inputs(cfg, B, 12, N, L, seed=13, ragged=True)
    d = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    return cfg, sd, inp, d, m.eval()


def _nav(m, txt, d, sel=slice(None)):
    return m.forward_navigation(txt, d["txt_masks"][sel], None, d["gmap_step_ids"][sel], d["gmap_img_fts"][sel],
                                d["gmap_pos_fts"][sel], d["gmap_masks"][sel], d["gmap_visited_masks"][sel], d["gmap_pair_dists"][sel])


def test_cached_text_kv_equals_uncached_and_tracks_the_shrinking_batch():
    from oracle import planner_port as P
    cfg, sd, inp, d, m = _setup()
    with torch.no_grad():
        kv = m.encode_text_kv(d["txt_embeds"])                    # once per episode
        ref = _nav(m, d["txt_embeds"], d)
        got = _nav(m, kv, d)
        assert torch.equal(got["global_logits"], ref["global_logits"]) and torch.equal(got["gmap_embeds"], ref["gmap_embeds"])
        # episodes 1 and 4 finished: the trainer keeps rows not_done_index of every per-episode tensor
        keep = torch.tensor([0, 2, 3, 5], device="cuda")
        sub = _nav(m, kv[keep], d, keep)
        fin = ~torch.isinf(ref["global_logits"][keep])
        assert torch.equal(torch.isinf(sub["global_logits"]), ~fin)
        assert (sub["global_logits"][fin] - ref["global_logits"][keep][fin]).abs().max().item() < 1e-5
        assert (sub["gmap_embeds"] - ref["gmap_embeds"][keep]).abs().max().item() < 1e-4
        # a second shrink composes with the first (rows of the ORIGINAL episode batch are tracked)
        keep2 = torch.tensor([True, False, True, True], device="cuda")
        sub2 = _nav(m, kv[keep][keep2], d, keep[keep2])
        assert (sub2["gmap_embeds"] - ref["gmap_embeds"][keep[keep2]]).abs().max().item() < 1e-4
        # bf16 instruction embeddings feed the cache as well
        kvb = m.encode_text_kv(d["txt_embeds"].bfloat16())
        assert torch.equal(kvb.kv_all, kv.kv_all)
    # and the whole thing sits inside the bf16 envelope of the fp32 oracle
    o = P.forward_navigation(sd, cfg, inp["txt_embeds"], inp["txt_masks"], None, inp["gmap_step_ids"], inp["gmap_img_fts"],
                             inp["gmap_pos_fts"], inp["gmap_masks"], inp["gmap_visited_masks"], inp["gmap_pair_dists"])
    lo = o["global_logits"]
    f = ~torch.isinf(lo)
    assert torch.equal(got["global_logits"].cpu().argmax(1), lo.argmax(1))
    assert (got["global_logits"].cpu()[f] - lo[f]).abs().max().item() < BF16_LOGIT_TOL


def test_text_kv_is_inference_only_and_checks_shapes():
    from etpnav_b200 import lib as L
    cfg, sd, inp, d, m = _setup(B=3, N=12, L=40, X=1)
    with torch.no_grad():
        kv = m.encode_text_kv(d["txt_embeds"])
    m.train()
    with pytest.raises(L.EtpError):
        _nav(m, kv, d)                                            # parameters require grad: the backward would need txt_embeds
    m.eval()
    with torch.no_grad(), pytest.raises(ValueError):
        _nav(m, kv[torch.tensor([0, 1], device="cuda")], d)       # handle of 2 episodes, step batch of 3
