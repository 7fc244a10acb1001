"""GPU: host-side pieces of the training / e2e path that touch the device: the one-blob input stager, the bf16
instruction-embedding input of the navigation step, and the flagged / clipped fused AdamW against torch.optim.AdamW with
the reference's parameter groups (pretrain_src/pretrain_src/optim/misc.py:14-20) and clip_grad_norm_ (train_r2r.py:279-284)."""
import ctypes as C

import pytest
import torch

from tests.common import no_dropout

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_stager_one_copy_views_match_host():
    from etpnav_b200.pipeline import HostBatchStager
    dev = torch.device("cuda", 0)
    st = HostBatchStager(dev, bf16_keys=("txt_embeds",))
    g = torch.Generator().manual_seed(3)
    hosts = [{"txt_embeds": torch.randn(5, 7, 768, generator=g), "a": torch.randn(257, 33, generator=g),
              "ids": torch.randint(0, 9, (64, 7), generator=g), "m": torch.rand(64, 9, generator=g) > 0.5, "tag": i}
             for i in range(5)]
    blobs = [st.pack(h) for h in hosts]
    st.submit(*blobs[0])
    for i in range(5):
        d = st.get()
        if i + 1 < 5:
            st.submit(*blobs[i + 1])
        acc = d["a"].clone()
        for _ in range(20):      # keep the compute stream busy while the next blob crosses
            acc = acc * 1.0001 + d["a"]
        assert d["tag"] == i
        torch.cuda.synchronize()
        assert torch.equal(d["a"].cpu(), hosts[i]["a"]) and torch.equal(d["ids"].cpu(), hosts[i]["ids"])
        assert torch.equal(d["m"].cpu(), hosts[i]["m"]) and d["m"].dtype == torch.bool
        assert d["txt_embeds"].dtype == torch.bfloat16
        assert torch.equal(d["txt_embeds"].cpu(), hosts[i]["txt_embeds"].bfloat16())     # host RNE cast == the kernels' cast
    with pytest.raises(RuntimeError):
        st.get()


def test_navigation_takes_bf16_txt_embeds_bit_exactly():
    """etp_nav_inputs.txt_embeds_bf16: handing the step the bf16 image of txt_embeds gives the same bits as handing it the
    fp32 tensor that image was rounded from (the kernels' first act on txt_embeds is that cast), forward and backward."""
    from etpnav_b200.config import PlannerConfig
    from etpnav_b200.planner import B200Planner
    from etpnav_b200.synth import make_inputs, make_weights
    cfg = no_dropout(PlannerConfig(vocab_size=2048, num_l_layers=0, num_x_layers=2))
    m = B200Planner(cfg, device="cuda")
    m.load_state_dict(make_weights(cfg, seed=4), strict=True)
    d = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in make_inputs(cfg, 3, 12, 20, 40, seed=4, ragged=True).items()}
    args = lambda t: (t, d["txt_masks"], None, d["gmap_step_ids"], d["gmap_img_fts"], d["gmap_pos_fts"], d["gmap_masks"],
                      d["gmap_visited_masks"], d["gmap_pair_dists"])
    tb = d["txt_embeds"].bfloat16()
    m.eval()
    with torch.no_grad():
        a = m.forward_navigation(*args(tb))
        b = m.forward_navigation(*args(tb.float()))
    assert torch.equal(a["global_logits"], b["global_logits"]) and torch.equal(a["gmap_embeds"], b["gmap_embeds"])
    m.train()
    grads = []
    for t in (tb.clone().requires_grad_(True), tb.float().requires_grad_(True)):
        m.zero_grad()
        out = m.forward_navigation(*args(t))
        torch.nn.functional.cross_entropy(out["global_logits"], d["labels"], reduction="sum").backward()
        grads.append((t.grad, m._pmap["global_encoder.encoder.x_layers.0.visual_attention.att.key.weight"].grad.clone()))
    assert grads[0][0].dtype == torch.bfloat16 and grads[1][0].dtype == torch.float32
    assert torch.equal(grads[0][0], grads[1][0].bfloat16())
    assert torch.equal(grads[0][1], grads[1][1])


def test_adamw_ex_matches_torch_param_groups_and_clipping():
    from etpnav_b200 import lib as L
    lib = L.lib()
    f32, i32, pv = C.c_float, C.c_int32, C.c_void_p
    lib.etp_adamw_step_ex.argtypes = [pv, pv, pv, pv, pv, C.c_int64, f32, f32, f32, f32, f32, i32, f32, pv, pv, f32, pv]
    lib.etp_grad_sumsq.argtypes = [pv, C.c_int64, pv, pv, pv]
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 64 * 40
    p0 = torch.randn(n, device="cuda", generator=g)
    # three tensors of the flat layout: decayed weight [0, 1280), no-decay bias [1280, 1920), frozen [1920, 2560)
    flags = torch.zeros(n // 64, dtype=torch.uint8, device="cuda")
    flags[:20] = 3
    flags[20:30] = 1
    lr, b1, b2, eps, wd, max_norm, gscale = 1e-2, 0.9, 0.98, 1e-6, 0.01, 5.0, 0.5
    p = p0.clone()
    pb = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    m1, v1 = torch.zeros_like(p), torch.zeros_like(p)
    tw, tb_, tf = (p0[:1280].clone().requires_grad_(True), p0[1280:1920].clone().requires_grad_(True),
                   p0[1920:].clone().requires_grad_(False))
    opt = torch.optim.AdamW([{"params": [tw], "weight_decay": wd}, {"params": [tb_], "weight_decay": 0.0}], lr=lr,
                            betas=(b1, b2), eps=eps)
    for step in range(1, 4):
        graw = torch.randn(n, device="cuda", generator=g) * 20.0          # summed over 2 ranks; the mean is graw * 0.5
        nsq = torch.zeros(1, device="cuda")
        L._check(lib.etp_grad_sumsq(L.ptr(graw), n, L.ptr(flags), L.ptr(nsq), L.stream_ptr()), "sumsq")
        L._check(lib.etp_adamw_step_ex(L.ptr(p), L.ptr(pb), L.ptr(graw), L.ptr(m1), L.ptr(v1), n, lr, b1, b2, eps, wd, step, gscale,
                                       L.ptr(flags), L.ptr(nsq), max_norm, L.stream_ptr()), "adamw_ex")
        tw.grad, tb_.grad = graw[:1280] * gscale, graw[1280:1920] * gscale
        tn = torch.nn.utils.clip_grad_norm_([tw, tb_], max_norm)
        assert abs(float(tn) - float(nsq.sqrt() * gscale)) < 1e-3 * float(tn)       # frozen block not in the norm
        opt.step()
    torch.cuda.synchronize()
    assert (p[:1280] - tw.detach()).abs().max().item() < 2e-6
    assert (p[1280:1920] - tb_.detach()).abs().max().item() < 2e-6
    assert torch.equal(p[1920:], p0[1920:])                                 # frozen: no update, no decay
    assert torch.equal(pb[:1920], p[:1920].bfloat16())


def test_overlapped_trainer_step_equals_the_sequential_one():
    """PlannerTrainer runs the panorama branch on its own stream next to the instruction-side GEMMs of the navigation
    call (no autograd graph, events for gmap_img_fts / d_gmap_img_fts).  Same dropout seeds, same kernels: logits and the
    whole flat gradient must agree with the autograd-driven sequential step up to the re-association of the masked mean
    (and fp32 atomics order); repeated steps keep agreeing (no stream race that would only show up later)."""
    from etpnav_b200.config import PlannerConfig
    from etpnav_b200.planner import B200Planner
    from etpnav_b200.synth import make_inputs, make_weights
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=0, num_x_layers=3)       # dropout on (the reference's 0.1)
    sd = make_weights(cfg, seed=17)
    d = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in make_inputs(cfg, 16, 12, 40, 90, seed=17, ragged=True).items()}
    outs = []
    for overlap in (False, True):
        m = B200Planner(cfg, device="cuda")
        m.load_state_dict(sd, strict=True)
        m.train()
        m.set_dropout_seed(99)
        tr = m.make_trainer(lr=1e-4, overlap=overlap)
        assert tr.overlap == overlap
        rec = []
        for _ in range(3):
            tr.zero_grad()
            logits, loss = tr.forward_backward(d)
            torch.cuda.synchronize()
            rec.append((logits.detach().clone(), m._direct_grad[tr.lo:tr.hi].clone(), loss.clone()))
            tr.optimizer_step()
        torch.cuda.synchronize()
        outs.append((rec, m._flat[tr.lo:tr.hi].clone()))
    (seq, p_seq), (ovl, p_ovl) = outs
    assert torch.equal(seq[0][0], ovl[0][0])        # first step: same weights, inputs, masks, kernels -> the same logits, bit for bit
    assert ((seq[0][1] - ovl[0][1]).norm() / seq[0][1].norm()).item() < 1e-4   # gradients up to the order of fp32 atomics
    for (lg_a, g_a, loss_a), (lg_b, g_b, loss_b) in zip(seq, ovl):
        fin = ~torch.isinf(lg_a)
        assert torch.equal(torch.isinf(lg_a), torch.isinf(lg_b))
        assert (lg_a[fin] - lg_b[fin]).abs().max().item() < 5e-2      # later steps: bf16 rounding flips after tiny parameter differences
        assert abs(loss_a.item() - loss_b.item()) < 2e-2
        assert ((g_a - g_b).norm() / g_a.norm()).item() < 1e-1      # (5.1e-2 seen at the third step; a stream race is O(1))
    # (AdamW's first steps move every element by ~lr whatever the gradient's size: elements whose tiny gradients differ in
    # the last bits may move apart by 2 lr; the parameters as a whole stay together)
    assert ((p_seq - p_ovl).norm() / p_seq.norm()).item() < 5e-3


def test_step_clears_gradients_behind_the_update():
    """PlannerTrainer.step() clears every bucket's gradient slice on the update stream right behind its AdamW instead of
    starting the next step with a memset on the compute stream: the buffer reads zero after step(), and the gradient the
    NEXT step accumulates equals the one a fresh trainer computes from the same weights (nothing stale, nothing cleared
    late)."""
    from etpnav_b200.config import PlannerConfig
    from etpnav_b200.planner import B200Planner
    from etpnav_b200.synth import make_inputs, make_weights
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=0, num_x_layers=2)
    d = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in make_inputs(cfg, 8, 12, 30, 50, seed=3, ragged=True).items()}
    m = B200Planner(cfg, device="cuda")
    m.load_state_dict(make_weights(cfg, seed=3), strict=True)
    m.train()
    m.set_dropout_seed(11)
    tr = m.make_trainer(lr=1e-4)
    tr.step(d)
    tr.step(d)
    torch.cuda.synchronize()
    assert tr._grads_clean and m._direct_grad[tr.lo:tr.hi].abs().max().item() == 0.0
    w1 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.set_dropout_seed(5)
    tr.step(d, keep_grads=True)
    torch.cuda.synchronize()
    g2 = m._direct_grad[tr.lo:tr.hi].clone()
    assert not tr._grads_clean and g2.abs().max().item() > 0
    m3 = B200Planner(cfg, device="cuda")
    m3.load_state_dict(w1, strict=True)
    m3.train()
    tr3 = m3.make_trainer(lr=1e-4)
    m3.set_dropout_seed(5)
    tr3.zero_grad()
    tr3.forward_backward(d)
    torch.cuda.synchronize()
    g3 = m3._direct_grad[tr3.lo:tr3.hi]
    assert ((g2 - g3).norm() / g3.norm()).item() < 1e-4
    tr.step(d)           # a kept gradient is cleared explicitly before the next accumulation
    torch.cuda.synchronize()
    assert m._direct_grad[tr.lo:tr.hi].abs().max().item() == 0.0


def test_pipelined_steps_match_plain_steps():
    """step(pipelined=True): the compute stream returns after the navigation buckets' update, the panorama buckets finish
    under the next step (picked up by its panorama stream).  (1) Ordering, made visible: a 20 ms spin is appended to the
    update stream after a pipelined step — the next step's panorama branch must not start before it ends.  (2) The logits of every step and the final parameters stay with those of plain
    steps (same seeds; differences = fp32 atomics order amplified by AdamW's sign-like first steps), and join() leaves
    nothing pending."""
    from etpnav_b200.config import PlannerConfig
    from etpnav_b200.planner import B200Planner
    from etpnav_b200.synth import make_inputs, make_weights
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=0, num_x_layers=3)
    sd = make_weights(cfg, seed=23)
    ds = [{k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in make_inputs(cfg, 16, 12, 40, 90, seed=23 + i, ragged=True).items()}
          for i in range(2)]
    outs = []
    for pipelined in (False, True):
        m = B200Planner(cfg, device="cuda")
        m.load_state_dict(sd, strict=True)
        m.train()
        m.set_dropout_seed(7)
        tr = m.make_trainer(lr=1e-4)
        lgs = [tr.step(ds[i % 2], pipelined=pipelined).detach().clone() for i in range(3)]
        assert tr._pending_join == pipelined
        if pipelined:
            spin_end = torch.cuda.Event(enable_timing=True)
            tr._debug_pano_start = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(tr.side):
                torch.cuda._sleep(40_000_000)            # ~20 ms at 1.9 GHz, behind the panorama buckets' update
                spin_end.record()
            lgs.append(tr.step(ds[1], pipelined=True).detach().clone())
            tr.join()
            torch.cuda.synchronize()
            assert spin_end.elapsed_time(tr._debug_pano_start) >= 0.0       # panorama branch started after the spin ended
            tr._debug_pano_start = None
        else:
            lgs.append(tr.step(ds[1]).detach().clone())
        tr.join()
        assert not tr._pending_join
        torch.cuda.synchronize()
        assert m._direct_grad[tr.lo:tr.hi].abs().max().item() == 0.0
        outs.append((lgs, m._flat[tr.lo:tr.hi].clone()))
    (la, pa), (lb, pb) = outs
    assert torch.equal(la[0], lb[0])
    for x, y in zip(la, lb):
        fin = ~torch.isinf(x)
        assert torch.equal(torch.isinf(x), torch.isinf(y)) and (x[fin] - y[fin]).abs().max().item() < 1e-1   # 4.9e-2 seen after six steps
    assert ((pa - pb).norm() / pa.norm()).item() < 1e-2
