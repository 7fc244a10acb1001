"""GPU: train()-mode dropout of the CUDA path against the fp32 oracle fed with THE SAME masks.

The kernels never store a mask: a keep flag is a pure function of (seed, site, element index)
(etpnav_b200/csrc/common.cuh: Drop).  ``etp_dropout_mask`` evaluates that function for one site, the oracle's
``drop`` hook (oracle/planner_port.py) multiplies by mask / (1 - p) at the matching nn.Dropout of the reference
(vilmodel_cmt.py:76,127,152,191,346,657,711; common/transformer.py:163-181).  With identical masks the two paths
compute the same function, so outputs and gradients must agree to the usual bf16-operand tolerances; a wrong site
id, index formula or a mask missing in the backward shows up as an O(1) error.
"""
import ctypes as C

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


class _KernelMasks:
    """drop(x, kind, site) hook for the oracle that reproduces the kernels' masks of ONE forward call."""

    def __init__(self, drop_struct, cfg):
        from etpnav_b200 import lib as L
        self.L, self.d, self.cfg = L, drop_struct, cfg
        self.lib = L.lib()
        self.lib.etp_dropout_mask.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p]
        self.sites = []

    def __call__(self, x, kind, site):
        p = {"hidden": self.cfg.hidden_dropout_prob, "attn": self.cfg.attention_probs_dropout_prob,
             "head": self.cfg.pred_head_dropout_prob}[kind]
        if p <= 0:
            return x
        n = x.numel()
        m = torch.empty(n, dtype=torch.uint8, device="cuda")
        self.L._check(self.lib.etp_dropout_mask(C.byref(self.d), p, site, n, m.data_ptr(), None), "etp_dropout_mask")
        torch.cuda.synchronize()
        self.sites.append((site, kind, float(m.float().mean().item())))
        return x * (m.view(x.shape).to(x.dtype).to(x.device) / (1.0 - p))


def _next_struct(m):
    """The etp_dropout the module will use for its NEXT forward call (same seed stream, not advanced)."""
    from etpnav_b200.planner import Dropout
    calls = m._drop_calls
    d = m._next_dropout()
    m._drop_calls = calls
    return Dropout(d.seed, d.p_hidden, d.p_attn, d.p_head)


def _setup(x_layers=2, l_layers=1, p=(0.1, 0.1, 0.1)):
    from etpnav_b200.config import PlannerConfig
    from etpnav_b200.planner import B200Planner
    from etpnav_b200.synth import make_inputs, make_weights
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=l_layers, num_x_layers=x_layers, hidden_dropout_prob=p[0],
                        attention_probs_dropout_prob=p[1], pred_head_dropout_prob=p[2])
    sd = make_weights(cfg, seed=31)
    inp = make_inputs(cfg, 3, 13, 21, 45, seed=31, ragged=True)
    m = B200Planner(cfg, device="cuda")
    m.load_state_dict(sd, strict=True)
    m.train()
    m.set_dropout_seed(1234)
    return cfg, sd, inp, m


def test_mask_statistics_and_determinism():
    from etpnav_b200 import lib as L
    from etpnav_b200.planner import Dropout
    lib = L.lib()
    lib.etp_dropout_mask.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p]
    d = Dropout(99, 0.1, 0.1, 0.1)
    n = 1 << 20
    a, b, c = (torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(3))
    L._check(lib.etp_dropout_mask(C.byref(d), 0.1, 1001, n, a.data_ptr(), None), "mask")
    L._check(lib.etp_dropout_mask(C.byref(d), 0.1, 1001, n, b.data_ptr(), None), "mask")
    L._check(lib.etp_dropout_mask(C.byref(d), 0.1, 1002, n, c.data_ptr(), None), "mask")
    torch.cuda.synchronize()
    assert torch.equal(a, b)                                   # pure function of (seed, site, index)
    keep = a.float().mean().item()
    assert abs(keep - 0.9) < 3e-3, keep                         # Bernoulli(0.9): sigma = 3e-4 at n = 2^20
    assert abs((a != c).float().mean().item() - 0.18) < 5e-3    # another site: independent stream
    # no visible structure between neighbours (the two halves of one hash word)
    x = a.float().view(-1, 2)
    cov = ((x[:, 0] - keep) * (x[:, 1] - keep)).mean().item()
    assert abs(cov) < 1e-3, cov
    d0 = Dropout(99, 0.0, 0.0, 0.0)
    L._check(lib.etp_dropout_mask(C.byref(d0), 0.0, 5, n, a.data_ptr(), None), "mask")
    assert bool(a.all())


def test_eval_mode_has_no_dropout_and_train_mode_does():
    cfg, sd, inp, m = _setup()
    d = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    args = (d["rgb_fts"], d["dep_fts"], d["loc_fts"], d["nav_types"], d["view_lens"])
    with torch.no_grad():
        m.eval()
        a, _ = m.forward_panorama(*args)
        b, _ = m.forward_panorama(*args)
        assert torch.equal(a, b)
        m.train()
        m.set_dropout_seed(7)
        c, _ = m.forward_panorama(*args)
        e, _ = m.forward_panorama(*args)          # next seed of the stream
        m.set_dropout_seed(7)
        f, _ = m.forward_panorama(*args)          # same seed again
    assert not torch.equal(a, c) and not torch.equal(c, e) and torch.equal(c, f)


def test_panorama_and_navigation_match_oracle_with_the_same_masks():
    from oracle import planner_port as P
    cfg, sd, inp, m = _setup()
    d = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    sdc = {k: v.cuda().clone().requires_grad_(True) for k, v in sd.items()}
    leaves = {k: d[k].clone().requires_grad_(True) for k in ("rgb_fts", "dep_fts", "gmap_img_fts")}
    with torch.no_grad():
        m.eval()
        txt = m.forward_txt(d["txt_ids"], d["txt_masks"])
        m.train()
    txt_k = txt.clone().requires_grad_(True)

    # ---- CUDA path (records which etp_dropout each call used)
    dp = _next_struct(m)
    pano, pmask = m.forward_panorama(leaves["rgb_fts"], leaves["dep_fts"], d["loc_fts"], d["nav_types"], d["view_lens"])
    dn = _next_struct(m)
    nav = m.forward_navigation(txt_k, d["txt_masks"], None, d["gmap_step_ids"], leaves["gmap_img_fts"], d["gmap_pos_fts"],
                               d["gmap_masks"], d["gmap_visited_masks"], d["gmap_pair_dists"])
    g = torch.Generator().manual_seed(3)
    wp = (torch.randn(pano.shape, generator=g).cuda() * pmask[..., None])
    we = (torch.randn(nav["gmap_embeds"].shape, generator=g).cuda() * d["gmap_masks"][..., None])
    loss = (torch.nn.functional.cross_entropy(nav["global_logits"], d["labels"], reduction="sum")
            + (pano * wp).sum() * 0.01 + (nav["gmap_embeds"] * we).sum() * 0.01)
    loss.backward()

    # ---- oracle with the kernels' masks
    o_leaves = {k: d[k].clone().requires_grad_(True) for k in leaves}
    txt_o = txt.clone().requires_grad_(True)
    hp, hn = _KernelMasks(dp, cfg), _KernelMasks(dn, cfg)
    pano_o, pm_o = P.forward_panorama(sdc, cfg, o_leaves["rgb_fts"], o_leaves["dep_fts"], d["loc_fts"], d["nav_types"],
                                      d["view_lens"], drop=hp)
    nav_o = P.forward_navigation(sdc, cfg, txt_o, d["txt_masks"], None, d["gmap_step_ids"], o_leaves["gmap_img_fts"],
                                 d["gmap_pos_fts"], d["gmap_masks"], d["gmap_visited_masks"], d["gmap_pair_dists"], drop=hn)
    loss_o = (torch.nn.functional.cross_entropy(nav_o["global_logits"], d["labels"], reduction="sum")
              + (pano_o * wp).sum() * 0.01 + (nav_o["gmap_embeds"] * we).sum() * 0.01)
    loss_o.backward()
    assert len(hp.sites) == 1 + 4 * cfg.num_pano_layers and len(hn.sites) == 5 * cfg.num_x_layers + 1
    for site, kind, keep in hp.sites + hn.sites:
        assert 0.85 < keep < 0.95, (site, kind, keep)

    # forward: same tolerances as the dropout-free parity tests (tests/test_planner_gpu.py)
    valid = pmask[..., None].expand_as(pano)
    assert (pano - pano_o)[valid].abs().max() < 6e-2
    lg, lo = nav["global_logits"], nav_o["global_logits"]
    assert torch.equal(torch.isinf(lg), torch.isinf(lo))
    fin = ~torch.isinf(lo)
    assert (lg[fin] - lo[fin]).abs().max() < 5e-2
    gm = d["gmap_masks"][..., None].expand_as(nav["gmap_embeds"])
    assert (nav["gmap_embeds"] - nav_o["gmap_embeds"])[gm].abs().max() < 8e-2
    # backward: relative L2 of the activation gradients and of every parameter gradient norm
    assert _rel(txt_k.grad, txt_o.grad) < 8e-2, _rel(txt_k.grad, txt_o.grad)
    for k in leaves:
        assert _rel(leaves[k].grad, o_leaves[k].grad) < 8e-2, (k, _rel(leaves[k].grad, o_leaves[k].grad))
    worst = ("", 0.0)
    for n, prm in m.named_parameters():
        if prm.grad is None or sdc[n].grad is None:
            continue
        go = sdc[n].grad
        if go.norm() < 1e-5:   # analytically zero (key biases: softmax shift invariance): fp32 round-off in the oracle
            assert prm.grad.norm() < 1e-2, n
            continue
        r = _rel(prm.grad, go)
        if r > worst[1]:
            worst = (n, r)
        # the two sprel_linear scalars are sums with heavy cancellation over only B*N*N terms here: bf16 noise is larger
        assert r < (0.25 if go.numel() == 1 else 0.12), (n, r)
    print("dropout parity: worst parameter-gradient rel L2", worst)


def test_txt_matches_oracle_with_the_same_masks():
    from oracle import planner_port as P
    cfg, sd, inp, m = _setup(x_layers=0, l_layers=2)
    d = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    sdc = {k: v.cuda().clone().requires_grad_(True) for k, v in sd.items()}
    dt = _next_struct(m)
    out = m.forward_txt(d["txt_ids"], d["txt_masks"])
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(out.shape, generator=g).cuda() * d["txt_masks"][..., None])
    (out * w).sum().backward()
    ht = _KernelMasks(dt, cfg)
    out_o = P.forward_txt(sdc, cfg, d["txt_ids"], d["txt_masks"], drop=ht)
    (out_o * w).sum().backward()
    tm = d["txt_masks"][..., None].expand_as(out)
    assert (out - out_o)[tm].abs().max() < 8e-2
    for n, prm in m.named_parameters():
        if prm.grad is None or sdc[n].grad is None:
            continue
        if sdc[n].grad.norm() < 1e-5:   # analytically zero (key biases)
            assert prm.grad.norm() < 1e-2, n
            continue
        assert _rel(prm.grad, sdc[n].grad) < 0.12, (n, _rel(prm.grad, sdc[n].grad))


def test_benched_c3_configuration_with_dropout_matches_oracle():
    """The configuration bench.py times (BASELINE.json configs[2]: B=64, 12 views, 80 nodes, 200 tokens, SIX cross-modal
    layers, train() dropout on) against the fp32 oracle fed with the kernels' own masks.  The CUDA path runs the full
    batch; the oracle runs the first S episodes — a dropout mask is a function of the element index in the full
    tensor with the batch outermost, so the first S episodes' flags are a prefix of every site's stream."""
    from etpnav_b200.config import PlannerConfig
    from etpnav_b200.planner import B200Planner
    from etpnav_b200.synth import make_inputs, make_weights
    from oracle import planner_port as P
    from tests.common import BF16_LOGIT_TOL
    B, V, N, L, X, S = 64, 12, 80, 200, 6, 2
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=0, num_x_layers=X)   # dropout probabilities: the reference's 0.1
    assert cfg.hidden_dropout_prob == 0.1 and cfg.attention_probs_dropout_prob == 0.1
    sd = make_weights(cfg, seed=41)
    inp = make_inputs(cfg, B, V, N, L, seed=41, ragged=False)
    m = B200Planner(cfg, device="cuda")
    m.load_state_dict(sd, strict=True)
    m.train()
    m.set_dropout_seed(4321)
    d = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    txt_k = d["txt_embeds"].clone().requires_grad_(True)
    img_k = d["gmap_img_fts"].clone().requires_grad_(True)
    rgb_k = d["rgb_fts"].clone().requires_grad_(True)
    dp = _next_struct(m)
    pano, pmask = m.forward_panorama(rgb_k, d["dep_fts"], d["loc_fts"], d["nav_types"], d["view_lens"])
    dn = _next_struct(m)
    nav = m.forward_navigation(txt_k, d["txt_masks"], None, d["gmap_step_ids"], img_k, d["gmap_pos_fts"], d["gmap_masks"],
                               d["gmap_visited_masks"], d["gmap_pair_dists"])
    g = torch.Generator().manual_seed(3)
    wp = torch.randn(S, V, 768, generator=g).cuda()
    we = torch.randn(S, N, 768, generator=g).cuda()
    # the loss only looks at the first S episodes, so the two sides differentiate the same function
    loss = (torch.nn.functional.cross_entropy(nav["global_logits"][:S], d["labels"][:S], reduction="sum")
            + (pano[:S] * wp).sum() * 0.01 + (nav["gmap_embeds"][:S] * we).sum() * 0.01)
    loss.backward()
    torch.cuda.synchronize()

    sdc = {k: v.cuda().clone().requires_grad_(True) for k, v in sd.items()}
    s = {k: (v[:S].clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    txt_o, img_o, rgb_o = (s[k].clone().requires_grad_(True) for k in ("txt_embeds", "gmap_img_fts", "rgb_fts"))
    hp, hn = _KernelMasks(dp, cfg), _KernelMasks(dn, cfg)
    pano_o, _ = P.forward_panorama(sdc, cfg, rgb_o, s["dep_fts"], s["loc_fts"], s["nav_types"], s["view_lens"], drop=hp)
    nav_o = P.forward_navigation(sdc, cfg, txt_o, s["txt_masks"], None, s["gmap_step_ids"], img_o, s["gmap_pos_fts"],
                                 s["gmap_masks"], s["gmap_visited_masks"], s["gmap_pair_dists"], drop=hn)
    loss_o = (torch.nn.functional.cross_entropy(nav_o["global_logits"], s["labels"], reduction="sum")
              + (pano_o * wp).sum() * 0.01 + (nav_o["gmap_embeds"] * we).sum() * 0.01)
    loss_o.backward()
    assert len(hn.sites) == 5 * X + 1
    lg, lo = nav["global_logits"][:S], nav_o["global_logits"]
    assert torch.equal(torch.isinf(lg), torch.isinf(lo))
    fin = ~torch.isinf(lo)
    assert torch.equal(lg.argmax(1), lo.argmax(1))
    e_log = (lg[fin] - lo[fin]).abs().max().item()
    e_emb = (nav["gmap_embeds"][:S] - nav_o["gmap_embeds"]).abs().max().item()
    e_pano = (pano[:S] - pano_o).abs().max().item()
    print(f"c3 as benched (X=6, dropout on): logit {e_log:.4g} embed {e_emb:.4g} pano {e_pano:.4g}")
    # dropout scales kept activations by 1/(1-p) and six layers compound: allow 2x the dropout-free envelope
    assert e_log < 2 * BF16_LOGIT_TOL, e_log
    assert e_emb < 8e-2 and e_pano < 6e-2, (e_emb, e_pano)
    assert _rel(txt_k.grad[:S], txt_o.grad) < 0.1, _rel(txt_k.grad[:S], txt_o.grad)
    assert _rel(img_k.grad[:S], img_o.grad) < 0.1
    assert _rel(rgb_k.grad[:S], rgb_o.grad) < 0.1
    assert txt_k.grad[S:].abs().max().item() == 0.0 and img_k.grad[S:].abs().max().item() == 0.0   # episodes are independent
    worst = ("", 0.0)
    for n, prm in m.named_parameters():
        go = sdc[n].grad
        if prm.grad is None or go is None:
            continue
        if go.norm() < 1e-5:
            assert prm.grad.norm() < 1e-2, n
            continue
        r = _rel(prm.grad, go)
        worst = max(worst, (n, r), key=lambda t: t[1])
        assert r < (0.3 if go.numel() == 1 else 0.15), (n, r)
    print("c3 as benched: worst full-tensor parameter-gradient rel L2", worst)
