"""GPU: the drop-in boundary driven the way the reference trainer drives it.

ss_trainer_ETP.py:211-213 wraps ``self.policy.net`` (whose ``vln_bert`` member is the planner) in
``DistributedDataParallel`` and creates ``torch.optim.AdamW``; :499-506 runs one training iteration as
    with autocast(): loss = rollout(...)          # MANY net(mode=...) calls, every step's graph kept alive
    scaler.scale(loss).backward(); scaler.step(optimizer); scaler.update()
This test builds the same stack around ``B200Planner`` (a stand-in ``Net`` with the ETP.forward mode dispatch of
Policy_ViewSelection_ETP.py:157-358, DDP over NCCL with world size 1, autocast, GradScaler, torch AdamW), accumulates
three rollout steps before ONE backward, and checks the gradients that reach ``p.grad`` against the fp32 oracle port
differentiating the same three-step loss, then that the optimizer's update is seen by the next forward."""
import os

import pytest
import torch
import torch.nn as nn

from tests.common import no_dropout

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


class _Net(nn.Module):
    """The slice of ETP(Net).forward the planner sits behind (Policy_ViewSelection_ETP.py:166-170,344-358)."""

    def __init__(self, planner):
        super().__init__()
        self.vln_bert = planner

    def forward(self, mode=None, **kw):
        if mode == "language":
            return self.vln_bert.forward_txt(kw["txt_ids"], kw["txt_masks"])
        if mode == "panorama":
            return self.vln_bert.forward_panorama(kw["rgb_fts"], kw["dep_fts"], kw["loc_fts"], kw["nav_types"], kw["view_lens"])
        if mode == "navigation":
            return self.vln_bert.forward_navigation(kw["txt_embeds"], kw["txt_masks"], kw["gmap_vp_ids"], kw["gmap_step_ids"],
                                                    kw["gmap_img_fts"], kw["gmap_pos_fts"], kw["gmap_masks"],
                                                    kw["gmap_visited_masks"], kw["gmap_pair_dists"])
        raise NotImplementedError(mode)


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _rollout_loss(call, steps, masked_mean=True):
    """Three steps of the rollout's planner calls (ss_trainer_ETP.py:801,837-839,878,890-892,1055): language once, then
    per step panorama -> masked mean of the view embeddings fed into a map node -> navigation -> CE(sum)."""
    txt = call(mode="language", txt_ids=steps[0]["txt_ids"], txt_masks=steps[0]["txt_masks"])
    loss, n = 0.0, 0
    for d in steps:
        pano, pmask = call(mode="panorama", rgb_fts=d["rgb_fts"], dep_fts=d["dep_fts"], loc_fts=d["loc_fts"],
                           nav_types=d["nav_types"], view_lens=d["view_lens"])
        w = pmask.unsqueeze(-1).to(pano.dtype)
        avg = (pano * w).sum(1) / w.sum(1)
        img = d["gmap_img_fts"].clone()
        img[:, 1] = img[:, 1] + avg.to(img.dtype)     # the map keeps live autograd tensors (graph_utils.py:206)
        out = call(mode="navigation", txt_embeds=txt, txt_masks=steps[0]["txt_masks"], gmap_vp_ids=None, gmap_step_ids=d["gmap_step_ids"],
                   gmap_img_fts=img, gmap_pos_fts=d["gmap_pos_fts"], gmap_masks=d["gmap_masks"],
                   gmap_visited_masks=d["gmap_visited_masks"], gmap_pair_dists=d["gmap_pair_dists"])
        loss = loss + torch.nn.functional.cross_entropy(out["global_logits"].float(), d["labels"], reduction="sum", ignore_index=-100)
        n += d["labels"].numel()
    return loss / n


def test_ddp_autocast_gradscaler_rollout_matches_oracle():
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from etpnav_b200.config import PlannerConfig
    from etpnav_b200.planner import B200Planner
    from etpnav_b200.synth import make_inputs, make_weights
    from oracle import planner_port as P
    torch.cuda.set_device(0)
    own_pg = not dist.is_initialized()
    if own_pg:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29731")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        cfg = no_dropout(PlannerConfig(vocab_size=2048, num_l_layers=2, num_x_layers=2))
        sd = make_weights(cfg, seed=51)
        planner = B200Planner(cfg, device="cuda")
        planner.load_state_dict(sd, strict=True)
        net = _Net(planner).train()
        # the reference's wrapper call, argument for argument (ss_trainer_ETP.py:211-212)
        ddp = DDP(net, device_ids=[0], output_device=0, find_unused_parameters=False, broadcast_buffers=False)
        opt = torch.optim.AdamW(ddp.parameters(), lr=1e-3)          # :213
        scaler = torch.cuda.amp.GradScaler()                          # :463
        steps = []
        for t in range(3):
            inp = make_inputs(cfg, 4, 12, 12 + 3 * t, 24, seed=60 + t, ragged=True)   # the map grows over the rollout
            steps.append({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()})
        opt.zero_grad()
        with torch.cuda.amp.autocast():                                # :502
            loss = _rollout_loss(lambda **kw: ddp(**kw), steps)
        scaler.scale(loss).backward()                                  # :504  (ONE backward through three steps' graphs)
        scaler.unscale_(opt)
        grads = {k: p.grad.detach().clone() for k, p in planner.named_parameters() if p.grad is not None}
        before = planner._flat.clone()
        scaler.step(opt)                                               # :505
        scaler.update()                                                # :506
        torch.cuda.synchronize()

        # oracle: the same three-step loss in fp32 autograd
        sdc = {k: v.cuda().clone().requires_grad_(True) for k, v in sd.items()}

        def oracle_call(mode=None, **kw):
            if mode == "language":
                return P.forward_txt(sdc, cfg, kw["txt_ids"], kw["txt_masks"])
            if mode == "panorama":
                return P.forward_panorama(sdc, cfg, kw["rgb_fts"], kw["dep_fts"], kw["loc_fts"], kw["nav_types"], kw["view_lens"])
            return P.forward_navigation(sdc, cfg, kw["txt_embeds"], kw["txt_masks"], None, kw["gmap_step_ids"], kw["gmap_img_fts"],
                                        kw["gmap_pos_fts"], kw["gmap_masks"], kw["gmap_visited_masks"], kw["gmap_pair_dists"])
        loss_o = _rollout_loss(oracle_call, steps)
        loss_o.backward()
        assert abs(loss.item() - loss_o.item()) < 2e-2 * max(1.0, abs(loss_o.item())), (loss.item(), loss_o.item())
        assert len(grads) == len(sd), (len(grads), len(sd))            # every parameter got a gradient (DDP needs that)
        worst = ("", 0.0)
        for k, g in grads.items():
            go = sdc[k].grad
            assert torch.isfinite(g).all(), k
            if go is None or go.norm() < 1e-6:
                continue
            r = _rel(g.float(), go)
            worst = max(worst, (k, r), key=lambda t: t[1])
            assert r < (0.3 if go.numel() == 1 else 0.12), (k, r)
        print("boundary (DDP + autocast + GradScaler, 3-step rollout): worst parameter-gradient rel L2", worst)
        # the optimizer wrote through the parameter views into the flat buffer; the next forward sees the new weights
        assert not torch.equal(before, planner._flat)
        net.eval()
        with torch.no_grad():
            d = steps[0]
            a, _ = ddp(mode="panorama", rgb_fts=d["rgb_fts"], dep_fts=d["dep_fts"], loc_fts=d["loc_fts"], nav_types=d["nav_types"],
                       view_lens=d["view_lens"])
            new_sd = {k: v.detach().clone() for k, v in planner.state_dict().items()}
            b, _ = P.forward_panorama(new_sd, cfg, d["rgb_fts"], d["dep_fts"], d["loc_fts"], d["nav_types"], d["view_lens"])
            old, _ = P.forward_panorama({k: v.cuda() for k, v in sd.items()}, cfg, d["rgb_fts"], d["dep_fts"], d["loc_fts"],
                                        d["nav_types"], d["view_lens"])
        assert (a - b).abs().max().item() < 3e-2
        assert (a - old).abs().max().item() > 2 * (a - b).abs().max().item()   # lr 1e-3 AdamW moved the weights visibly
        # a second iteration goes through the same wrappers (DDP's reducer was re-armed)
        net.train()
        opt.zero_grad()
        with torch.cuda.amp.autocast():
            loss2 = _rollout_loss(lambda **kw: ddp(**kw), steps)
        scaler.scale(loss2).backward()
        scaler.step(opt)
        scaler.update()
        assert torch.isfinite(loss2).item()
    finally:
        if own_pg:
            dist.destroy_process_group()
