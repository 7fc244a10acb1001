"""CPU, world_size 2 over gloo: the host-side data-parallel logic (batch sharding, per-rank shards, flat
gradient all-reduce + 1/world scale == DDP mean)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from etpnav_b200.config import PlannerConfig
from etpnav_b200.dist import allreduce_buckets_, allreduce_flat_, gradient_buckets, rank_seed, shard_batch
from etpnav_b200.synth import make_inputs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=0)
    inp = make_inputs(cfg, 2, 12, 8, 16, seed=rank_seed(100, rank), ragged=False)
    g = torch.full((1000,), float(rank + 1))
    g[:10] = inp["gmap_img_fts"].flatten()[:10]
    scale = allreduce_flat_(g, world)
    # bucketed variant (the trainer's path): same result as one flat all-reduce
    from etpnav_b200.layout import FlatLayout
    cfg2 = PlannerConfig(vocab_size=64, num_l_layers=0, num_x_layers=3)
    lay = FlatLayout(cfg2)
    lo = min(lay.group_ranges[k][0] for k in ("pano", "nav"))
    hi = max(lay.group_ranges[k][1] for k in ("pano", "nav"))
    gb = torch.arange(hi - lo, dtype=torch.float32) * (rank + 1)
    buckets = [(n, a - lo, b - lo) for n, a, b in gradient_buckets(lay, cfg2, lo, hi)]
    sb = allreduce_buckets_(gb, buckets, world)
    torch.save({"g": g * scale, "img": inp["gmap_img_fts"][:, :2, :4].clone(), "gb": gb * sb,
                "cover": sorted((a, b) for _, a, b in buckets), "n": hi - lo,
                "names": [n for n, _, _ in buckets]}, f"{out}/r{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_mean_and_distinct_shards(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["g"], r1["g"])                      # every rank holds the same reduced gradient
    assert torch.allclose(r0["g"][10:], torch.full((990,), 1.5))  # mean of 1 and 2
    assert not torch.equal(r0["img"], r1["img"])              # ranks drew different shards
    # buckets: completion order (last x-layer first, "rest" last), disjoint, covering the whole slice; mean of x and 2x
    assert r0["names"] == ["x_layer_2", "x_layer_1", "x_layer_0", "nav_head", "rest"]
    cov = r0["cover"]
    assert cov[0][0] == 0 and cov[-1][1] == r0["n"] and all(cov[i][1] == cov[i + 1][0] for i in range(len(cov) - 1))
    assert torch.equal(r0["gb"], r1["gb"])
    assert torch.allclose(r0["gb"], torch.arange(r0["n"], dtype=torch.float32) * 1.5)


def test_shard_batch_covers_global_batch():
    for gb, w in [(512, 8), (256, 8), (10, 4), (7, 2)]:
        spans = [shard_batch(gb, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
