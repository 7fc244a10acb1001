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
    assert r0["names"] == ["x_layer_2", "x_layer_1", "x_layer_0", "nav_head", "pano_layer_1", "rest"]
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


# ----------------------------------------------------------------------------------------------------
# pre-training trainer: the one exchange of its data-parallel path is the SUM all-reduce of the flat
# gradient buffer followed by AdamW with grad_scale = 1 / world.  Two gloo ranks drive PretrainTrainer.step with the C
# library stubbed (tests/test_pretrain_dryrun_cpu.py), each filling the flat gradient with a rank-dependent pattern where
# the backward would: after the step both ranks must hold the same reduced buffer and have passed 1 / world to AdamW.
# ----------------------------------------------------------------------------------------------------
def _pretrain_worker(rank, world, port, out):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from etpnav_b200 import lib as L
    from etpnav_b200 import planner, pretrain
    from etpnav_b200.synth import make_traj_batch
    from tests.test_pretrain_dryrun_cpu import _StubLib, _fake_refresh
    stub = _StubLib()
    seen = {}

    class Adam:
        restype = None
        argtypes = None

        def __call__(self, *a):
            seen["grad_scale"] = a[12].value if hasattr(a[12], "value") else a[12]
            return 0
    stub.etp_adamw_step_ex = Adam()
    L.lib = lambda: stub
    L.require_device = lambda: None
    L.stream_ptr = lambda: C.c_void_p(0)
    L.ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    planner._declared = pretrain._declared = True
    planner.B200Planner._refresh_cache = _fake_refresh
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=1, num_x_layers=2, hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, pred_head_dropout_prob=0.0)
    model = pretrain.B200PreTraining(cfg, device="cpu").train()
    tr = pretrain.PretrainTrainer(model, world_size=world)
    # stand-in for the backward kernels: a deterministic rank-dependent gradient, written when the loss is formed
    orig = model.forward

    def fwd(batch, task, compute_loss=True):
        loss = orig(batch, task, compute_loss)
        tr.m._direct_grad.copy_(torch.arange(tr.m._direct_grad.numel(), dtype=torch.float32) % 97 * (rank + 1))
        return torch.nan_to_num(loss, nan=0.0, posinf=0.0, neginf=0.0)
    model.forward = fwd
    model.__call__ = fwd
    b = make_traj_batch(cfg, 2, 2, 6, 10, seed=rank_seed(7, rank))

    class M:   # PretrainTrainer calls self.model(batch, task, compute_loss=True)
        bert = model.bert

        def __call__(self, batch, task, compute_loss=True):
            return fwd(batch, task, compute_loss)
    tr.model = M()
    tr.step(b, "sap")
    torch.save({"g": tr.m._direct_grad.clone(), "scale": seen.get("grad_scale")}, f"{out}/p{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_pretrain_trainer_allreduce(tmp_path):
    world = 2
    mp.spawn(_pretrain_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "p0.pt"), torch.load(tmp_path / "p1.pt")
    assert torch.equal(r0["g"], r1["g"])
    n = r0["g"].numel()
    want = torch.arange(n, dtype=torch.float32) % 97 * 3.0       # SUM of the rank patterns (x1 + x2)
    # the stubbed backward kernels add nothing; the loss.backward() of the stub graph may not touch the buffer either
    assert torch.equal(r0["g"], want)
    assert abs(float(r0["scale"]) - 0.5) < 1e-7 and abs(float(r1["scale"]) - 0.5) < 1e-7


# ----------------------------------------------------------------------------------------------------
# PlannerTrainer's data-parallel host logic with the C library stubbed: rank 0's parameters are broadcast at
# construction (ranks built from different seeds end up equal, like under DDP), the update runs bucket by bucket
# (all-reduce, then AdamW over the bucket's trainable runs with grad_scale = 1 / world), frozen groups are skipped.
# ----------------------------------------------------------------------------------------------------
def _planner_trainer_worker(rank, world, port, out):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from etpnav_b200 import lib as L
    from etpnav_b200 import planner
    from tests.test_pretrain_dryrun_cpu import _StubLib, _fake_refresh
    stub = _StubLib()
    calls = []

    class Adam:
        restype = None
        argtypes = None

        def __call__(self, *a):
            v = lambda x: x.value if hasattr(x, "value") else x
            calls.append((v(a[0]), int(v(a[5])), float(v(a[12]))))     # (param pointer, element count, grad_scale)
            return 0
    stub.etp_adamw_step = Adam()
    L.lib = lambda: stub
    L.require_device = lambda: None
    L.stream_ptr = lambda: C.c_void_p(0)
    planner._declared = True
    planner.B200Planner._refresh_cache = _fake_refresh
    torch.manual_seed(100 + rank)                                        # different initial weights per rank
    cfg = PlannerConfig(vocab_size=64, num_l_layers=0, num_x_layers=2, fix_pano_embedding=True)
    m = planner.B200Planner(cfg, device="cpu")
    before = m._flat.clone()
    tr = m.make_trainer(lr=1e-3, world_size=world)
    m._direct_grad[tr.lo:tr.hi] = torch.arange(tr.hi - tr.lo, dtype=torch.float32) % 13 * (rank + 1)
    tr.optimizer_step()
    base = m._flat.data_ptr()
    torch.save({"before": before, "after_bcast": m._flat.clone(), "g": m._direct_grad[tr.lo:tr.hi].clone(),
                "calls": [((p - base) // 4, n, s) for p, n, s in calls], "active": tr.active, "lo": tr.lo, "hi": tr.hi,
                "nav": m.layout.group_ranges["nav"], "drop_base": m._drop_base}, f"{out}/t{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_planner_trainer_update_pipeline(tmp_path):
    world = 2
    mp.spawn(_planner_trainer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "t0.pt"), torch.load(tmp_path / "t1.pt")
    assert not torch.equal(r0["before"], r1["before"])                   # built from different seeds ...
    assert torch.equal(r0["after_bcast"], r1["after_bcast"]) and torch.equal(r0["after_bcast"], r0["before"])   # ... rank 0 wins
    n = r0["hi"] - r0["lo"]
    assert torch.equal(r0["g"], r1["g"]) and torch.equal(r0["g"], torch.arange(n, dtype=torch.float32) % 13 * 3.0)   # SUM
    assert r0["drop_base"] != r1["drop_base"]                            # every rank draws its own dropout masks
    # frozen panorama group: only the nav group is stepped, bucket by bucket, every element exactly once, scale 1 / world
    assert r0["active"] == [tuple(r0["nav"])]
    spans = sorted((a, a + cnt) for a, cnt, _ in r0["calls"])
    assert spans[0][0] == r0["nav"][0] and spans[-1][1] == r0["nav"][1]
    assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
    assert all(abs(s - 0.5) < 1e-7 for _, _, s in r0["calls"]) and len(r0["calls"]) >= 1   # (per bucket on a CUDA device)


def _adamw_ref(p, g, m, v, lr, b1, b2, eps, wd, t):
    """torch.optim.AdamW's update (the arithmetic of adamw_kernel / peer_reduce_adamw_kernel)."""
    p = p * (1.0 - lr * wd)
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    denom = v.sqrt() / (1.0 - b2 ** t) ** 0.5 + eps
    return p - (lr / (1.0 - b1 ** t)) * (m / denom), m, v


def _peer_scheme_worker(rank, world, port, out):
    """The owner-computes scheme of the peer-memory update (csrc/peer.cu), restated with gloo collectives: every rank sums
    the sub-slices it OWNS (planner.peer_partition) from all ranks' gradients in rank order, applies AdamW there with
    owner-local state, and every rank receives the owners' values — against all-reduce + the same AdamW everywhere."""
    from etpnav_b200.planner import peer_partition
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, runs = 64 * 40, [(0, 64 * 11), (64 * 12, 64 * 30), (64 * 30, 64 * 31)]       # a frozen gap and a one-granule run
    torch.manual_seed(7)
    p0 = torch.randn(n)
    g = torch.randn(n, generator=torch.Generator().manual_seed(100 + rank))
    hyper = dict(lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=1e-2)
    # replicated path: all-reduce, 1/world, AdamW on every rank (two steps)
    pr, mr, vr = p0.clone(), torch.zeros(n), torch.zeros(n)
    # owner path
    po, mo, vo = p0.clone(), torch.zeros(n), torch.zeros(n)      # mo / vo are only ever touched on owned sub-slices
    owned_total = 0
    for t in (1, 2):
        gs = g * t
        red = gs.clone()
        dist.all_reduce(red)
        for x, y in runs:
            pr[x:y], mr[x:y], vr[x:y] = _adamw_ref(pr[x:y], red[x:y] / world, mr[x:y], vr[x:y], t=t, **hyper)
        every = [torch.empty(n) for _ in range(world)]
        dist.all_gather(every, gs)                               # "peer loads": what the owner reads from each rank
        new = po.clone()
        for x, y in runs:
            a, b = peer_partition(x, y, world, rank)
            if b > a:
                acc = every[0][a:b].clone()
                for r in range(1, world):
                    acc += every[r][a:b]
                new[a:b], mo[a:b], vo[a:b] = _adamw_ref(po[a:b], acc / world, mo[a:b], vo[a:b], t=t, **hyper)
                owned_total += (b - a) if t == 1 else 0
        allp = [torch.empty(n) for _ in range(world)]
        dist.all_gather(allp, new)                               # "peer stores": every rank ends up with the owners' values
        for x, y in runs:
            for r in range(world):
                a, b = peer_partition(x, y, world, r)
                po[a:b] = allp[r][a:b]
    torch.save({"pr": pr, "po": po, "p0": p0, "owned": owned_total, "runs": runs}, f"{out}/r{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def test_owner_computes_update_equals_allreduce_plus_replicated_adamw(tmp_path):
    world = 2
    mp.spawn(_peer_scheme_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["po"], r1["po"])                                  # replicas end bit-identical
    assert torch.allclose(r0["po"], r0["pr"], rtol=0, atol=1e-6)            # same update as the replicated path
    assert r0["owned"] + r1["owned"] == sum(y - x for x, y in r0["runs"])   # the sub-slices partition the trainable runs
    frozen = torch.ones(64 * 40, dtype=torch.bool)
    for x, y in r0["runs"]:
        frozen[x:y] = False
    assert torch.equal(r0["po"][frozen], r0["p0"][frozen])                  # nothing outside the runs moved
