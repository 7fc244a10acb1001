"""Worker of tests/test_dp_peer_gpu.py (launched with torch.distributed.run, one process per GPU): the peer-memory
data-parallel update (PlannerTrainer grad_comm="peer": reduce-scatter + AdamW + parameter all-gather in one kernel over
NVLink, csrc/peer.cu) against the NCCL all-reduce + replicated AdamW path on the SAME steps — every rank its own data
shard and dropout stream, frozen parameters in one group — and the bit-equality of the replicas.  Prints one JSON line on
rank 0."""
import json
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from etpnav_b200.config import PlannerConfig
    from etpnav_b200.planner import B200Planner
    from etpnav_b200.synth import make_inputs, make_weights
    cfg = PlannerConfig(vocab_size=2048, num_l_layers=0, num_x_layers=2)
    sd = make_weights(cfg, seed=1)
    data = [{k: (v.to(dev) if isinstance(v, torch.Tensor) else v)
             for k, v in make_inputs(cfg, 6, 12, 24, 40, seed=50 + 7 * s + rank, ragged=True).items()} for s in range(4)]
    out = {}
    for comm in ("fp32", "peer"):
        m = B200Planner(cfg, device=dev)
        m.load_state_dict(sd, strict=True)
        m.train()
        # a frozen tensor in the middle of a bucket: the trainable runs (and their per-rank sub-slices) must skip it
        m._pmap["global_encoder.encoder.x_layers.1.visn_self_att.output.dense.weight"].requires_grad_(False)
        m.set_dropout_seed(1234 + rank)
        tr = m.make_trainer(lr=1e-3, world_size=world, grad_comm=comm)
        if tr._peer is not None:
            tr._peer_write_reduced = 1        # owners leave the summed gradient in their own buffer (compared below)
        m.set_dropout_seed(1234 + rank)
        p0 = m._flat[tr.lo:tr.hi].clone()
        tr.step(data[0], keep_grads=True)
        torch.cuda.synchronize()
        g1, p1 = m._direct_grad[tr.lo:tr.hi].clone(), m._flat[tr.lo:tr.hi].clone()
        for d in data[1:]:
            tr.step(d, pipelined=True)        # training-loop form: panorama buckets finish under the next step
        tr.join()
        torch.cuda.synchronize()
        p = m._flat[tr.lo:tr.hi].clone()
        pb = m._flat_bf16[tr.lo:tr.hi].clone()
        lo, hi = p.clone(), p.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out[comm] = dict(p=p, g1=g1, p1=p1, ranges=tr.owned_ranges(), moved=float((p - p0).abs().max()),
                         cross=float((hi - lo).abs().max()), image_ok=bool(torch.equal(pb, p.bfloat16())),
                         effective=tr.grad_comm, fallback=tr.peer_fallback, err=tr.peer_error(),
                         owned=sum(b - a for a, b in tr.owned_ranges()), total=tr.hi - tr.lo)
        frozen = m._pmap["global_encoder.encoder.x_layers.1.visn_self_att.output.dense.weight"]
        out[comm]["frozen_same"] = bool(torch.equal(frozen.detach().cpu(), sd["global_encoder.encoder.x_layers.1.visn_self_att.output.dense.weight"]))
        del tr, m
    # step 1: the owner's summed gradient == NCCL's all-reduced gradient up to fp32 summation order; the first AdamW step
    # moves an element by lr * g / (|g| + eps'), so wherever the gradient is above the noise floor both paths write the same value
    gn, gp = out["fp32"]["g1"], out["peer"]["g1"]
    gmax = float(gn.abs().max())
    gerr = max(float((gp[x:y] - gn[x:y]).abs().max()) for x, y in out["peer"]["ranges"] if y > x) / gmax
    big = gn.abs() > 1e-4 * gmax
    p1err = float((out["peer"]["p1"] - out["fp32"]["p1"])[big].abs().max())
    stats = torch.tensor([gerr, p1err], device=dev, dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    d = (out["peer"]["p"] - out["fp32"]["p"]).abs()
    drop = ("p", "g1", "p1", "ranges")
    res = {"world": world, "step1_grad_max_rel": float(stats[0]), "step1_param_max_abs_where_grad_significant": float(stats[1]),
           "frac_grad_significant": float(big.float().mean()), "max_abs_peer_vs_nccl_after_4_steps": float(d.max()),
           "peer": {k: v for k, v in out["peer"].items() if k not in drop},
           "nccl": {k: v for k, v in out["fp32"].items() if k not in drop}}
    owned = torch.tensor([out["peer"]["owned"]], device=dev)
    dist.all_reduce(owned)
    res["owned_total_over_ranks"] = int(owned.item())
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
