"""Data-parallel plumbing for the planner step: the batch axis is the only sharded axis, weights
are replicated, and the one collective is a SUM all-reduce of the flat gradient slice once per optimizer step
(the reference: DDP bucketed all-reduce, ss_trainer_ETP.py:211-212).  The mean over ranks is applied as
``grad_scale = 1/world`` inside the fused AdamW kernel, so no extra pass over the buffer is needed."""
import torch


def rank_seed(base: int, rank: int) -> int:
    """Each rank draws its own synthetic shard (mirrors per-rank environments, ss_trainer_ETP.py:159-163)."""
    return base + rank


def shard_batch(global_batch: int, world: int, rank: int):
    """Contiguous [start, end) rows of the global batch owned by ``rank`` (even split, remainder to low ranks)."""
    per, rem = divmod(global_batch, world)
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)


def allreduce_flat_(flat_grad: torch.Tensor, world: int) -> float:
    """In-place SUM all-reduce of a flat gradient slice; returns the scale that turns the sum into DDP's mean."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world


def gradient_buckets(layout, cfg, lo, hi):
    """Slices [start, end) of the flat gradient buffer in the order the backward completes them, covering [lo, hi):
    x-layer X-1 (with the SAP head that follows it in the layout), X-2, ..., 0, then the rest (panorama group,
    node-packing parameters and the stacked text key|value projections, all finished last).  Each x-layer's own
    parameters are one contiguous run of the layout (etpnav_b200/layout.py).  When the slice also holds the panorama
    group, the remainder is split into ``nav_head`` (final at the end of the navigation backward), ``pano_layer_i`` for
    the panorama layers above 0 (the last one carries pano_encoder.norm) and ``rest``."""
    X = cfg.num_x_layers
    buckets = []
    if X == 0:
        return [("rest", lo, hi)]
    starts = [layout.offset(f"global_encoder.encoder.x_layers.{i}.visual_attention.att.query.weight") for i in range(X)]
    ends = starts[1:] + [hi]     # the last layer's run extends over the SAP head to the end of the nav group
    for i in range(X - 1, -1, -1):
        buckets.append((f"x_layer_{i}", starts[i], ends[i]))
    nav_lo = layout.group_ranges["nav"][0]
    if lo < nav_lo < starts[0]:
        buckets.append(("nav_head", nav_lo, starts[0]))   # node packing + stacked text K|V: done when the nav backward is
        # panorama group, finished last and in the order layer P-1 ... 0: every layer above 0 is its own bucket (final at
        # the event etp_backward_panorama records for it), layer 0 + the view embeddings are the exposed remainder
        P, end = cfg.num_pano_layers, nav_lo
        for i in range(P - 1, 0, -1):
            st = layout.offset(f"img_embeddings.pano_encoder.layers.{i}.self_attn.in_proj_weight")
            if lo < st < end:
                buckets.append((f"pano_layer_{i}", st, end))
                end = st
        buckets.append(("rest", lo, end))
    elif starts[0] > lo:
        buckets.append(("rest", lo, starts[0]))
    return buckets


def allreduce_buckets_(flat_grad, buckets, world, side_stream=None, wait_fns=None):
    """SUM all-reduce of ``flat_grad`` bucket by bucket.  With a CUDA ``side_stream``, bucket i is enqueued on it
    after ``wait_fns[i](side_stream)`` (e.g. a cudaStreamWaitEvent on the event the backward records when that
    bucket is final), so the collectives overlap the rest of the backward; the caller joins the streams afterwards.
    Without a side stream (CPU / gloo tests) the buckets are reduced in order on the current stream."""
    if world <= 1:
        return 1.0
    import torch.distributed as dist
    if side_stream is None:
        for _, a, b in buckets:
            dist.all_reduce(flat_grad[a:b], op=dist.ReduceOp.SUM)
        return 1.0 / world
    for i, (_, a, b) in enumerate(buckets):
        if wait_fns is not None and wait_fns[i] is not None:
            wait_fns[i](side_stream)
        with torch.cuda.stream(side_stream):
            dist.all_reduce(flat_grad[a:b], op=dist.ReduceOp.SUM)
    return 1.0 / world
