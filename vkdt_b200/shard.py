"""frame-parallel sharding of raw sequences over the GPUs of one box.

The default darkroom graphs have no `feedback` connector, so frames (and stills) are independent units:
frame f goes to rank f mod N, every rank owns one graph on its own GPU, there is no data-path collective;
the host gathers per-frame results in frame order.  torch.distributed is plumbing only (barrier, gather of
small per-frame records, max-over-ranks timing)."""
import torch.distributed as dist


def frames_for_rank(n_frames, rank, world):
    """round robin keeps the output order interleaved and the load balanced to within one frame."""
    return list(range(rank, n_frames, world))


def graph_is_frame_parallel(cfg_text):
    """graphs with feedback connectors carry state from frame to frame (align, accum, svgf ...): not shardable."""
    return not any(l.startswith("feedback:") for l in cfg_text.splitlines())


def gather_in_frame_order(local, n_frames, rank=None, world=None, dst=0, group=None):
    """local: {frame: small record}.  returns the list of records in frame order on rank `dst`, None elsewhere.
    group: a process group for the (host side) gather, e.g. a gloo group next to an nccl default group."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return [local[f] for f in range(n_frames)]
    parts = [None] * world if rank == dst else None
    dist.gather_object(local, parts, dst=dst, group=group)
    if rank != dst:
        return None
    merged = {}
    for p in parts:
        merged.update(p)
    missing = [f for f in range(n_frames) if f not in merged]
    if missing:
        raise RuntimeError("frames missing after gather: %s" % missing[:8])
    return [merged[f] for f in range(n_frames)]


def max_over_ranks(value, device=None):
    """device timings are reported as the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
