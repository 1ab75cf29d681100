"""host side of the N > 1 path on CPU: two gloo ranks shard an MLV clip frame-parallel, each 'develops' its frames with
the CPU oracle (stand-in for its GPU), rank 0 gathers per-frame records in frame order.  no data-path collective."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vkdt_b200 import shard


def _worker(rank, world, port, path, n_frames, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as O
    mine = shard.frames_for_rank(n_frames, rank, world)
    local = {}
    for f in mine:
        raw, info = (O.ref_mlv_decode(path, f) if O.ref_lib() is not None else (None, None))
        if raw is None:  # no reference decoder on this box: the frames are regenerated from their seeds
            from vkdt_b200 import synth
            raw = synth.mosaic(96, 64, seed=100 + f)
        d = O.darkroom_defaults(raw.shape[1], raw.shape[0])
        out = O.darkroom_run(d, raw)
        local[f] = (f, float(out[..., :3].sum()), rank)
    t = shard.max_over_ranks(rank + 1.0)
    res = shard.gather_in_frame_order(local, n_frames)
    if rank == 0:
        ret["frames"] = res
        ret["tmax"] = t
    dist.barrier()
    dist.destroy_process_group()


def test_frames_for_rank():
    assert shard.frames_for_rank(10, 0, 4) == [0, 4, 8] and shard.frames_for_rank(10, 3, 4) == [3, 7]
    got = sorted(f for r in range(8) for f in shard.frames_for_rank(1000, r, 8))
    assert got == list(range(1000))
    assert max(len(shard.frames_for_rank(1000, r, 8)) for r in range(8)) - min(len(shard.frames_for_rank(1000, r, 8)) for r in range(8)) <= 1
    assert shard.frames_for_rank(3, 5, 8) == []


def test_feedback_graphs_are_not_sharded():
    assert shard.graph_is_frame_parallel("module:i-mlv:main\nconnect:a:b:c:d:e:f\n")
    assert not shard.graph_is_frame_parallel("module:align:01\nfeedback:align:01:output:align:01:input\n")


def test_two_rank_frame_parallel(tmp_path):
    from vkdt_b200 import synth
    n_frames = 5
    frames = [synth.mosaic(96, 64, seed=100 + f) for f in range(n_frames)]
    path = str(tmp_path / "clip.mlv")
    synth.write_mlv(path, frames)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, path, n_frames, ret), nprocs=2, join=True)
    res = ret["frames"]
    assert [r[0] for r in res] == list(range(n_frames))          # frame order restored
    assert [r[2] for r in res] == [f % 2 for f in range(n_frames)]  # round robin ownership
    assert ret["tmax"] == 2.0                                     # max over ranks
    # same result as a single process
    from oracle import oracle_py as O
    for f in (0, 3):
        d = O.darkroom_defaults(96, 64)
        assert abs(float(O.darkroom_run(d, frames[f])[..., :3].sum()) - res[f][1]) < 1e-3
