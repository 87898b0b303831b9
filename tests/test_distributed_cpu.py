"""The N>1 path on CPU: world_size-2 gloo processes shard a sequence by frame and all_gather the final poses."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from skelsplat_b200 import configs, synthetic


def _fake_optimize(sub, device):
    # stand-in for the CUDA optimiser (no GPU here): a deterministic function of each frame's input
    return np.stack([f.pose_3d_init.astype(np.float32) * 2.0 + 1.0 for f in sub.frames]) if sub.frames else np.zeros((0, sub.cfg.n_joints, 3), np.float32)


def _worker(rank, world, n_frames, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from skelsplat_b200.distributed import optimize_sequence_sharded, shard_bounds
    seq = synthetic.make_sequence(configs.H36M, n_frames, seed=4)
    out = optimize_sequence_sharded(seq, "cpu", optimize_fn=_fake_optimize)
    q.put((rank, out.numpy(), shard_bounds(n_frames, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def _run(n_frames, world=2, port=29731):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, n_frames, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_ranks_shard_and_gather_even_and_ragged():
    for n_frames, port in ((6, 29731), (5, 29732), (1, 29733)):
        res = _run(n_frames, port=port)
        seq = synthetic.make_sequence(configs.H36M, n_frames, seed=4)
        want = np.stack([f.pose_3d_init.astype(np.float32) * 2.0 + 1.0 for f in seq.frames])
        bounds = sorted(b for _, _, b in res)
        assert bounds[0][0] == 0 and bounds[-1][1] == n_frames and bounds[0][1] == bounds[1][0]
        for rank, out, _ in res:
            assert out.shape == want.shape and np.array_equal(out, want)       # every rank holds all final poses
