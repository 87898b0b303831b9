"""Frame sharding across the GPUs of one box (one process per GPU, torch.distributed).

Frames are fully independent (fresh model/optimiser per frame, train.py:74-89), so the sequence is split by
frame with no data-path collective; the only exchange is ONE all_gather of the final poses [F_local,J,3]
(204 B/frame at J=17) when a shard finishes.  Backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_frames, rank, world_size):
    """Contiguous, balanced split: the first (n_frames % world_size) ranks get one extra frame."""
    base, extra = divmod(n_frames, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_size_padded(n_frames, world_size):
    return (n_frames + world_size - 1) // world_size


def gather_poses(local_xyz, n_frames, group=None):
    """all_gather of the shards' final poses into [n_frames, J, 3] on every rank.  Shards are padded to equal
    length for the collective (all_gather_into_tensor needs equal sizes) and the padding is dropped afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_xyz[:n_frames]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    pad = shard_size_padded(n_frames, world)
    s, e = shard_bounds(n_frames, rank, world)
    buf = torch.zeros((pad,) + tuple(local_xyz.shape[1:]), dtype=local_xyz.dtype, device=local_xyz.device)
    buf[:e - s] = local_xyz[:e - s]
    out = torch.empty((world * pad,) + tuple(local_xyz.shape[1:]), dtype=local_xyz.dtype, device=local_xyz.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = []
    for r in range(world):
        rs, re = shard_bounds(n_frames, r, world)
        parts.append(out[r * pad:r * pad + (re - rs)])
    return torch.cat(parts, 0)


def optimize_sequence_sharded(seq, device, optimize_fn=None, group=None):
    """Each rank optimises its contiguous shard of ``seq.frames`` and every rank receives all final poses.
    ``optimize_fn(sub_sequence, device) -> [F_local,J,3]`` defaults to the fused CUDA optimiser."""
    from .synthetic import Sequence
    if optimize_fn is None:
        from .trainer import optimize_sequence as optimize_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(seq.frames)
    s, e = shard_bounds(n, rank, world)
    sub = Sequence(cfg=seq.cfg, cameras=seq.cameras, frames=seq.frames[s:e])
    J = seq.cfg.n_joints
    local = optimize_fn(sub, device) if e > s else np.zeros((0, J, 3), np.float32)
    local = torch.as_tensor(np.asarray(local, np.float32)).to(device).reshape(-1, J, 3)
    return gather_poses(local, n, group)
