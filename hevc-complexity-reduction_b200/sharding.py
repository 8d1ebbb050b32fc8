"""Frame sharding + gather for the one-process-per-GPU launch (torchrun / torch.distributed).

The path shards on independent units: every frame is independent and the per-sub-batch gates never
cross a frame (video_to_cu_depth.py:64-70, net_CNN.py:175,187), so rank r takes a contiguous frame
range and the only exchange is the gather of the per-rank cu_depth rows to rank 0, which serialises
them (video_to_cu_depth.py:114-116).  Over NCCL this is a grouped send/recv (torch.distributed.gather);
the same code runs over gloo on CPU tensors in the tests.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np


def frame_range(n_frames: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first_frame, n_frames_of_rank): contiguous ranges, the first n_frames % world ranks take one
    extra frame (50 frames over 8 ranks -> 7,7,6,6,6,6,6,6).  Same rule as csrc/engine.cu frame_range()."""
    base, rem = divmod(n_frames, world_size)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def all_frame_ranges(n_frames: int, world_size: int) -> List[Tuple[int, int]]:
    return [frame_range(n_frames, world_size, r) for r in range(world_size)]


def gather_rows(local_rows, n_frames: int, rows_per_frame: int, row_width: int, dst: int = 0, group=None):
    """Gather each rank's [frames_r * rows_per_frame, row_width] float32 tensor to rank `dst`, in frame
    order.  `local_rows` is a torch tensor on the rank's device (cuda for nccl, cpu for gloo).  Returns
    the full [n_frames * rows_per_frame, row_width] tensor on rank dst, None elsewhere.  Shards are
    padded to the largest shard so one gather call moves everything."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = all_frame_ranges(n_frames, world)
    max_rows = max(nf for _, nf in ranges) * rows_per_frame
    send = torch.zeros((max_rows, row_width), dtype=torch.float32, device=local_rows.device)
    if local_rows.numel():
        send[: local_rows.shape[0]].copy_(local_rows.reshape(-1, row_width))
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([recv[r][: ranges[r][1] * rows_per_frame] for r in range(world)], dim=0)


def predict_sharded(predict_frames: Callable[[int, int], "object"], n_frames: int, rows_per_frame: int,
                    row_width: int = 21, dst: int = 0, group=None):
    """Run `predict_frames(first_frame, n)` (returns this rank's rows as a torch tensor) on this rank's
    frame range and gather to rank dst."""
    import torch.distributed as dist

    f0, nf = frame_range(n_frames, dist.get_world_size(group), dist.get_rank(group))
    local = predict_frames(f0, nf)
    return gather_rows(local, n_frames, rows_per_frame, row_width, dst, group)


def write_cu_depth(path: str, rows: np.ndarray) -> None:
    """video_to_cu_depth.py:114-116: one raw little-endian float32 write, no header; via a temporary
    file so a failure never leaves a truncated cu_depth.dat."""
    import os

    tmp = "%s.tmp.%d" % (path, os.getpid())
    with open(tmp, "wb") as f:
        f.write(np.ascontiguousarray(rows, dtype="<f4").tobytes())
    os.replace(tmp, path)
