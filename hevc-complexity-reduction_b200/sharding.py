"""Frame sharding + gather for the one-process-per-GPU launch (torchrun / torch.distributed).

The path shards on independent units: every frame is independent and the per-sub-batch gates never
cross a frame (video_to_cu_depth.py:64-70, net_CNN.py:175,187), so rank r takes a contiguous frame
range and the only exchange is the gather of the per-rank cu_depth rows to rank 0, which serialises
them (video_to_cu_depth.py:114-116).  Two transports:

* ``gather_rows``  -- a collective after the kernels: grouped send/recv (torch.distributed.gather) over NCCL, or
  over gloo on CPU tensors in the tests;
* ``PeerGather``   -- no collective on the data path: rank dst exports ONE buffer for the whole sequence
  (include/ethcnn.h, ethcnn_peer_buffer_*), every rank maps it over NVLink / NVSwitch and hands
  ``buffer + first_row_of_rank * 84`` to the kernels as their output pointer, so the gate kernel's coalesced
  stores are the transfer.  torch.distributed only carries the 64-byte handle and the closing barrier.
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


class _DevicePointerView(object):
    """__cuda_array_interface__ wrapper so torch can view a raw device pointer without copying."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerGather(object):
    """The gather buffer of one sharded prediction, living on rank `dst` and written in place by every rank.

    `net` is this rank's EthCnn (or anything with peer_buffer_create / peer_buffer_open / peer_buffer_release).
    After construction `ok` says whether EVERY rank could map the buffer (agreed with a MIN all-reduce); if not,
    nothing stays mapped and the caller uses gather_rows().  Usage on every rank:

        pg = PeerGather(net, total_rows)
        net.set_option(OPT_STAGED_OUTPUT, 1)
        net.predict_luma_device(..., d_out=pg.row_ptr(first_row_of_this_rank), stream=...)
        pg.complete()                       # stream sync + barrier: all rows are now in rank dst's memory
        rows = pg.rows()                    # rank dst: zero-copy torch view [total_rows, row_width]; None elsewhere
    """

    def __init__(self, net, total_rows: int, row_width: int = 21, dst: int = 0, group=None, device=None):
        import torch
        import torch.distributed as dist

        self.net, self.total_rows, self.row_width, self.dst, self.group = net, int(total_rows), int(row_width), dst, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        self.ptr = 0
        self.error = None
        n_bytes = max(4, self.total_rows * self.row_width * 4)
        box = [None]
        if self.rank == dst:
            try:
                self.ptr, box[0] = net.peer_buffer_create(n_bytes)
            except Exception as e:  # no device / allocation failure: every rank learns about it below
                self.error = str(e)
        dist.broadcast_object_list(box, src=dst, group=group)
        mine = 1
        if self.rank != dst:
            if box[0] is None:
                mine = 0
            else:
                try:
                    self.ptr = net.peer_buffer_open(box[0])
                except Exception as e:
                    self.error, mine = str(e), 0
        elif box[0] is None:
            mine = 0
        flag = torch.tensor([mine], dtype=torch.int32, device=device if device is not None else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(int(flag.item()) == 1)
        if not self.ok:
            self.close()

    def row_ptr(self, first_row: int) -> int:
        """Address of row `first_row` (frame-major CTU raster order) inside the gather buffer."""
        if not self.ok:
            raise RuntimeError("peer gather buffer is not mapped on every rank: %s" % (self.error or "a peer failed"))
        if not 0 <= first_row <= self.total_rows:
            raise ValueError("row outside the gather buffer")
        return self.ptr + int(first_row) * self.row_width * 4

    def complete(self) -> None:
        """Every rank's stores have landed on rank dst: local device sync, then a barrier."""
        import torch
        import torch.distributed as dist

        if self.device is not None and torch.cuda.is_available():
            torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def rows(self):
        """Rank dst: the gathered [total_rows, row_width] float32 rows as a zero-copy torch tensor on its device."""
        import torch

        if self.rank != self.dst or not self.ok:
            return None
        return torch.as_tensor(_DevicePointerView(self.ptr, (self.total_rows, self.row_width)), device=self.device)

    def close(self) -> None:
        """Collective: every rank calls it.  Importers unmap first, the exporter frees after a barrier."""
        import torch.distributed as dist

        def release():
            if self.ptr:
                try:
                    self.net.peer_buffer_release(self.ptr)
                except Exception:
                    pass
                self.ptr = 0

        if self.rank != self.dst:
            release()
        dist.barrier(group=self.group)
        if self.rank == self.dst:
            release()
        self.ok = False


class SharedHostRows(object):
    """The HOST-side twin of PeerGather for the host-pointer API (ethcnn_predict_luma): one POSIX shared-memory block
    holds the whole sequence's rows, every rank maps it, page-locks it (cudaHostRegister, so the library's D2H copies go
    straight into it) and passes `ptr + first_row_of_rank * 84` as the `out` pointer.  Rank dst then owns the complete
    cu_depth rows in host memory without any gather: what video_to_cu_depth.py:114-116 writes out.  torch.distributed
    only carries the block's name and the closing barrier."""

    def __init__(self, total_rows: int, row_width: int = 21, dst: int = 0, group=None, register: bool = True):
        import ctypes
        from multiprocessing import shared_memory

        import torch.distributed as dist

        self.total_rows, self.row_width, self.dst, self.group = int(total_rows), int(row_width), dst, group
        self.rank = dist.get_rank(group)
        self.n_bytes = max(4, self.total_rows * self.row_width * 4)
        box = [None]
        if self.rank == dst:
            self.shm = shared_memory.SharedMemory(create=True, size=self.n_bytes)
            box[0] = self.shm.name
        dist.broadcast_object_list(box, src=dst, group=group)
        if self.rank != dst:
            self.shm = shared_memory.SharedMemory(name=box[0])
            try:   # only the creator owns (and unlinks) the block; keep this process's resource tracker out of it
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self._view = (ctypes.c_char * self.n_bytes).from_buffer(self.shm.buf)
        self.ptr = ctypes.addressof(self._view)
        self.registered = False
        if register:
            import torch

            if torch.cuda.is_available():
                rc = torch.cuda.cudart().cudaHostRegister(self.ptr, self.n_bytes, 0)
                self.registered = int(rc) == 0
        dist.barrier(group=group)

    def row_ptr(self, first_row: int) -> int:
        if not 0 <= first_row <= self.total_rows:
            raise ValueError("row outside the shared block")
        return self.ptr + int(first_row) * self.row_width * 4

    def rows(self):
        """Rank dst: numpy view [total_rows, row_width] of the block (valid after every rank's call returned and a barrier)."""
        import numpy as np

        if self.rank != self.dst:
            return None
        return np.frombuffer(self.shm.buf, dtype="<f4", count=self.total_rows * self.row_width).reshape(-1, self.row_width)

    def close(self) -> None:
        """Collective: every rank calls it."""
        import torch.distributed as dist

        if self.registered:
            import torch

            torch.cuda.cudart().cudaHostUnregister(self.ptr)
            self.registered = False
        dist.barrier(group=self.group)
        self._view = None
        try:
            self.shm.close()
        except BufferError:
            pass   # a numpy view handed out by rows() is still alive; the mapping goes with the process
        if self.rank == self.dst:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass


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
