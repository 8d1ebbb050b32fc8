"""The N > 1 path on CPU: world_size-2 (and 3) gloo process groups run the product's sharding + gather
code (hevc-complexity-reduction_b200/sharding.py); the per-rank compute is the oracle here (no GPU), so
what is tested is the partitioning, ordering and gather -- gathered rows must equal the 1-way result
byte for byte."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ethcnn_b200 as eb
    from oracle import ethcnn_oracle as eo

    torch.set_num_threads(1)
    W, H = 200, 136
    weights = eo.random_weights(3)
    yuv = eo.synth_yuv(W, H, n_frames, seed0=50)
    fb = W * H * 3 // 2

    def predict_frames(f0, nf):  # stand-in for EthCnn.predict_luma on this rank's frame range
        if nf == 0:
            return torch.zeros((0, 21), dtype=torch.float32)
        return torch.from_numpy(eo.get_prob(yuv[f0 * fb:(f0 + nf) * fb], W, H, 32, weights, eo.MODE_AI, (0.5, 0.5)))

    full = eb.sharding.predict_sharded(predict_frames, n_frames, rows_per_frame=12, row_width=21, dst=0)
    if rank == 0:
        eb.sharding.write_cu_depth(out_path, full.numpy())
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 5), (3, 2), (2, 1)])
def test_sharded_gather_equals_single_process(tmp_path, world, n_frames):
    from oracle import ethcnn_oracle as eo

    out = str(tmp_path / "cu_depth.dat")
    mp.spawn(_worker, args=(world, _free_port(), n_frames, out), nprocs=world, join=True)
    W, H = 200, 136
    want = eo.get_prob(eo.synth_yuv(W, H, n_frames, seed0=50), W, H, 32, eo.random_weights(3), eo.MODE_AI, (0.5, 0.5))
    got = np.fromfile(out, dtype="<f4").reshape(-1, 21)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert os.path.getsize(out) == n_frames * 12 * 21 * 4


class _ShmNet(object):
    """Stand-in for EthCnn's peer-buffer calls on a box without GPUs: POSIX shared memory plays the NVLink-mapped
    buffer, so the handle exchange, the row offsets and the completion protocol of PeerGather run for real."""

    def __init__(self, fail_open=False):
        self.fail_open, self.shm = fail_open, {}

    def _addr(self, shm):
        import ctypes
        return ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))

    def peer_buffer_create(self, n_bytes):
        from multiprocessing import shared_memory
        shm = shared_memory.SharedMemory(create=True, size=n_bytes)
        p = self._addr(shm)
        self.shm[p] = (shm, True)
        return p, shm.name.encode().ljust(64, b"\0")

    def peer_buffer_open(self, handle):
        from multiprocessing import shared_memory
        if self.fail_open:
            raise RuntimeError("peer access not available")
        shm = shared_memory.SharedMemory(name=handle.rstrip(b"\0").decode())
        p = self._addr(shm)
        self.shm[p] = (shm, False)
        return p

    def peer_buffer_release(self, p):
        shm, owner = self.shm.pop(p)
        # the exported ctypes view pins the mapping; dropping the object is enough for the test
        if owner:
            try:
                shm.unlink()
            except FileNotFoundError:
                pass


def _peer_worker(rank, world, port, n_frames, out_path, fail_rank):
    import ctypes
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ethcnn_b200 as eb
    from oracle import ethcnn_oracle as eo

    torch.set_num_threads(1)
    W, H, rows_per_frame = 200, 136, 12
    weights = eo.random_weights(3)
    yuv = eo.synth_yuv(W, H, n_frames, seed0=50)
    fb = W * H * 3 // 2
    f0, nf = eb.sharding.frame_range(n_frames, world, rank)
    local = eo.get_prob(yuv[f0 * fb:(f0 + nf) * fb], W, H, 32, weights, eo.MODE_AI, (0.5, 0.5)) if nf else np.zeros((0, 21), np.float32)
    net = _ShmNet(fail_open=(rank == fail_rank))
    pg = eb.sharding.PeerGather(net, n_frames * rows_per_frame, 21, dst=0)
    if fail_rank >= 0:
        # one rank could not map the buffer: EVERY rank must see ok == False and fall back to the collective
        assert not pg.ok and pg.ptr == 0
        full = eb.sharding.gather_rows(torch.from_numpy(local), n_frames, rows_per_frame, 21, dst=0)
        if rank == 0:
            eb.sharding.write_cu_depth(out_path, full.numpy())
    else:
        assert pg.ok
        local = np.ascontiguousarray(local, dtype="<f4")
        ctypes.memmove(pg.row_ptr(f0 * rows_per_frame), local.ctypes.data, local.nbytes)   # "the kernels store their rows"
        pg.complete()
        if rank == 0:
            buf = (ctypes.c_float * (n_frames * rows_per_frame * 21)).from_address(pg.row_ptr(0))
            eb.sharding.write_cu_depth(out_path, np.frombuffer(buf, dtype="<f4").reshape(-1, 21).copy())
        with pytest.raises(ValueError):
            pg.row_ptr(n_frames * rows_per_frame + 1)
        pg.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames,fail_rank", [(2, 5, -1), (3, 4, -1), (2, 3, 1)])
def test_peer_gather_equals_single_process(tmp_path, world, n_frames, fail_rank):
    from oracle import ethcnn_oracle as eo

    out = str(tmp_path / "cu_depth.dat")
    mp.spawn(_peer_worker, args=(world, _free_port(), n_frames, out, fail_rank), nprocs=world, join=True)
    W, H = 200, 136
    want = eo.get_prob(eo.synth_yuv(W, H, n_frames, seed0=50), W, H, 32, eo.random_weights(3), eo.MODE_AI, (0.5, 0.5))
    got = np.fromfile(out, dtype="<f4").reshape(-1, 21)
    assert got.shape == want.shape and np.array_equal(got, want)


def _host_rows_worker(rank, world, port, n_frames, out_path):
    import ctypes
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ethcnn_b200 as eb
    from oracle import ethcnn_oracle as eo

    torch.set_num_threads(1)
    W, H, rows_per_frame = 200, 136, 12
    yuv = eo.synth_yuv(W, H, n_frames, seed0=50)
    fb = W * H * 3 // 2
    f0, nf = eb.sharding.frame_range(n_frames, world, rank)
    hr = eb.sharding.SharedHostRows(n_frames * rows_per_frame, 21, dst=0, register=False)
    if nf:   # "ethcnn_predict_luma(..., out = hr.row_ptr(first row of this rank))"
        local = np.ascontiguousarray(eo.get_prob(yuv[f0 * fb:(f0 + nf) * fb], W, H, 32, eo.random_weights(3), eo.MODE_AI, (0.5, 0.5)), "<f4")
        ctypes.memmove(hr.row_ptr(f0 * rows_per_frame), local.ctypes.data, local.nbytes)
    dist.barrier()
    if rank == 0:
        eb.sharding.write_cu_depth(out_path, hr.rows().copy())
    else:
        assert hr.rows() is None
    hr.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 5), (3, 2)])
def test_shared_host_rows_equal_single_process(tmp_path, world, n_frames):
    from oracle import ethcnn_oracle as eo

    out = str(tmp_path / "cu_depth.dat")
    mp.spawn(_host_rows_worker, args=(world, _free_port(), n_frames, out), nprocs=world, join=True)
    W, H = 200, 136
    want = eo.get_prob(eo.synth_yuv(W, H, n_frames, seed0=50), W, H, 32, eo.random_weights(3), eo.MODE_AI, (0.5, 0.5))
    got = np.fromfile(out, dtype="<f4").reshape(-1, 21)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_frame_ranges(eb):
    fr = eb.sharding.all_frame_ranges
    assert [n for _, n in fr(50, 8)] == [7, 7, 6, 6, 6, 6, 6, 6]          # BASELINE config 3
    assert [n for _, n in fr(425, 4)] == [107, 106, 106, 106]              # BASELINE config 4
    for n, w in ((0, 4), (1, 8), (7, 3), (50, 1)):
        r = fr(n, w)
        assert sum(k for _, k in r) == n
        assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(w - 1)) and r[0][0] == 0
