"""One process per GPU (the torchrun form) on a box with >= 2 devices: rank r predicts its frame
range and the rows reach rank 0 (a) through the peer-memory gather buffer -- the gate kernel's stores cross NVLink, no
collective on the data path -- and (b) through the NCCL gather.  Both must equal the single-GPU rows bit for bit; a staged
single-GPU run (the same export kernel, local destination) is checked in test_staged_output_is_bit_identical."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import ethcnn_oracle as eo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, QP = 768, 512, 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _clip(n_frames):
    return np.stack([eo.synth_frame(W, H, 300 + k) for k in range(n_frames)])


def _worker(rank, world, port, model_dir, n_frames, out_dir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import ethcnn_b200 as eb

    rows_per_frame = 12 * 8
    net = eb.EthCnn(model_dir, None, eb.MODE_AI, device=rank)
    f0, nf = eb.sharding.frame_range(n_frames, world, rank)
    luma = torch.from_numpy(_clip(n_frames)[f0:f0 + nf].copy()).to(dev)
    stream = torch.cuda.current_stream().cuda_stream

    # (a) peer-memory gather: the kernels write straight into rank 0's buffer
    pg = eb.sharding.PeerGather(net, n_frames * rows_per_frame, 21, dst=0, device=dev)
    assert pg.ok, pg.error
    net.set_option(eb.OPT_STAGED_OUTPUT, 1)
    for _ in range(2):   # the second pass overwrites the same rows (what bench.py does every step)
        if nf:
            net.predict_luma_device(luma.data_ptr(), W, H, W, W * H, nf, QP, pg.row_ptr(f0 * rows_per_frame), stream)
    pg.complete()
    if rank == 0:
        np.save(os.path.join(out_dir, "peer.npy"), pg.rows().cpu().numpy())
    pg.close()

    # (a') host-pointer API: every rank's D2H lands in one shared, page-locked host block
    hr = eb.sharding.SharedHostRows(n_frames * rows_per_frame, 21, dst=0)
    assert hr.registered
    if nf:
        host_luma = torch.from_numpy(_clip(n_frames)[f0:f0 + nf].copy()).pin_memory()
        net.predict_luma_ptr(host_luma.data_ptr(), W, H, W * H, nf, QP, hr.row_ptr(f0 * rows_per_frame))
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "host.npy"), hr.rows().copy())
    hr.close()

    # (b) the collective
    net.set_option(eb.OPT_STAGED_OUTPUT, 0)
    local = torch.empty((nf * rows_per_frame, 21), dtype=torch.float32, device=dev)
    if nf:
        net.predict_luma_device(luma.data_ptr(), W, H, W, W * H, nf, QP, local.data_ptr(), stream)
    full = eb.sharding.gather_rows(local, n_frames, rows_per_frame, 21, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "nccl.npy"), full.cpu().numpy())
    net.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 1])
def test_peer_gather_and_nccl_gather_equal_single_gpu(eb, ai_model_dir, tmp_path, n_frames):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d, _ = ai_model_dir
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), d, n_frames, str(tmp_path)), nprocs=world, join=True)
    with eb.EthCnn(d, None, eb.MODE_AI, device=0) as net:
        want = net.predict_luma(_clip(n_frames), W, H, n_frames, QP)
    assert np.array_equal(np.load(tmp_path / "peer.npy"), want)
    assert np.array_equal(np.load(tmp_path / "nccl.npy"), want)
    assert np.array_equal(np.load(tmp_path / "host.npy"), want)


def test_staged_output_is_bit_identical(eb, ai_model_dir):
    """ETHCNN_OPT_STAGED_OUTPUT on one GPU: dense kernel -> local staging -> gate-and-export kernel gives the same bytes
    as gating in place, including closed gates (a flat frame) and a non-multiple-of-4 row offset."""
    import torch

    d, _ = ai_model_dir
    frames = _clip(3)
    frames[1] = 128                                   # flat frame: every gate closes
    dev = torch.device("cuda", 0)
    luma = torch.from_numpy(frames).to(dev)
    n = 3 * 96
    for fc_path in (3, 2, 1, 0):   # every dense path feeds the export kernel the same way
        with eb.EthCnn(d, None, eb.MODE_AI, device=0) as net:
            net.set_option(eb.OPT_FC1_PATH, fc_path)
            a = torch.full((n + 1, 21), -1.0, dtype=torch.float32, device=dev)
            b = torch.full((n + 1, 21), -1.0, dtype=torch.float32, device=dev)
            s = torch.cuda.current_stream().cuda_stream
            net.predict_luma_device(luma.data_ptr(), W, H, W, W * H, 3, QP, a.data_ptr() + 84, s)
            net.set_option(eb.OPT_STAGED_OUTPUT, 1)
            net.predict_luma_device(luma.data_ptr(), W, H, W, W * H, 3, QP, b.data_ptr() + 84, s)
            torch.cuda.synchronize()
            a, b = a.cpu().numpy(), b.cpu().numpy()
        assert np.array_equal(a, b), "fc path %d" % fc_path
        assert (a[0] == -1).all() and (a[97:193, 1:] == 0).all() and (a[1:97, 0] > 0).all()
