#!/usr/bin/env python
"""bench.py -- CTUs/sec of the ETH-CNN CU-partition predictor (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input.  Workloads (BASELINE.json `configs`):

  config 2 (headline, default)  1920x1080, 50 frames PER RANK per step, QP cycling 22/27/32/37        weak scaling
  config 3                      4928x3264, 50 frames, QP 32, frames split 7,7,6,... over the ranks     strong scaling
  config 4                      2880x1920, 425 frames, QP 27, split 107,106,...; also looped >= 2 s    strong scaling
  config 5                      1920x1080 residue stream, 240 frames, QP 37, LDP residual-CNN weights  strong scaling

The JSON line describes `--config` (default 2); without `--config` the other three are measured briefly as well and
reported under "other_configs", and configs 2 and 4 are additionally looped back to back for >= 2 s ("sustained").

  value     device-resident: luma already in HBM, kernels + gather, CUDA events on the launching stream, max over ranks
  e2e       the public host API (ethcnn_predict_luma through ctypes) from PINNED HOST memory: H2D of the luma planes,
            kernels, D2H of the probabilities (+ gather), every step
  roofline  dominant kernel: algorithmic FLOPs (SURVEY.md section 8(d) per-CTU figures x the CTUs of a launch) / its
            CUDA-event time, against the measured bf16 tensor peak (burst if the timed region is < 1 s, else sustained)
  cpu_baseline   the oracle port of the reference's script on the host cores (rank 0, N = 1, bounded sample)

--impl reference times that CPU port alone (the reference's TensorFlow cannot be installed here), same workload string.
The product arm never imports oracle/: frames and the encoder-like model directory come from tools/synth.py.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_CTU = 4096 + 84
ALG_FLOP_PER_CTU = 3104298              # 2 * 1 552 149 MAC
CONV_FLOP_PER_CTU = 2 * 279552
SCRATCH_BYTES_PER_CTU = 2 * 2 * 2688    # the features as fp16 hi + lo, written by the conv kernel and read by the FC kernel
# dram__bytes_read.sum + dram__bytes_write.sum of one launch, per CTU, from the `ncu --set full` captures under profiles/
NCU_DRAM_BYTES_PER_CTU = {"conv": (104.592384e6 + 223.902464e6) / 25500, "fc1": (289.413376e6 + 5.936640e6) / 25500}
NCU_DRAM_SOURCE = {"conv": "profiles/r02j_conv.md", "fc1": "profiles/r02j_fc_pair.md"}

MODE_AI, MODE_LDP = 0, 1
CONFIGS = {
    2: dict(w=1920, h=1080, frames=50, qps=(22, 27, 32, 37), mode=MODE_AI, scaling="weak", residue=False,
            workload="config2: 1920x1080 4:2:0, 50 frames per rank per step, QP cycling 22/27/32/37"),
    3: dict(w=4928, h=3264, frames=50, qps=(32,), mode=MODE_AI, scaling="strong", residue=False,
            workload="config3: 4928x3264 4:2:0, 50 frames per step sharded in contiguous frame ranges, QP 32"),
    4: dict(w=2880, h=1920, frames=425, qps=(27,), mode=MODE_AI, scaling="strong", residue=False,
            workload="config4: 2880x1920 4:2:0, 425 frames per step sharded in contiguous frame ranges, QP 27"),
    5: dict(w=1920, h=1080, frames=240, qps=(37,), mode=MODE_LDP, scaling="strong", residue=True,
            workload="config5: 1920x1080 residue stream, 240 frames per step, QP 37, LDP residual-CNN weights"),
}


def ctus_per_frame(cfg):
    return -(-cfg["w"] // 64) * -(-cfg["h"] // 64)


def frame_range(n_frames, world, rank):
    base, rem = divmod(n_frames, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def rank_frames(cfg, world, rank):
    """(first frame, frames) this rank holds per step: weak = every rank its own `frames`; strong = a contiguous range."""
    if cfg["scaling"] == "weak":
        return rank * cfg["frames"], cfg["frames"]
    return frame_range(cfg["frames"], world, rank)


def clip_frames(cfg, f0, nf, seed0):
    """Frames [f0, f0 + nf) of the config's global synthetic sequence (deterministic, independent of the sharding)."""
    from tools import synth

    gen = synth.synth_residue_frame if cfg["residue"] else synth.synth_frame
    nb = 5
    base = {}
    out = np.empty((nf, cfg["h"], cfg["w"]), np.uint8)
    for i in range(nf):
        k = f0 + i
        if k % nb not in base:
            base[k % nb] = gen(cfg["w"], cfg["h"], seed0 + k % nb)
        out[i] = np.roll(base[k % nb], shift=(8 * (k // nb), 16 * (k // nb)), axis=(0, 1))
    return out


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while a timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                pw.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.remove(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        order = np.argsort(pw)[len(pw) // 2:]  # the half of the samples with the highest power draw = under load
        return {"sm_mhz": float(np.median(np.array(sm)[order])), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw))}


# --------------------------------------------------------------------------------------------- CPU arm
_POOL_W = {}
_POOL_CFG = {}


def _pool_init(model_dir, mode, qps):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import assets, tf_bundle
    for qp in qps:
        name = assets.LDP_MODEL if mode == MODE_LDP else assets.AI_MODELS[min(assets.AI_MODELS, key=lambda k: abs(k - qp))]
        _POOL_W[qp] = tf_bundle.read_bundle(os.path.join(model_dir, name))
    _POOL_CFG["mode"] = mode


def _pool_frame(args):
    from oracle import ethcnn_oracle as eo
    luma, qp = args
    h, w = luma.shape
    pad = np.zeros((-(-h // 64) * 64, -(-w // 64) * 64), np.uint8)
    pad[:h, :w] = luma
    mode = eo.MODE_LDP if _POOL_CFG["mode"] == MODE_LDP else eo.MODE_AI
    return eo.predict_frame(pad, qp, _POOL_W[qp], mode, (0.5, 0.5))


class CpuReference(object):
    """The oracle port run the reference's way (per frame: pad, slice 64x64 CTUs in raster order, sub-batches of <= 1024
    through the fp32 net, gates), frames farmed out to one worker process per host core.  Checker code: only the
    cpu_baseline leg and --impl reference come here."""

    def __init__(self, model_dir, cfg):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor

        self.cores = os.cpu_count() or 1
        self.pool = ProcessPoolExecutor(max_workers=self.cores, mp_context=mp.get_context("fork"),
                                        initializer=_pool_init, initargs=(model_dir, cfg["mode"], cfg["qps"]))
        list(self.pool.map(_pool_frame, [(np.zeros((64, 64), np.uint8), cfg["qps"][0])] * self.cores))  # start the workers

    def run(self, frames: np.ndarray, qp: int) -> np.ndarray:
        return np.concatenate(list(self.pool.map(_pool_frame, [(f, qp) for f in frames])), axis=0)

    def close(self):
        self.pool.shutdown()


def cpu_model_dir(cfg):
    """Encoder-like directory for the CPU port (checker side: oracle.assets)."""
    from oracle import assets

    d = tempfile.mkdtemp(prefix="ethcnn_ref_")
    assets.materialize(d, "LDP" if cfg["mode"] == MODE_LDP else "AI")
    return d


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config or 2]
    from tools import synth

    d = cpu_model_dir(cfg)
    missing = [qp for qp in cfg["qps"] if cfg["mode"] == MODE_AI and not os.path.exists(os.path.join(d, synth.AI_MODELS[qp] + ".index"))]
    for qp in missing:
        synth.write_bundle(os.path.join(d, synth.AI_MODELS[qp]), synth.random_cnn_weights(100 + qp))
    ref = CpuReference(d, cfg)
    clip = clip_frames(cfg, 0, min(cfg["frames"], max(ref.cores, 8)), 0)
    # bounded sample: size the per-step frame count so that steps + warm-up finish in a couple of minutes
    t = time.time()
    ref.run(clip[:ref.cores], cfg["qps"][0])
    per_frame = (time.time() - t) / min(ref.cores, len(clip))
    budget = 120.0 / max(1, args.steps + args.warmup)
    n = int(max(1, min(len(clip), budget / max(per_frame, 1e-6))))
    nq = len(cfg["qps"])
    for i in range(args.warmup):
        ref.run(clip[:n], cfg["qps"][i % nq])
    t0 = time.time()
    for i in range(args.steps):
        ref.run(clip[:n], cfg["qps"][i % nq])
    dt = time.time() - t0
    ref.close()
    v = args.steps * n * ctus_per_frame(cfg) / dt
    line = {
        "impl": "reference", "metric": "CTUs/sec (ETH-CNN inference)", "value": v, "unit": "CTU/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"]},
        "cpu_baseline": {"value": v, "unit": "CTU/s", "cores": ref.cores, "kind": "port",
                         "sample": "%d frames of %dx%d per step, %d steps; TensorFlow 1.x is not installable here, this is the oracle "
                                   "port of video_to_cu_depth.py / net_CNN.py (numpy fp32, per-frame CTU slicing, sub-batches of "
                                   "1024, gates), one worker process per core" % (n, cfg["w"], cfg["h"], args.steps)},
        "e2e": {"value": v, "unit": "CTU/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------- GPU arm
class Ctx(object):
    pass


def pin_rank_to_cores(local, n_local):
    """Give every rank of the node its own slice of the host cores (the H2D staging and the launch thread stay put)."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = len(cpus) // max(1, n_local)
        if per >= 2:
            os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
            return per
    except Exception:
        pass
    return 0


def measure_config(cx, cfg_id, steps, warmup, want_e2e=True, want_check=True, sustain_s=0.0, h2d_probe=False):
    """All numbers of one config at the current world size.  Returns a dict (rank 0 fills it completely)."""
    import torch
    import torch.distributed as dist

    import ethcnn_b200 as eb

    cfg = CONFIGS[cfg_id]
    world, rank, dev, stream = cx.world, cx.rank, cx.dev, cx.stream
    net = cx.nets[cfg["mode"]]
    W, H, cpf = cfg["w"], cfg["h"], ctus_per_frame(cfg)
    f0, nf = rank_frames(cfg, world, rank)
    total_frames = cfg["frames"] * world if cfg["scaling"] == "weak" else cfg["frames"]
    n_ctus = nf * cpf                 # this rank, per step
    total_ctus = total_frames * cpf   # the job, per step
    qps = cfg["qps"]
    nq = len(qps)
    # one clip per QP for the weak config (4 x 103.7 MB of luma > 126 MB L2, so a step never finds its input in L2);
    # the strong configs hold one sequence that is itself several times the L2
    n_clips = nq if cfg["scaling"] == "weak" else 1
    clips_host, clips_dev = [], []
    for i in range(n_clips):
        seed0 = (1000 * rank + 10 * i) if cfg["scaling"] == "weak" else 7000 * cfg_id
        c = torch.from_numpy(clip_frames(cfg, 0 if cfg["scaling"] == "weak" else f0, nf, seed0))
        c = c.pin_memory() if (want_e2e or h2d_probe) else c
        clips_host.append(c)
        clips_dev.append(c.to(dev))
    out_dev = torch.empty((max(1, n_ctus), 21), dtype=torch.float32, device=dev)
    out_host = torch.empty((max(1, n_ctus), 21), dtype=torch.float32).pin_memory()
    first_row = f0 * cpf

    peer = host_rows = None
    if world > 1 and not cx.args.nccl_gather:
        peer = eb.sharding.PeerGather(net, total_ctus, 21, dst=0, device=dev)
        if not peer.ok:
            if rank == 0:
                print("[bench] peer gather buffer unavailable (%s): NCCL gather instead" % peer.error, file=sys.stderr)
            peer = None
        if want_e2e:
            host_rows = eb.sharding.SharedHostRows(total_ctus, 21, dst=0)
            ok = torch.tensor([1 if host_rows.registered else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                host_rows.close()
                host_rows = None
    peer_out = peer.row_ptr(first_row) if peer is not None else 0
    host_out = host_rows.row_ptr(first_row) if host_rows is not None else 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_nccl():
        return eb.sharding.gather_rows(out_dev[:n_ctus], total_frames, cpf, 21, dst=0)

    def step_device(i):
        k = i % nq
        c = clips_dev[k % n_clips]
        if nf == 0:
            return
        net.predict_luma_device(c.data_ptr(), W, H, W, W * H, nf, qps[k], peer_out if peer is not None else out_dev.data_ptr(),
                                stream.cuda_stream)
        if world > 1 and peer is None:
            gather_nccl()

    def step_e2e(i):
        k = i % nq
        c = clips_host[k % n_clips]
        if nf:
            net.predict_luma_ptr(c.data_ptr(), W, H, W * H, nf, qps[k], host_out if host_rows is not None else out_host.data_ptr())
        if world > 1 and host_rows is None:
            out_dev.copy_(out_host, non_blocking=True)
            g = gather_nccl()
            if rank == 0:
                g.cpu()

    def timed(step_fn, n_steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(n_steps):
            step_fn(i)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    res = {"workload": cfg["workload"], "scaling": cfg["scaling"], "ctus_per_step": total_ctus, "frames_of_rank0": nf if rank == 0 else None}
    net.set_option(eb.OPT_STAGED_OUTPUT, 1 if peer is not None else 0)
    # set-up, not steps: touch every QP range once so that its checkpoint is parsed, packed and resident
    for k in range(nq):
        step_device(k)
    for i in range(warmup):
        step_device(i)
    barrier()
    # `value`: the timed region holds nothing but the library's own launches (no per-stage events between the kernels: the
    # FC and gate kernels are programmatic dependents of their predecessors and start while those drain)
    launches0 = net.kernel_launches
    dev_ms, _ = timed(step_device, steps)
    launches = net.kernel_launches - launches0
    # per-stage CUDA-event times: a second, identical loop with the library's stage timers on
    net.profile_enable(True)
    for s in range(4):
        net.profile_read(s, reset=True)
    prof_ms, _ = timed(step_device, steps)
    stage = {}
    for s, name in enumerate(eb.STAGE_NAMES):
        ms, n = net.profile_read(s, reset=True)
        stage[name] = {"ms_total": ms, "launches": n}
    net.profile_enable(False)
    res["ms_per_step_with_stage_events"] = max_over_ranks(prof_ms) / steps
    dev_ms = max_over_ranks(dev_ms)
    res["value"] = steps * total_ctus / (dev_ms * 1e-3)
    res["ms_per_step"] = dev_ms / steps
    res["timed_region_s"] = dev_ms * 1e-3
    res["gpu_launches"] = int(sum_over_ranks(launches))
    res["stages"] = {name: {"ms_per_step": v["ms_total"] / steps, "launches_per_step": v["launches"] / steps} for name, v in stage.items()}
    res["_stage_raw"] = stage

    # ---- sustained: the same device-resident step looped back to back for >= sustain_s seconds, clocks sampled meanwhile
    if sustain_s > 0:
        n_loop = max(steps, int(np.ceil(sustain_s * 1e3 / max(1e-3, dev_ms / steps))))
        sampler = ClockSampler(cx.local).start()
        sus_ms, _ = timed(step_device, n_loop)
        # stage split in the same thermal state: a short loop with the library's stage timers on, right behind the long one
        net.profile_enable(True)
        for s in range(4):
            net.profile_read(s, reset=True)
        n_prof = max(5, min(50, n_loop // 10))
        timed(step_device, n_prof)
        sstage = {}
        for s, name in enumerate(eb.STAGE_NAMES):
            ms, n = net.profile_read(s, reset=True)
            sstage[name] = ms / n_prof
        net.profile_enable(False)
        clocks = sampler.stop()
        sus_ms = max_over_ranks(sus_ms)
        res["sustained"] = {"value": n_loop * total_ctus / (sus_ms * 1e-3), "unit": "CTU/s", "steps": n_loop, "seconds": sus_ms * 1e-3,
                            "ms_per_step": sus_ms / n_loop, "stage_ms_per_step": sstage, "clocks": clocks}

    # ---- N-GPU == 1-GPU bytes (outside every timed region): rank 0 recomputes ALL frames of the job on its own GPU
    if world > 1 and want_check:
        step_device(0)
        if peer is not None:
            peer.complete()
            gathered = peer.rows() if rank == 0 else None
        else:
            gathered = gather_nccl()
            barrier()
        if rank == 0:
            net.set_option(eb.OPT_STAGED_OUTPUT, 0)
            same, worst = True, 0.0
            for r in range(world):
                rf0, rnf = rank_frames(cfg, world, r)
                if rnf == 0:
                    continue
                seed0 = 1000 * r if cfg["scaling"] == "weak" else 7000 * cfg_id
                cr = torch.from_numpy(clip_frames(cfg, 0 if cfg["scaling"] == "weak" else rf0, rnf, seed0)).to(dev)
                alone = torch.empty((rnf * cpf, 21), dtype=torch.float32, device=dev)
                net.predict_luma_device(cr.data_ptr(), W, H, W, W * H, rnf, qps[0], alone.data_ptr(), stream.cuda_stream)
                torch.cuda.synchronize()
                part = gathered[rf0 * cpf:(rf0 + rnf) * cpf]
                if not torch.equal(part, alone):
                    same = False
                    worst = max(worst, float((part - alone).abs().max().item()))
                del cr, alone
            res["gather_check"] = ("rank 0 recomputed all %d frames of the step alone on one GPU: the gathered rows are bit-identical"
                                   % total_frames) if same else "MISMATCH against the single-GPU recomputation (max |d| %g)" % worst
            net.set_option(eb.OPT_STAGED_OUTPUT, 1 if peer is not None else 0)
            if not same:
                raise SystemExit("N-GPU rows differ from the 1-GPU rows: " + res["gather_check"])
        barrier()

    # ---- end to end through the host API (pinned host buffers; H2D + kernels + D2H inside the timed region)
    if want_e2e:
        net.set_option(eb.OPT_STAGED_OUTPUT, 0)
        for i in range(max(3, warmup)):
            step_e2e(i)
        _, wall_ms = timed(step_e2e, steps)
        e2e_ms = max_over_ranks(wall_ms)
        res["e2e"] = {"value": steps * total_ctus / (e2e_ms * 1e-3), "unit": "CTU/s", "ms_per_step": e2e_ms / steps,
                      "h2d_bytes_per_step": total_frames * W * H, "d2h_bytes_per_step": total_ctus * 84,
                      "timing": "wall clock between device-synchronised barriers, max over ranks",
                      "gather": (None if world == 1 else "every rank's D2H lands in one page-locked shared host block (no collective)"
                                 if host_rows is not None else "H2D + NCCL gather + D2H on rank 0")}
        if host_rows is not None and want_check:
            # rank 0's shared block must hold every rank's rows of the last e2e step
            k = (steps - 1) % nq
            if nf:
                net.predict_luma_device(clips_dev[k % n_clips].data_ptr(), W, H, W, W * H, nf, qps[k], out_dev.data_ptr(), stream.cuda_stream)
            g = gather_nccl()
            torch.cuda.synchronize()
            if rank == 0:
                same = bool(np.array_equal(host_rows.rows(), g.cpu().numpy()))
                res["e2e"]["gather_check"] = "rows in the shared host block are bit-identical to an NCCL gather" if same else "MISMATCH"
                if not same:
                    raise SystemExit("shared host rows differ from the NCCL gather")
        if h2d_probe and nf:
            # what the box can deliver: every rank copies the same luma bytes from pinned memory, nothing else running
            buf = clips_dev[0]
            for _ in range(3):
                buf.copy_(clips_host[0], non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                buf.copy_(clips_host[0], non_blocking=True)
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            gbs = steps * total_frames * W * H / dt / 1e9
            res["e2e"]["h2d_ceiling"] = {"aggregate_gb_s": gbs, "ctu_per_s": gbs * 1e9 / 4096,
                                         "e2e_fraction_of_ceiling": res["e2e"]["value"] / (gbs * 1e9 / 4096),
                                         "how": "bare pinned-host -> device copies of the same bytes on all ranks at once, max over ranks"}
    if peer is not None:
        peer.close()
    if host_rows is not None:
        host_rows.close()
    del clips_dev, clips_host, out_dev, out_host
    torch.cuda.empty_cache()
    return res


def roofline_of(res, peaks):
    """`roofline` of the dominant kernel + the other kernel + the whole path against the HBM roofline."""
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    burst, sustained = float(peaks.get("bf16_tflops", 1590.0)), float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    long_region = res["timed_region_s"] >= 1.0
    tf_peak = sustained if long_region else burst
    stage = res["_stage_raw"]
    steps_ctus = res["ctus_per_step"] * res["timed_region_s"] / (res["ms_per_step"] * 1e-3) / max(1, res.get("_world", 1))  # per rank
    fc_s = max(1e-12, (stage["fc1"]["ms_total"] + stage["heads"]["ms_total"]) * 1e-3)
    conv_s = max(1e-12, stage["conv"]["ms_total"] * 1e-3)
    kern = {"conv": (CONV_FLOP_PER_CTU, conv_s), "fc1": (ALG_FLOP_PER_CTU - CONV_FLOP_PER_CTU, fc_s)}
    dom = "conv" if conv_s >= fc_s else "fc1"
    names = {"conv": "conv", "fc1": "fc (FC1+FC2+FC3)"}
    out = {}
    for key, k in (("roofline", dom), ("roofline_other_kernel", "fc1" if dom == "conv" else "conv")):
        flop, secs = kern[k]
        n_launch = max(1, stage[k]["launches"])
        r = {"kernel": names[k], "bound": "tensor", "achieved": steps_ctus * flop / secs / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
             "peak_source": "%s, %s bf16 (timed region %.3f s)" % (src, "sustained" if long_region else "burst", res["timed_region_s"]),
             "avg_launch_ms": stage[k]["ms_total"] / n_launch,
             "traffic": NCU_DRAM_BYTES_PER_CTU[k] * steps_ctus / n_launch,
             "traffic_source": "dram bytes read + written per launch, scaled per CTU from " + NCU_DRAM_SOURCE[k],
             "algorithmic_bytes_per_launch": ALG_BYTES_PER_CTU * steps_ctus / n_launch,
             "scratch_bytes_per_launch": SCRATCH_BYTES_PER_CTU * steps_ctus / n_launch,
             "note": "algorithmic FLOPs of the kernel (SURVEY 8(d)) / its CUDA-event time; the kernel issues ~3x these FLOPs in fp16 "
                     "(hi*hi + hi*lo + lo*hi split passes for fp32-class accuracy); scratch = the fp16 hi/lo features exchanged "
                     "between the two kernels through HBM, not part of the algorithmic bytes"}
        r["frac"] = r["achieved"] / r["peak"]
        out[key] = r
    per_gpu = res["value"] / max(1, res.get("_world", 1))
    out["roofline_hbm"] = {"kernel": "whole path", "bound": "hbm", "achieved": per_gpu * ALG_BYTES_PER_CTU / 1e9, "peak": hbm_peak,
                           "unit": "GB/s", "peak_source": src, "frac": per_gpu * ALG_BYTES_PER_CTU / 1e9 / hbm_peak}
    out["whole_path_fraction"] = {"hbm_frac": out["roofline_hbm"]["frac"], "tensor_frac": per_gpu * ALG_FLOP_PER_CTU / 1e12 / tf_peak}
    return out


def drop_in_wallclock(model_dir, cfg):
    """Wall clock of the drop-in the way HM uses it (TAppEncCfg.cpp:2319 forks one process per encode) on the config's file:
    the C++ CLI working in-process (pays CUDA start-up) and as a client of the resident server (`--serve`), plus the Python
    shim as a client.  Everything inside: process start, file read, H2D, kernels, D2H, the cu_depth.dat write."""
    pkg = os.path.join(ROOT, "hevc-complexity-reduction_b200")
    cli = os.path.join(pkg, "bin", "video_to_cu_depth")
    if not os.path.exists(cli):
        return None
    work = model_dir
    W, H, nf = cfg["w"], cfg["h"], cfg["frames"]
    yuv = os.path.join(work, "bench_clip.yuv")
    clip = clip_frames(cfg, 0, nf, 4242)
    uv = bytes([128]) * (W * H // 2)
    with open(yuv, "wb") as f:
        for k in range(nf):
            f.write(clip[k].tobytes())
            f.write(uv)
    shim = os.path.join(work, "video_to_cu_depth.py")
    if not os.path.lexists(shim):
        os.symlink(os.path.join(pkg, "video_to_cu_depth.py"), shim)
    n = nf * ctus_per_frame(cfg)
    argv = ["bench_clip.yuv", str(W), str(H), str(cfg["qps"][min(2, len(cfg["qps"]) - 1)])]
    env = dict(os.environ)
    env.pop("ETHCNN_SERVER", None)

    def once(cmd, e):
        out = os.path.join(work, "cu_depth.dat")
        if os.path.exists(out):
            os.remove(out)
        t = time.time()
        r = subprocess.run(cmd, cwd=work, env=e, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        dt = time.time() - t
        if r.returncode != 0 or os.path.getsize(out) != n * 84:
            raise RuntimeError("drop-in failed: " + r.stderr.decode()[-300:])
        return dt
    res = {"file": "%dx%d x %d frames (%.1f MB of 4:2:0)" % (W, H, nf, os.path.getsize(yuv) / 1e6), "ctus": n}
    try:
        res["cli_in_process_s"] = once([cli] + argv, env)
        sock = os.path.join(work, "bench.sock")
        srv = subprocess.Popen([cli, "--serve", sock], cwd=work, stderr=subprocess.DEVNULL)
        try:
            for _ in range(1200):
                if os.path.exists(sock) or srv.poll() is not None:
                    break
                time.sleep(0.05)
            e2 = dict(env, ETHCNN_SERVER=sock)
            once([cli] + argv, e2)   # first request loads the checkpoint
            res["cli_to_resident_server_s"] = min(once([cli] + argv, e2) for _ in range(3))
            res["python_shim_to_resident_server_s"] = min(once([sys.executable, "video_to_cu_depth.py"] + argv, e2) for _ in range(2))
            res["ctu_per_s_through_the_server"] = n / res["cli_to_resident_server_s"]
        finally:
            subprocess.run([cli, "--quit", sock], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            try:
                srv.wait(timeout=30)
            except Exception:
                srv.kill()
    except Exception as e:   # a measurement, not a gate
        res["error"] = str(e)[:200]
    finally:
        for fn in ("bench_clip.yuv", "cu_depth.dat"):
            try:
                os.remove(os.path.join(work, fn))
            except OSError:
                pass
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    import ethcnn_b200 as eb
    from tools import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    cores_per_rank = pin_rank_to_cores(local, world) if (world > 1 and not args.no_pin) else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    main_id = args.config or 2
    todo = [main_id] + ([c for c in (3, 4, 5) if c != main_id] if (args.config is None and not args.only_main) else [])
    cx = Ctx()
    cx.world, cx.rank, cx.local, cx.dev, cx.args = world, rank, local, dev, args
    cx.stream = torch.cuda.current_stream()
    cx.nets = {}
    model_dirs, synthetic = {}, []
    for cid in todo:
        mode = CONFIGS[cid]["mode"]
        if mode not in cx.nets:
            model_dirs[mode] = tempfile.mkdtemp(prefix="ethcnn_bench_")
            synthetic += synth.prepare_models(model_dirs[mode], "LDP" if mode == MODE_LDP else "AI")
            cx.nets[mode] = eb.EthCnn(model_dirs[mode], None, eb.MODE_LDP if mode == MODE_LDP else eb.MODE_AI, device=local)

    dense_path_id = cx.nets[CONFIGS[main_id]["mode"]].query(3)
    sampler = ClockSampler(local).start()   # from the warm-up to the end of the headline's e2e region (both under load)
    main = measure_config(cx, main_id, args.steps, args.warmup, want_e2e=True, want_check=True,
                          sustain_s=0.0, h2d_probe=True)
    clocks = sampler.stop()
    main["_world"] = world
    others, sustained = {}, {}
    if args.config is None and not args.only_main:
        sus = measure_config(cx, main_id, 5, 3, want_e2e=False, want_check=False, sustain_s=args.sustain)
        if "sustained" in sus:
            sustained["config%d" % main_id] = sus["sustained"]
        for cid in todo[1:]:
            r = measure_config(cx, cid, max(3, min(args.steps, 10)), 3, want_e2e=True, want_check=(cid == 3),
                               sustain_s=(args.sustain if cid == 4 else 0.0))
            r["_world"] = world
            if "sustained" in r:
                sustained["config%d" % cid] = r.pop("sustained")
            others["config%d" % cid] = r

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        rl = roofline_of(main, peaks)
        for name, r in others.items():
            rr = roofline_of(r, peaks)
            r["roofline"] = {k: rr["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "avg_launch_ms")}
            r["roofline_hbm_frac"] = rr["roofline_hbm"]["frac"]
            r.pop("_stage_raw"), r.pop("_world")
        sus_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        for name, s in sustained.items():
            cid = int(name[-1])
            per_gpu = s["value"] / world
            c = CONFIGS[cid]
            ctus_rank = c["frames"] * ctus_per_frame(c) * (1.0 if c["scaling"] == "weak" else 1.0 / world)   # per rank per step
            s["whole_path_tensor_frac_of_sustained_peak"] = per_gpu * ALG_FLOP_PER_CTU / 1e12 / sus_peak
            s["conv_tflops"] = CONV_FLOP_PER_CTU * ctus_rank / max(1e-9, s["stage_ms_per_step"]["conv"] * 1e-3) / 1e12
            s["conv_frac_of_sustained_peak"] = s["conv_tflops"] / sus_peak
            s["peak"] = {"bf16_tflops_sustained": sus_peak}

        cpu_baseline = None
        cfg = CONFIGS[main_id]
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(cpu_model_dir(cfg), cfg)
            clip = clip_frames(cfg, 0, min(cfg["frames"], 2 * ref.cores), 0)
            t = time.time()
            qp_b = cfg["qps"][min(2, len(cfg["qps"]) - 1)]
            ref.run(clip[:ref.cores], qp_b)
            per_frame = (time.time() - t) / min(ref.cores, len(clip))
            n = int(max(1, min(len(clip), 15.0 / max(per_frame, 1e-6))))
            t = time.time()
            ref.run(clip[:n], qp_b)
            dt = time.time() - t
            cpu_baseline = {"value": n * ctus_per_frame(cfg) / dt, "unit": "CTU/s", "cores": ref.cores, "kind": "port",
                            "sample": "%d frames of %dx%d at QP %d (oracle port, one worker process per core)" % (n, cfg["w"], cfg["h"], qp_b)}
            ref.close()

        drop_in = None
        if world == 1 and args.config is None and not args.only_main:
            for net in cx.nets.values():   # the CLI / server below open their own handles
                net.close()
            cx.nets_closed = True
            drop_in = drop_in_wallclock(model_dirs[cfg["mode"]], cfg)
        dense = ("simt", "tcgen05 FC1 + heads kernel", "fused tcgen05 FC1+FC2+FC3",
                 "fused tcgen05 FC1+FC2+FC3 on CTA pairs (cta_group::2)")[dense_path_id]
        line = {
            "metric": "CTUs/sec (ETH-CNN inference)", "value": main["value"], "unit": "CTU/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f32-class (3-pass split fp16 on tensor cores, fp32 accumulate: convs mma.sync, FC tcgen05)",
            "data": "synthetic luma; weights: %s" % ("deployed checkpoints" if not synthetic else "synthetic checkpoints for QP %s" % synthetic),
            "config": {"workload": cfg["workload"], "ctus_per_step": main["ctus_per_step"],
                       "sharding": ("single GPU" if world == 1 else
                                    "contiguous frame ranges; rows stored by the gate kernel straight into rank 0's gather buffer over NVLink "
                                    "peer memory (no collective on the data path)" if not args.nccl_gather else
                                    "contiguous frame ranges, NCCL gather to rank 0"),
                       "gather_check": main.get("gather_check"),
                       "l2": "inputs are several times the 126 MB L2 (config 2 rotates over 4 clips = 415 MB; the others hold one sequence "
                             "of 0.5-2.35 GB) + ~11 KB of scratch traffic per CTU",
                       "dense_path": dense, "cores_per_rank": cores_per_rank or None},
            "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"],
            "clocks": clocks,
            "roofline": rl["roofline"], "roofline_other_kernel": rl["roofline_other_kernel"], "roofline_hbm": rl["roofline_hbm"],
            "whole_path_fraction": rl["whole_path_fraction"],
            "stages": main["stages"],
            "ms_per_step_with_stage_events": main.get("ms_per_step_with_stage_events"),
            "timed_region_s": main["timed_region_s"],
            "sustained": sustained or None,
            "other_configs": others or None,
            "drop_in_wallclock": drop_in,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if not getattr(cx, "nets_closed", False):
        for net in cx.nets.values():
            net.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--config", type=int, default=None, choices=sorted(CONFIGS), help="BASELINE config of the JSON line (default 2, "
                    "with the other configs measured briefly under other_configs)")
    ap.add_argument("--only-main", action="store_true", help="skip other_configs and the sustained legs")
    ap.add_argument("--sustain", type=float, default=2.5, help="seconds of back-to-back steps for the sustained legs")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pin", action="store_true", help="N > 1: do not give every rank its own slice of the host cores")
    ap.add_argument("--nccl-gather", action="store_true", help="N > 1: gather with NCCL after the kernels instead of peer-memory stores")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
