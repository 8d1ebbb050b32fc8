"""Per-stage device times of the hot path on a device-resident clip (CUDA events inside the library).

    [ETHCNN_LIB=variant.so] python tools/stage_times.py [--w 1920 --h 1080 --frames 50 --steps 20 --qp 32] [--tag name]

Prints one JSON line: ms per step of the conv / fc / gate stages and of the whole step.  Used to compare kernel variants
(warp counts, ring depths, ...) without the rest of bench.py.  Synthetic frames come from the product-side generator
(tools/synth.py), checkpoints from the staged deployed files when present, synthetic ones otherwise.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ethcnn_b200 as eb  # noqa: E402
from tools import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=50)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--qp", type=int, default=32)
    ap.add_argument("--tag", default="")
    ap.add_argument("--opt", action="append", default=[], help="option=value passed to ethcnn_set_option (numeric ids)")
    args = ap.parse_args()
    d = tempfile.mkdtemp(prefix="stage_times_")
    synth.prepare_models(d)
    W, H, nf = args.w, args.h, args.frames
    cols, rows = (W + 63) // 64, (H + 63) // 64
    base = [synth.synth_frame(W, H, 70 + k) for k in range(min(nf, 5))]
    luma = np.stack([np.roll(base[k % len(base)], 8 * (k // len(base)), axis=0) for k in range(nf)])
    dev = torch.device("cuda", 0)
    dl = torch.from_numpy(luma).to(dev)
    out = torch.empty((nf * cols * rows, 21), dtype=torch.float32, device=dev)
    with eb.EthCnn(d, None, eb.MODE_AI, device=0) as net:
        for o in args.opt:
            k, v = o.split("=")
            net.set_option(int(k), int(v))
        s = torch.cuda.current_stream()
        for _ in range(3):
            net.predict_luma_device(dl.data_ptr(), W, H, W, W * H, nf, args.qp, out.data_ptr(), s.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)   # the step time is taken WITHOUT the per-stage events (they sit between the kernels)
        for _ in range(args.steps):
            net.predict_luma_device(dl.data_ptr(), W, H, W, W * H, nf, args.qp, out.data_ptr(), s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / args.steps
        net.profile_enable(True)
        for st in range(4):
            net.profile_read(st, reset=True)
        e0.record(s)
        for _ in range(args.steps):
            net.predict_luma_device(dl.data_ptr(), W, H, W, W * H, nf, args.qp, out.data_ptr(), s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
        res = {"tag": args.tag, "lib": os.path.basename(eb.library_path()), "ctus": nf * cols * rows,
               "step_ms": step_ms, "step_ms_with_stage_events": e0.elapsed_time(e1) / args.steps}
        for st, name in enumerate(eb.STAGE_NAMES):
            ms, n = net.profile_read(st, reset=True)
            res[name + "_ms"] = ms / args.steps
        res["ctu_per_s"] = res["ctus"] / (res["step_ms"] * 1e-3)
        res["checksum"] = float(out.double().sum().item())
        net.profile_enable(False)
        lib = eb.load_library()
        if hasattr(lib, "ethcnn_debug_conv_phases"):   # measurement builds (-DETHCNN_EXP_TIMING): cycles per phase of the conv warp tasks
            import ctypes as C
            buf = (C.c_ulonglong * 24)()
            lib.ethcnn_debug_conv_phases(buf, 1)
            net.predict_luma_device(dl.data_ptr(), W, H, W, W * H, nf, args.qp, out.data_ptr(), s.cuda_stream)
            torch.cuda.synchronize()
            lib.ethcnn_debug_conv_phases(buf, 0)
            names = ("setup", "load_x", "conv1+act+conv2", "act2+stores", "conv3", "act3+stores", "wait_tiles", "-")
            res["conv_phase_cycles_per_ctu"] = {br: {names[k]: round(buf[8 * b + k] / res["ctus"], 1) for k in range(7)}
                                                for b, br in enumerate(("S", "M", "L"))}
        if hasattr(lib, "ethcnn_debug_fc_issuer"):     # -DETHCNN_EXP_FC_TIMING: where the MMA issuer of the fused FC kernel waits
            import ctypes as C
            b8 = (C.c_ulonglong * 8)()
            e8 = (C.c_ulonglong * 8)()
            lib.ethcnn_debug_fc_issuer(b8, 1)
            lib.ethcnn_debug_fc_epi(e8, 1)
            net.predict_luma_device(dl.data_ptr(), W, H, W, W * H, nf, args.qp, out.data_ptr(), s.cuda_stream)
            torch.cuda.synchronize()
            lib.ethcnn_debug_fc_issuer(b8, 0)
            lib.ethcnn_debug_fc_epi(e8, 0)
            en = ("wait_acc1_full", "epi1_wait_stage", "epi1_ld_math_store", "epi1_fence_arrive", "epi2")
            res["fc_epilogue_warp_cycles_per_leader_cta"] = {en[k]: round(e8[k] / max(1, e8[6])) for k in range(5)}
            nm = ("wait_acc1_empty", "fc1_wait_operands", "fc1_issue", "wait_acc2_empty", "fc2_wait", "fc2_issue")
            res["fc_issuer_cycles_per_leader_cta"] = {nm[k]: round(b8[k] / max(1, b8[6])) for k in range(6)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
