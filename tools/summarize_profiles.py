#!/usr/bin/env python
"""Turn the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py <tag> [--launches gpurun_out/launches_X.csv] [--rep name=gpurun_out/prof_X.ncu-rep ...]

Writes profiles/<tag>_launches.md (per-kernel launch count, mean duration, share of the step) and
profiles/<tag>_<name>.md (the raw-page metrics that matter for the roofline: duration, DRAM bytes, pipe
utilisation, shared-memory wavefronts, I-cache hit rate, stall reasons, registers)."""
import argparse
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEEP = (
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "sm__icc_request_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
)


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def write_rep(tag, name, rep):
    m = raw_metrics(rep)
    path = os.path.join(ROOT, "profiles", "%s_%s.md" % (tag, name))
    with open(path, "w") as f:
        f.write("# ncu --set full: %s (%s)\n\n" % (name, os.path.basename(rep)))
        f.write("Kernel: `%s`  grid %s block %s\n\n" % (m.get("Kernel Name", ("?",))[0], m.get("Grid Size", ("?",))[0], m.get("Block Size", ("?",))[0]))
        f.write("Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200; "
                "absolute times are cold-cache / serialised, use them for shares and ratios.\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEEP:
            if k in m:
                f.write("| %s | %s | %s |\n" % (k, m[k][0], m[k][1]))
        f.write("\n## warp stall reasons (warps per issue-active cycle)\n\n| reason | value |\n|---|---|\n")
        st = [(k, m[k][0]) for k in m if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
        for k, v in sorted(st, key=lambda kv: -float(kv[1] or 0)):
            f.write("| %s | %s |\n" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    print("wrote", path)


def write_launches(tag, path_csv):
    rows = [r for r in csv.reader(open(path_csv)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    per_kernel = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki], [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
        per_kernel.setdefault(r[ki], []).append(v)
    tot = sum(a[1] for a in agg.values())
    path = os.path.join(ROOT, "profiles", "%s_launches.md" % tag)
    with open(path, "w") as f:
        f.write("# ncu launch list (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` around "
                "`python bench.py --steps 2 --warmup 3 --no-cpu-baseline`; per-launch times are cold-cache and serialised, "
                "so compare SHARES with bench.py's CUDA-event stage times, not absolutes.\n\n"
                "| kernel | launches | mean us | total us | share | grid (first) | block |\n|---|---|---|---|---|---|---|\n" % os.path.basename(path_csv))
        for k, (n, t, g, b) in agg.items():
            f.write("| `%s` | %d | %.1f | %.1f | %.3f | %s | %s |\n" % (k[:90], n, t / n / 1e3, t / 1e3, t / tot, g, b))
        # The command also runs the end-to-end arm, whose launches work on 8-frame slabs.  The device-resident steps
        # (one launch per kernel over all 25 500 CTUs) are the longest launches of each kernel: their shares are what
        # bench.py's `stages` report.
        big = collections.OrderedDict((k, [v for v in vs if v >= 0.5 * max(vs)]) for k, vs in per_kernel.items())
        tot_big = sum(sum(v) / len(v) for v in big.values())
        f.write("\n## whole-step launches only (duration >= half of the kernel's longest launch)\n\n"
                "| kernel | launches | mean us | share of the step |\n|---|---|---|---|\n")
        for k, vs in big.items():
            f.write("| `%s` | %d | %.1f | %.3f |\n" % (k[:60], len(vs), sum(vs) / len(vs) / 1e3, sum(vs) / len(vs) / tot_big))
    print("wrote", path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep", action="append", default=[])
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if a.launches:
        write_launches(a.tag, a.launches)
    for spec in a.rep:
        name, rep = spec.split("=", 1)
        write_rep(a.tag, name, rep)


if __name__ == "__main__":
    sys.exit(main())
