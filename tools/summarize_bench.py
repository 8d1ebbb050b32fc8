#!/usr/bin/env python
"""Turn the bench.py JSON lines kept under profiles/ into profiles/<round>_summary.md.

    python tools/summarize_bench.py r02 1=profiles/r02j_bench_n1.json 2=profiles/r02i_bench_n2.json ... \
        [--reference profiles/r02j_bench_reference.json] [--cli8 profiles/r02i_cli_wallclock_8gpu_server.txt]

Every number in the output is read from those files; nothing is typed in by hand."""
import argparse
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def last_json_line(path):
    with open(path) as f:
        lines = [l for l in f.read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def e6(v):
    return "%.1fe6" % (v / 1e6) if v >= 20e6 else "%.2fe6" % (v / 1e6)


def check_word(s):
    if not s:
        return "-"
    return "recomputed on one GPU: bit-identical" if "bit-identical" in s else s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("round")
    ap.add_argument("lines", nargs="+", help="N=path of a bench.py line")
    ap.add_argument("--reference")
    ap.add_argument("--cli8")
    ap.add_argument("--note", help="a sentence appended to the introduction (e.g. which lines predate a kernel change)")
    a = ap.parse_args()
    runs = {}
    for spec in a.lines:
        n, path = spec.split("=", 1)
        runs[int(n)] = (os.path.relpath(os.path.join(ROOT, path), os.path.join(ROOT, "profiles")), last_json_line(path))
    ns = sorted(runs)
    d1 = runs[1][1]
    out = []
    w = out.append
    w("# Round %s measurements (B200; full JSON lines: %s%s)\n" % (
        a.round[1:].lstrip("0"), ", ".join(runs[n][0] for n in ns),
        ", " + os.path.basename(a.reference) if a.reference else ""))
    w("`python bench.py --steps %d --warmup %d` (N = 1) / the same under torchrun (N > 1), each on a fresh box. `value` = "
      "device-resident CTU/s (CUDA events on the launching stream, max over ranks, nothing but the library's launches in the "
      "timed region), `e2e` = host API from pinned memory: H2D + kernels + D2H inside the timed region.%s\n" % (
          d1["steps"], d1["warmup"], "  " + a.note if a.note else ""))

    w("## Headline: BASELINE %s, weak scaling\n" % d1["config"]["workload"])
    w("| N | value CTU/s | per GPU | vs N x (N = 1) | ms per step | e2e CTU/s | bare pinned-H2D ceiling of the box | e2e / ceiling | N-GPU == 1-GPU rows |")
    w("|---|---|---|---|---|---|---|---|---|")
    for n in ns:
        d = runs[n][1]
        c = d["e2e"].get("h2d_ceiling") or {}
        w("| %d | %s | %s | %.3f | %.4f | %s | %.1f GB/s = %s CTU/s | %.3f | %s |" % (
            n, e6(d["value"]), e6(d["value"] / n), d["value"] / (n * d1["value"]), d["ms_per_step"], e6(d["e2e"]["value"]),
            c.get("aggregate_gb_s", 0), e6(c.get("ctu_per_s", 0)), c.get("e2e_fraction_of_ceiling", 0),
            check_word(d["config"].get("gather_check"))))
    w("")
    cb = d1["cpu_baseline"]
    ref = last_json_line(a.reference) if a.reference else None
    w("CPU port of the reference on the same box (%s, %d cores): %.1fe3 CTU/s inside the N = 1 run%s.  e2e is bound by the "
      "host-to-device copy of 4 KB of luma per CTU; the last two columns compare it with bare pinned copies of the same bytes "
      "issued on all ranks at once on the same box.\n" % (
          cb["sample"], cb["cores"], cb["value"] / 1e3,
          ", %.1fe3 as `--impl reference`" % (ref["value"] / 1e3) if ref else ""))

    st = d1["stages"]
    r, ro, rh, wp = d1["roofline"], d1["roofline_other_kernel"], d1["roofline_hbm"], d1["whole_path_fraction"]
    w("Stages at N = 1 (CUDA events in a separate profiled loop, ms per %d-CTU step): conv %.4f, fused FC %.4f, gate %.4f; the "
      "step itself %.4f ms (programmatic dependent launch: the prologue of each kernel overlaps the tail of its predecessor; "
      "%.4f ms with the stage events between the kernels).  Roofline (dominant kernel = conv): %.1f TFLOP/s algorithmic = %.3f "
      "of the %s; fused FC %.1f TFLOP/s = %.3f; whole path %.3f of the tensor peak, %.3f of the HBM roofline (%.0f of %.0f GB/s) "
      "on its algorithmic bytes.  DRAM bytes per launch (ncu, %s): conv %.1f MB, FC %.1f MB against %.1f MB algorithmic per "
      "step (the difference is the fp16 hi/lo feature scratch, %.0f MB written and read once).\n" % (
          d1["config"]["ctus_per_step"], st["conv"]["ms_per_step"], st["fc1"]["ms_per_step"], st["gate"]["ms_per_step"],
          d1["ms_per_step"], d1["ms_per_step_with_stage_events"], r["achieved"], r["frac"], r["peak_source"],
          ro["achieved"], ro["frac"], wp["tensor_frac"], wp["hbm_frac"], rh["achieved"], rh["peak"],
          r["traffic_source"].split("from ")[-1], r["traffic"] / 1e6, ro["traffic"] / 1e6,
          r["algorithmic_bytes_per_launch"] / 1e6, 2 * r["scratch_bytes_per_launch"] / 1e6))

    dw = d1.get("drop_in_wallclock")
    if dw:
        s = ("Drop-in wall clock at N = 1 on the config-2 file (%s): C++ CLI in-process %.2f s (CUDA start-up), CLI -> resident "
             "server %.4f s, Python shim -> server %.3f s." % (dw["file"], dw["cli_in_process_s"], dw["cli_to_resident_server_s"],
                                                               dw["python_shim_to_resident_server_s"]))
        if a.cli8:
            rows = []
            for line in open(a.cli8):
                m = re.match(r"(config\d)\s+(\S+)\s+x(\d+)\s+\d+ CTUs\s+CLI->server\s+wall ([\d.]+) s", line)
                if m:
                    rows.append("%s x %s frames %s s" % (m.group(2), m.group(3), m.group(4)))
            s += "  Through a server sharding over 8 GPUs (`%s`): %s." % (os.path.basename(a.cli8), ", ".join(rows))
        w(s + "\n")

    w("## Sustained (the same step looped back to back for >= 2.5 s, clocks sampled meanwhile)\n")
    w("| N | config | CTU/s | SM clock under load | throttle reasons | conv TFLOP/s (frac of the measured sustained bf16 peak) | whole path frac of sustained peak |")
    w("|---|---|---|---|---|---|---|")
    for n in ns:
        for k, v in sorted(runs[n][1].get("sustained", {}).items()):
            w("| %d | %s | %s | %.0f MHz | %s | %.1f (%.3f) | %.3f |" % (
                n, k, e6(v["value"]), v["clocks"]["sm_mhz"], ", ".join(v["clocks"]["reasons"]) or "none",
                v["conv_tflops"], v["conv_frac_of_sustained_peak"], v["whole_path_tensor_frac_of_sustained_peak"]))
    w("")

    w("## The other BASELINE configs (strong scaling: the sequence is split in contiguous frame ranges)\n")
    w("| N | config | value CTU/s | speed-up over N = 1 | ms per step | e2e CTU/s | check |")
    w("|---|---|---|---|---|---|---|")
    for n in ns:
        for k, v in sorted(runs[n][1].get("other_configs", {}).items()):
            base = d1["other_configs"][k]["value"]
            w("| %d | %s | %s | %.2f | %.3f | %s | %s |" % (
                n, k, e6(v["value"]), v["value"] / base, v["ms_per_step"], e6(v["e2e"]["value"]), check_word(v.get("gather_check"))))
    w("")
    for k, v in sorted(d1.get("other_configs", {}).items()):
        w("* %s (%d CTUs per step)" % (v["workload"], v["ctus_per_step"]))
    path = os.path.join(ROOT, "profiles", "%s_summary.md" % a.round)
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
