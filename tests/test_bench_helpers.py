"""Host-side logic of bench.py that the N > 1 runs rely on (no GPU): frame sharding of the BASELINE configs, content that does
not depend on the sharding, identical workload strings in both arms."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_frame_ranges_of_the_baseline_configs():
    c3, c4 = bench.CONFIGS[3], bench.CONFIGS[4]
    assert [bench.rank_frames(c3, 8, r)[1] for r in range(8)] == [7, 7, 6, 6, 6, 6, 6, 6]
    assert [bench.rank_frames(c4, 4, r)[1] for r in range(4)] == [107, 106, 106, 106]             # config 4
    for cfg, world in ((c3, 8), (c4, 4), (bench.CONFIGS[5], 3)):
        nxt = 0
        for r in range(world):
            f0, nf = bench.rank_frames(cfg, world, r)
            assert f0 == nxt
            nxt += nf
        assert nxt == cfg["frames"]
    assert bench.rank_frames(bench.CONFIGS[2], 8, 5) == (250, 50)                                   # weak scaling: 50 frames per rank
    assert [bench.ctus_per_frame(bench.CONFIGS[k]) for k in (2, 3, 4, 5)] == [510, 3927, 1350, 510]


def test_strong_scaled_content_is_independent_of_the_sharding():
    cfg = dict(bench.CONFIGS[5], w=192, h=136, frames=11)
    whole = bench.clip_frames(cfg, 0, 11, 70)
    for world in (2, 3):
        parts = [bench.clip_frames(cfg, *bench.frame_range(11, world, r), 70) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), whole)
    assert not np.array_equal(whole[0], whole[5])        # frame k and k + 5 share a base frame but are shifted


def test_both_arms_name_the_same_workload():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"workload": cfg["workload"]') >= 2   # run_reference and run_ours take the string from CONFIGS
    assert len({c["workload"] for c in bench.CONFIGS.values()}) == 4


def test_committed_bench_lines_carry_the_contract_keys_and_the_summary_regenerates(tmp_path):
    """The bench lines kept under profiles/ (what DESIGN.md quotes) carry every key the bench contract names, both arms name the
    same workload, and profiles/r02_summary.md is reproducible from them with tools/summarize_bench.py."""
    import glob
    import json
    import subprocess
    import sys
    prof = os.path.join(ROOT, "profiles")
    lines = {}
    for path in sorted(glob.glob(os.path.join(prof, "r02[ij]_bench_n*.json"))):
        with open(path) as f:
            d = json.loads([l for l in f.read().splitlines() if l.startswith("{")][-1])
        lines[d["n_gpus"]] = (path, d)
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "roofline", "clocks", "e2e", "gpu_launches"):
            assert key in d, (path, key)
        assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert d["roofline"]["bound"] in ("hbm", "tensor") and 0 < d["roofline"]["frac"] < 1
        assert abs(d["roofline"]["achieved"] / d["roofline"]["peak"] - d["roofline"]["frac"]) < 1e-9
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] > 1:
            assert "bit-identical" in d["config"]["gather_check"]
    assert sorted(lines) == [1, 2, 4, 8]
    ref_path = glob.glob(os.path.join(prof, "r02[ij]_bench_reference.json"))[-1]
    with open(ref_path) as f:
        ref = json.loads([l for l in f.read().splitlines() if l.startswith("{")][-1])
    one = lines[1][1]
    assert ref["impl"] == "reference" and ref["config"]["workload"] == one["config"]["workload"]
    assert ref["metric"] == one["metric"] and ref["unit"] == one["unit"]
    assert one["cpu_baseline"]["kind"] in ("port", "reference") and one["cpu_baseline"]["cores"] >= 1

    want = open(os.path.join(prof, "r02_summary.md")).read()
    note = want.split("inside the timed region.  ", 1)[1].split("\n", 1)[0]
    cli8 = glob.glob(os.path.join(prof, "r02[ij]_cli_wallclock_8gpu_server.txt"))[-1]
    tool = open(os.path.join(ROOT, "tools", "summarize_bench.py")).read().replace(
        'os.path.join(ROOT, "profiles", "%s_summary.md" % a.round)', repr(str(tmp_path / "summary.md")))
    script = tmp_path / "summarize.py"
    script.write_text(tool.replace('ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))', "ROOT = %r" % ROOT))
    args = ["%d=%s" % (n, os.path.relpath(p, ROOT)) for n, (p, _) in sorted(lines.items())]
    subprocess.run([sys.executable, str(script), "r02", *args, "--reference", os.path.relpath(ref_path, ROOT),
                    "--cli8", os.path.relpath(cli8, ROOT), "--note", note], check=True, cwd=ROOT, capture_output=True)
    assert (tmp_path / "summary.md").read_text() == want
