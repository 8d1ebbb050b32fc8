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
