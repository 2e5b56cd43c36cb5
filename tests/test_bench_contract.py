"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the
driver reads; the workload generator is deterministic and shard-consistent."""
import json
import os
import subprocess
import sys

import numpy as np

from workloads import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--bytes", str(32 << 20)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "input_GB_per_s_scanned" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["dtype"] == "u8" and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_shards_are_slices_of_one_stream():
    import bench
    pats = synth.patterns_c2(1000)
    B = 1 << 20
    whole, owned = bench.make_shard(0, 1, 3 * B, pats)
    assert owned == 3 * B and whole.size == 3 * B
    halo = max(len(p) for p in pats) - 1
    for r in range(3):
        shard, own = bench.make_shard(r, 3, B, pats)
        assert own == B
        want = whole[r * B:min((r + 1) * B + halo, 3 * B)]
        assert np.array_equal(shard, want), r
    # patterns really are planted: the oracle finds about one match per 4 KiB
    from oracle import Oracle
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        o = Oracle(synth.write_pattern_file(os.path.join(td, "p"), pats))
        m = int((o.match(whole[:B]) > 0).sum())
    assert 200 <= m <= 320
