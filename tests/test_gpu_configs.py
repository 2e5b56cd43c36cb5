"""The BASELINE.json configs at their full sizes, through the C ABI, against the reference's own CPU
matcher (oracle/_ref when present, else the C restatement), every position checked.

What reference test/omp_PFAC.cpp:397-439 does for its multi-GPU run (element-by-element comparison
with a second implementation), here per config:
  C2  1,000 patterns, 1 GiB: dense array + reduced list
  C3  20,000 Snort-like patterns, 2**32 + 4099 bytes: dense array (64-bit indexing) and the 64-bit
      reduced list of an input >= 2**31 bytes
  C4  DNA 5,000 patterns, 2e9 bytes, reduce in both perf modes; and the 10 %-density variant
  C5  10,000 patterns sharded over 2 ranks with the NCCL count scan (needs 2 GPUs, else skipped)
Text comes from the device generator (tests/test_devgen.py pins it to the numpy definition); the
oracle sees it in 256 MiB chunks + halo (its input_size is an int).
"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import pytest

from tests import configs
from tests.helpers import CheckerOracle
from workloads import synth

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GIB = 1 << 30


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def _setup(name, tmp_path):
    from pfac_b200 import PFAC
    cfg = configs.CONFIGS[name]
    pats = cfg["patterns"]()
    pfile = synth.write_pattern_file(str(tmp_path / (name + ".pat")), pats)
    CheckerOracle.set_threads(os.cpu_count() or 1)
    orc = CheckerOracle(pfile)
    pf = PFAC()
    pf.readPatternFromFile(pfile)
    return cfg, pats, orc, pf


def _time_ms(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _check_dense_and_list(orc, d_text, n, halo, d_out, got_ids, got_pos):
    """Every position of the dense array and every entry of the reduced list vs the oracle."""
    dev = d_text.device
    mism = 0
    k = 0
    t0 = time.time()
    for c0, c1, want in configs.ChunkedCheck(orc, d_text, n, halo):
        if d_out is not None:
            mism += int((d_out[c0:c1] != torch.from_numpy(want).to(dev)).sum().item())
        if got_ids is not None:
            ids, pos = configs.nonzero_pairs(want, c0)
            m = ids.size
            g_ids = got_ids[k:k + m].cpu().numpy()
            g_pos = got_pos[k:k + m].cpu().numpy().astype(np.int64)
            mism += int(g_ids.size != m) + int((g_ids != ids[:g_ids.size]).sum()) + int((g_pos != pos[:g_pos.size]).sum())
            k += m
    return mism, k, time.time() - t0


def test_c2_full_size(cuda, tmp_path):
    """BASELINE config 2 at 1 GiB: PFAC_matchFromDevice and PFAC_matchFromDeviceReduce, all positions."""
    cfg, pats, orc, pf = _setup("c2", tmp_path)
    n = cfg["bytes"]
    d_text = configs.device_text(cfg, 0, n, n, pats, cuda)
    d_out = torch.full((n,), -7, dtype=torch.int32, device=cuda)
    ms_dense = _time_ms(lambda: pf.matchFromDevice(d_text, n, d_out))
    cap = n // 16
    d_id = torch.full((cap,), -7, dtype=torch.int32, device=cuda)
    d_pos = torch.full((cap,), -7, dtype=torch.int32, device=cuda)
    ms_red = _time_ms(lambda: pf.matchFromDeviceReduce(d_text, n, d_id, d_pos))
    m = pf.matchFromDeviceReduce(d_text, n, d_id, d_pos)
    mism, k, secs = _check_dense_and_list(orc, d_text, n, orc.max_pattern_len - 1, d_out, d_id, d_pos)
    configs.append_result({"test": "c2_full_size", "bytes": n, "matches": m, "oracle": orc.kind, "oracle_s": round(secs, 1),
                           "dense_ms": ms_dense, "dense_GBps": n / ms_dense / 1e6, "reduce_ms_call": ms_red,
                           "reduce_GBps": n / ms_red / 1e6, "mismatches": mism})
    assert m == k and mism == 0
    assert int(d_id[m].item()) == -7 and int(d_pos[m].item()) == -7   # nothing written past the M entries
    pf.destroy()


def test_c3_beyond_4gib_dense_and_reduce64(cuda, tmp_path):
    """BASELINE config 3 with N = 2**32 + 4099: dense (positions index past 32 bits) and the 64-bit
    reduced list of an input >= 2**31 bytes; 20,000 patterns, tables in L2."""
    cfg, pats, orc, pf = _setup("c3", tmp_path)
    n = (1 << 32) + 4099
    d_text = configs.device_text(cfg, 0, n, n, pats, cuda)
    d_out = torch.empty(n, dtype=torch.int32, device=cuda)
    d_out.fill_(-7)
    ms_dense = _time_ms(lambda: pf.matchFromDevice(d_text, n, d_out), reps=2)
    cap = n // 16
    d_id = torch.full((cap,), -7, dtype=torch.int32, device=cuda)
    d_pos = torch.full((cap,), -7, dtype=torch.int64, device=cuda)
    ms_red = _time_ms(lambda: pf.matchFromDeviceReduce64(d_text, n, d_id, d_pos), reps=2)
    m = pf.matchFromDeviceReduce64(d_text, n, d_id, d_pos)
    assert m < cap
    mism, k, secs = _check_dense_and_list(orc, d_text, n, orc.max_pattern_len - 1, d_out, d_id, d_pos)
    configs.append_result({"test": "c3_beyond_4gib", "bytes": n, "matches": m, "oracle": orc.kind, "oracle_s": round(secs, 1),
                           "states": orc.num_states, "dense_ms": ms_dense, "dense_GBps": n / ms_dense / 1e6,
                           "reduce64_ms_call": ms_red, "reduce64_GBps": n / ms_red / 1e6, "mismatches": mism})
    assert m == k and mism == 0
    assert int(d_pos[m - 1].item()) >= (1 << 32) - 4096   # the list reaches past 2**32 - one plant per 2 KiB
    # the legacy int API refuses what it cannot index
    from pfac_b200 import PFACError
    with pytest.raises(PFACError):
        pf.matchFromDeviceReduce(d_text, n, d_id, d_pos[:cap // 2].view(torch.int32))
    pf.destroy()


@pytest.mark.parametrize("name", ["c4", "c4dense"])
def test_c4_dna_reduce_both_perf_modes(cuda, tmp_path, name):
    """BASELINE config 4 at 2e9 bytes: PFAC_matchFromDeviceReduce under PFAC_TIME_DRIVEN (reduceOnDevice)
    and PFAC_SPACE_DRIVEN (reduceInplaceOnDevice) give the oracle's list."""
    from pfac_b200 import PerfMode
    cfg, pats, orc, pf = _setup(name, tmp_path)
    n = cfg["bytes"]
    d_text = configs.device_text(cfg, 0, n, n, pats, cuda)
    cap = n // 4 if name == "c4dense" else n // 16
    d_id = torch.full((cap,), -7, dtype=torch.int32, device=cuda)
    d_pos = torch.full((cap,), -7, dtype=torch.int32, device=cuda)
    ms_red = _time_ms(lambda: pf.matchFromDeviceReduce(d_text, n, d_id, d_pos))
    m = pf.matchFromDeviceReduce(d_text, n, d_id, d_pos)
    assert m < cap
    mism, k, secs = _check_dense_and_list(orc, d_text, n, orc.max_pattern_len - 1, None, d_id, d_pos)
    assert m == k and mism == 0
    first_ids, first_pos = d_id[:m].clone(), d_pos[:m].clone()
    d_id.fill_(-7)
    d_pos.fill_(-7)
    pf.setPerfMode(PerfMode.SPACE_DRIVEN)
    ms_space = _time_ms(lambda: pf.matchFromDeviceReduce(d_text, n, d_id, d_pos, alias="reduceInplaceOnDevice"), reps=1)
    m2 = pf.matchFromDeviceReduce(d_text, n, d_id, d_pos, alias="reduceInplaceOnDevice")
    assert m2 == m and torch.equal(d_id[:m], first_ids) and torch.equal(d_pos[:m], first_pos)
    configs.append_result({"test": name + "_full_size", "bytes": n, "matches": m, "oracle": orc.kind,
                           "oracle_s": round(secs, 1), "reduce_ms_call": ms_red, "reduce_GBps": n / ms_red / 1e6,
                           "space_driven_ms_call": ms_space, "space_driven_GBps": n / ms_space / 1e6, "mismatches": mism})
    pf.destroy()


def test_c5_two_ranks_nccl_global_list(cuda, tmp_path):
    """BASELINE config 5 on 2 ranks (one process per GPU, NCCL): every rank's reduced run against the
    oracle with global 64-bit positions, offsets = exclusive scan of the counts, one global list on rank 0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "c5.jsonl"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "run_configs.py"),
           "--config", "c5", "--bytes", str(2 * GIB), "--check-bytes", "-1", "--steps", "3", "--gather",
           "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [json.loads(l) for l in open(out)]
    assert len(lines) == 2 and all(l["bit_exact"] for l in lines)
    by_rank = {l["rank"]: l for l in lines}
    assert by_rank[0]["global_offset"] == 0 and by_rank[1]["global_offset"] == by_rank[0]["matches_rank"]
    assert by_rank[0]["matches_total"] == by_rank[0]["matches_rank"] + by_rank[1]["matches_rank"]
    assert by_rank[0].get("gather_ok") is True and by_rank[0].get("p2p_gather_ok") is True


@pytest.mark.parametrize("name", ["wprefix_dense", "wan_dense"])
def test_adversarial_texts(cuda, tmp_path, name):
    """Worst cases: a text of concatenated pattern prefixes (every piece passes the first stage and walks
    deep) and a^n against a^1..a^32 (every position matches after the longest walk; the parking and
    spill rings of the reduce kernel run full): dense array and reduced list equal the oracle's."""
    cfg, pats, orc, pf = _setup(name, tmp_path)
    n = (96 << 20) + 333
    d_text = configs.device_text(cfg, 0, n, n, pats, cuda)
    d_out = torch.full((n,), -7, dtype=torch.int32, device=cuda)
    ms_dense = _time_ms(lambda: pf.matchFromDevice(d_text, n, d_out), reps=2)
    d_id = torch.full((n,), -7, dtype=torch.int32, device=cuda)
    d_pos = torch.full((n,), -7, dtype=torch.int32, device=cuda)
    ms_red = _time_ms(lambda: pf.matchFromDeviceReduce(d_text, n, d_id, d_pos), reps=2)
    m = pf.matchFromDeviceReduce(d_text, n, d_id, d_pos)
    mism, k, secs = _check_dense_and_list(orc, d_text, n, orc.max_pattern_len - 1, d_out, d_id, d_pos)
    configs.append_result({"test": name, "bytes": n, "matches": m, "oracle": orc.kind, "oracle_s": round(secs, 1),
                           "dense_ms": ms_dense, "dense_GBps": n / ms_dense / 1e6, "reduce_ms_call": ms_red,
                           "reduce_GBps": n / ms_red / 1e6, "mismatches": mism})
    assert m == k and mism == 0
    if name == "wan_dense":
        assert m == n   # every position matches
    pf.destroy()
