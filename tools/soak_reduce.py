#!/usr/bin/env python
"""Soak of the reduce kernel's ordering protocol: many launches at random sizes, offsets and
dictionaries, each checked on the device against the dense kernel (count, order, ids).  A protocol
fault would show as a watchdog trap (error status) or a mismatch.

    python tools/soak_reduce.py [seconds]
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pfac_b200 import PFAC  # noqa: E402
from tests import configs  # noqa: E402
from workloads import synth  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
dev = torch.device("cuda:0")
rng = np.random.default_rng(2024)
tmp = tempfile.mkdtemp(prefix="pfac_soak_")
cases = []
for name in ("c2", "c4", "c4dense", "c5"):
    cfg = configs.CONFIGS[name]
    pats = cfg["patterns"]()
    pf = PFAC()
    pf.readPatternFromFile(synth.write_pattern_file(os.path.join(tmp, name + ".pat"), pats))
    big = configs.device_text(cfg, 0, (256 << 20) + 4096, 1 << 40, pats, dev)
    cases.append((name, pf, big))
d_out = torch.empty(256 << 20, dtype=torch.int32, device=dev)
d_id = torch.empty(256 << 20, dtype=torch.int32, device=dev)
d_pos = torch.empty(256 << 20, dtype=torch.int64, device=dev)
t0 = time.time()
n_runs = 0
while time.time() - t0 < budget:
    name, pf, big = cases[int(rng.integers(0, len(cases)))]
    n = int(rng.choice([int(rng.integers(1, 5000)), int(rng.integers(1 << 16, 1 << 22)), int(rng.integers(1 << 22, 256 << 20))]))
    off = int(rng.integers(0, 4096)) if rng.random() < 0.3 else int(rng.integers(0, 256)) * 16
    text = big[off:off + n]
    pf.matchFromDevice(text, n, d_out)
    m = pf.matchFromDeviceReduce64(text, n, d_id, d_pos)
    nz = torch.nonzero(d_out[:n]).flatten()
    assert m == nz.numel(), (name, n, off, m, nz.numel())
    assert torch.equal(nz, d_pos[:m]) and torch.equal(d_out[:n][nz], d_id[:m]), (name, n, off)
    n_runs += 1
print("soak ok: %d reduce launches in %.0f s, all equal to the dense kernel" % (n_runs, time.time() - t0))
