#!/usr/bin/env python
"""Time PFAC_matchFromHost / PFAC_matchFromHostReduce on pageable (malloc-style) and pinned host
buffers.  The reference's callers pass plain malloc'ed memory (reference test/simple_example.cpp),
so the pageable row is what a drop-in user sees; bench.py's `e2e` is the pinned row.

    python tools/host_path_bench.py [--mib 1024] [--reps 3]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--skip-pinned", action="store_true")
    args = ap.parse_args()
    import torch
    from pfac_b200 import PFAC
    from workloads import synth

    n = args.mib << 20
    pats = synth.patterns_c2(1000)
    text = synth.make_text("random", synth.SEED_BASE + 2, 0, n, n, pats, every=4096)
    tmp = tempfile.mkdtemp(prefix="pfac_hp_")
    pfile = synth.write_pattern_file(os.path.join(tmp, "p.txt"), pats)
    pf = PFAC()
    pf.readPatternFromFile(pfile)

    def best(fn):
        fn()
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = {"bytes": n, "cores": os.cpu_count(),
           "env": {k: v for k, v in os.environ.items() if k.startswith("PFAC_B200_")}}
    # pageable: numpy-owned memory, touched once so page faults are not in the timed region
    h_in = np.array(text, copy=True)
    h_out = np.zeros(n, dtype=np.int32)
    h_id = np.zeros(n // 8, dtype=np.int32)
    h_pos = np.zeros(n // 8, dtype=np.int32)
    t = best(lambda: pf.matchFromHost(h_in, h_out))
    out["dense_pageable_GBps"] = n / t / 1e9
    ref = h_out.copy()
    t = best(lambda: pf.matchFromHostReduce(h_in, h_id, h_pos))
    out["reduce_pageable_GBps"] = n / t / 1e9
    if args.skip_pinned:
        print(json.dumps(out))
        with open(os.path.join(ROOT, "gpurun_out", "host_path.jsonl"), "a") as f:
            f.write(json.dumps(out) + "\n")
        return
    # pinned
    p_in = torch.from_numpy(text).pin_memory()
    p_out = torch.zeros(n, dtype=torch.int32).pin_memory()
    p_id = torch.zeros(n // 8, dtype=torch.int32).pin_memory()
    p_pos = torch.zeros(n // 8, dtype=torch.int32).pin_memory()
    t = best(lambda: pf.matchFromHost(p_in, p_out))
    out["dense_pinned_GBps"] = n / t / 1e9
    t = best(lambda: pf.matchFromHostReduce(p_in, p_id, p_pos))
    out["reduce_pinned_GBps"] = n / t / 1e9
    out["pageable_equals_pinned"] = bool(np.array_equal(ref, p_out.numpy()))
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "host_path.jsonl"), "a") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
