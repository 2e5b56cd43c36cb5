#!/usr/bin/env python
"""Parity + timing runner for the BASELINE.json configs (lives under tests/ because it loads the
oracle).  Runs one config end to end: regenerate the synthetic text on the GPU (workloads/devgen),
check the CUDA path bit-exactly against the reference's CPU matcher (oracle/_ref when present, else the
C port; chunked, so inputs >= 2^31 bytes work), time it, and print one JSON line (also appended to
gpurun_out/configs.jsonl; tools/keep_results.py moves those lines into profiles/).  The same configs at
full size are part of the GPU suite (tests/test_gpu_configs.py); this script is for timing and ncu runs.

    python tests/run_configs.py --config c2|c3|c4|c4dense|c5|wprefix_*|wan_* [--bytes N] [--check-bytes M] [--steps K]
    python -m torch.distributed.run --nproc-per-node 8 ... tests/run_configs.py --config c5 [--gather]

Configs:
  c2       1,000 patterns len 4-32 (255-symbol alphabet), 1 GiB planted random text, dense
  c3       20,000 Snort-like patterns, 4 GiB ASCII-weighted text, dense (64-bit indexing)
  c4       DNA, 5,000 patterns len 8-24, 2e9 bytes, reduce (both perf modes), natural density
  c4dense  c4 + 64 short patterns (len 4-6): high match density compaction
  c5       10,000 Snort-like patterns, 32 GiB sharded over the ranks, reduce + global offset scan (both the
           NCCL all-gather cross-check and the in-kernel scan over peer memory, PFAC_comm)
  wprefix_dense / wprefix_reduce, wan_dense / wan_reduce   adversarial texts (tests/configs.py)
The oracle is test infrastructure (oracle/); the timed path is libpfac.so only.
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from workloads import synth  # noqa: E402
from pfac_b200.sharding import shard_bounds  # noqa: E402

from tests.configs import CONFIGS, GIB, ChunkedCheck, device_text, nonzero_pairs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=sorted(CONFIGS))
    ap.add_argument("--bytes", type=int, default=0, help="override the total input size")
    ap.add_argument("--check-bytes", type=int, default=-1,
                    help="bytes per rank to verify against the oracle (-1 = all, 0 = none)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--perf-mode", type=int, default=0)
    ap.add_argument("--out", default="", help="also append the JSON line to this file")
    ap.add_argument("--gather", action="store_true",
                    help="multi-GPU reduce: also deliver every rank's run to rank 0 (sharding.place_runs) and time it")
    args = ap.parse_args()

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    from tests.helpers import CheckerOracle
    from pfac_b200 import PFAC
    from pfac_b200.sharding import allgather_count_offsets, place_runs

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = CONFIGS[args.config]
    total_len = args.bytes or cfg["bytes"]
    pats = cfg["patterns"]()
    tmp = tempfile.mkdtemp(prefix="pfac_cfg_")
    pfile = synth.write_pattern_file(os.path.join(tmp, "p%d.txt" % rank), pats)
    t0 = time.time()
    pf = PFAC()
    pf.readPatternFromFile(pfile)
    compile_s = time.time() - t0
    if args.perf_mode:
        pf.setPerfMode(args.perf_mode)
    info = pf.tableInfo()
    maxlen = info["max_pattern_len"]

    start, owned, total = shard_bounds(total_len, world, rank, maxlen)
    t0 = time.time()
    d_in = device_text(cfg, start, total, total_len, pats, dev)   # regenerated on the GPU from the counter hash
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    api = cfg["api"]
    res = {"config": args.config, "api": api, "rank": rank, "world": world, "total_bytes": total_len,
           "owned_bytes": owned, "patterns": len(pats), "states": info["num_states"],
           "max_pattern_len": maxlen, "table": {k: info[k] for k in (
               "hash_edges", "num_chains", "tail_bytes", "hot_depth", "hot_buckets", "cold_buckets",
               "next2_hot", "chains_hot", "pre2_bits_set", "device_bytes")},
           "compile_s": round(compile_s, 3), "gen_s": round(gen_s, 3), "perf_mode": args.perf_mode}

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- run + time ---------------------------------------------------------------------------
    if api == "dense":
        d_out = torch.empty(owned, dtype=torch.int32, device=dev)
        run = lambda: pf.matchShardFromDevice(d_in, owned, total, d_out)  # noqa: E731
    else:
        cap = owned if args.config in ("c4dense", "wan_reduce") else max(owned // 8, 1 << 20)
        d_id = torch.empty(cap, dtype=torch.int32, device=dev)
        pos64 = api == "reduce64" or owned >= 2 ** 31
        d_pos = torch.empty(cap, dtype=torch.int64 if pos64 else torch.int32, device=dev)
        if pos64:
            run = lambda: pf.matchShardFromDeviceReduce64(d_in, owned, total, start, d_id, d_pos)  # noqa: E731
        else:
            run = lambda: pf.matchFromDeviceReduce(d_in, owned, d_id, d_pos)  # noqa: E731
    for _ in range(3):
        m = run()
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        m = run()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = ev0.elapsed_time(ev1) / args.steps if api == "dense" else wall / args.steps * 1e3
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    res["ms_per_step"] = ms_max
    res["input_GBps_all_ranks"] = total_len / (ms_max * 1e-3) / 1e9

    # ---- cross-GPU step for the reduced list ------------------------------------------------------
    if api != "dense":
        count = int(m)
        if world > 1:
            t0 = time.perf_counter()
            off, total_m, counts = allgather_count_offsets(count, device=dev)
            res["count_scan_ms"] = (time.perf_counter() - t0) * 1e3
        else:
            off, total_m = 0, count
        if world > 1 and args.gather and pos64:
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            g_ids, g_pos = place_runs(d_id, d_pos, counts, dst=0)
            torch.cuda.synchronize()
            dist.barrier()
            res["gather_ms"] = (time.perf_counter() - t0) * 1e3
            res["gather_bytes"] = 12 * total_m
            if rank == 0:   # one global list, ascending positions, this rank's run at its offset
                ok = bool((g_pos[1:] > g_pos[:-1]).all().item()) if total_m > 1 else True
                ok = ok and g_ids.numel() == total_m and bool(torch.equal(g_ids[:count], d_id[:count]))
                res["gather_ok"] = ok
            del g_ids, g_pos
        res["matches_rank"] = count
        res["matches_total"] = total_m
        res["global_offset"] = off
        res["algorithmic_bytes"] = owned + (12 if pos64 else 8) * count
    else:
        res["algorithmic_bytes"] = 5 * owned
    # ---- the same step inside the library: count exchange + scan fused into the reduce kernel over
    # peer memory (PFAC_comm, CUDA IPC mailboxes), then the runs placed into rank 0's list by P2P stores
    if api == "reduce64":
        from pfac_b200 import PFACComm
        comm = PFACComm.from_torch(list_capacity=int(total_m) + 16 if rank == 0 else 0, device=dev) if world > 1 \
            else PFACComm(0, 1, int(total_m) + 16)
        d_scan = torch.zeros(3, dtype=torch.int64, device=dev)
        for _ in range(2):
            g_off, g_total, g_count = pf.matchShardFromDeviceReduce64Global(comm, d_in, owned, total, start, d_id, d_pos)
        assert (g_off, g_total, g_count) == (off, total_m, count), ((g_off, g_total, g_count), (off, total_m, count))
        sync()
        ev0.record()
        for _ in range(args.steps):
            pf.matchShardFromDeviceReduce64Global(comm, d_in, owned, total, start, d_id, d_pos, d_scan=d_scan, sync=False)
        ev1.record()
        torch.cuda.synchronize()
        gms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        res["global_scan_ms_per_step"] = float(gms.item())          # kernel + in-kernel exchange, device-timed
        res["global_scan_input_GBps_all_ranks"] = total_len / (float(gms.item()) * 1e-3) / 1e9
        if args.gather:
            sync()
            ev0.record()
            pf.gatherRuns(comm, 0, d_id, d_pos, d_scan=d_scan, sync=False)
            ev1.record()
            torch.cuda.synchronize()
            res["p2p_gather_ms"] = ev0.elapsed_time(ev1)
            sync()
            if rank == 0:   # the list on rank 0: ascending, this rank's run first
                k = min(int(total_m), 1 << 24)
                ids0, pos0 = comm.read_global_list(k)
                okp = bool(np.all(pos0[1:] > pos0[:-1])) if k > 1 else True
                kk = min(k, count)
                okp = okp and np.array_equal(ids0[:kk], d_id[:kk].cpu().numpy()) and \
                    np.array_equal(pos0[:kk], d_pos[:kk].cpu().numpy())
                tail_ids, tail_pos = comm.read_global_list(min(int(total_m), 4096), first=max(int(total_m) - 4096, 0))
                res["p2p_gather_ok"] = bool(okp) and bool(np.all(tail_pos[1:] > tail_pos[:-1]))
                res["p2p_gather_last_pos"] = int(tail_pos[-1]) if tail_pos.size else -1
        sync()
        comm.destroy()
    res["roofline_frac_of_6548.5"] = res["algorithmic_bytes"] / (ms_max * 1e-3) / 1e9 / 6548.5

    # ---- parity vs the oracle, chunked ---------------------------------------------------------------
    check = owned if args.check_bytes < 0 else min(args.check_bytes, owned)
    if check:
        CheckerOracle.set_threads(max(1, (os.cpu_count() or 1) // world))  # torchrun exports OMP_NUM_THREADS=1
        orc = CheckerOracle(pfile)
        nmis = 0
        t0 = time.time()
        k = 0
        for c0, c1, want in ChunkedCheck(orc, d_in, owned, maxlen - 1, limit=check):
            if api == "dense":
                nmis += int((d_out[c0:c1] != torch.from_numpy(want).to(dev)).sum().item())
            else:
                ids, pos = nonzero_pairs(want, c0 + (start if pos64 else 0))
                g_ids = d_id[k:k + ids.size].cpu().numpy()
                g_pos = d_pos[k:k + ids.size].cpu().numpy().astype(np.int64)
                nmis += int(g_ids.size != ids.size) + int((g_ids != ids[:g_ids.size]).sum()) + \
                    int((g_pos != pos[:g_pos.size]).sum())
                k += ids.size
        if api != "dense":
            if check == owned and k != count:
                nmis += abs(k - count) + 1
            res["oracle_matches_checked"] = int(k)
        res["oracle"] = orc.kind
        res["checked_bytes"] = check
        res["mismatches"] = nmis
        res["oracle_s"] = round(time.time() - t0, 1)
        res["bit_exact"] = nmis == 0
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "a") as f:
        f.write(json.dumps(res) + "\n")
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(res) + "\n")
    pf.destroy()
    if world > 1:
        dist.destroy_process_group()
    if check and not res["bit_exact"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
