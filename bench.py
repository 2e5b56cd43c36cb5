#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE configs[1] = 1,000 synthetic patterns (len 4-32, 255-symbol
alphabet) over 1 GiB of planted random text per GPU, PFAC_matchFromDevice (dense int32 result).
One step = one pass of the hot path over the rank's resident 1 GiB shard (+ tail halo).  N > 1
is weak scaling: rank r owns bytes [r GiB, (r+1) GiB) of an N GiB stream, no data-path
collective.  value = total input GB (1e9 B) scanned by all ranks per second of the slowest rank.

Extra objects on the JSON line: roofline (dominant kernel vs the measured HBM peak),
cpu_baseline (reference CPU_OMP matcher on this box's host cores, rank 0, N=1; plus the reference's
single-thread PFAC_CPU on a smaller sample), e2e (same metric through PFAC_matchFromHost with pinned
host buffers, copies inside the timed region), reduce (the fused compaction path on the same shard with
the cross-GPU count scan done inside the kernel over peer memory), c5 (BASELINE configs[4]: 32 GiB of
text regenerated on the GPUs, 10,000 patterns, sharded reduce + in-kernel global offset scan + the runs
placed into one list by P2P stores, 512 MiB per rank checked against the oracle outside the timed region),
table (what the table compiler built).  `config` is identical for both arms.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from workloads import synth  # noqa: E402

GIB = 1 << 30
METRIC = "input_GB_per_s_scanned"
UNIT = "GB/s"
TEXT_SEED = synth.SEED_BASE + 2
N_PATTERNS = 1000


def workload_config(bytes_per_gpu, n_gpus):
    return {
        "workload": "C2: %d synthetic patterns (len 4-32, 255-symbol alphabet, 50 prefix pairs) over "
                    "%.3f GiB planted random text per GPU, PFAC_matchFromDevice dense int32 result"
                    % (N_PATTERNS, bytes_per_gpu / GIB),
        "patterns": N_PATTERNS,
        "bytes_per_gpu": bytes_per_gpu,
        "total_bytes": bytes_per_gpu * n_gpus,
        "plant_every": 4096,
        "sharding": "contiguous shards + (maxPatternLen-1)-byte tail halo, no data-path collective",
        "l2": "5 B/position of traffic per step (>= 5 GiB) is far larger than the 126 MB L2: no flush",
    }


def make_shard(rank, world, bytes_per_gpu, patterns):
    """Bytes [rank*B, (rank+1)*B + halo) of the world*B-byte stream, generated in 64 MiB pieces."""
    total_len = bytes_per_gpu * world
    halo = max(len(p) for p in patterns) - 1
    start = rank * bytes_per_gpu
    n = min(bytes_per_gpu + halo, total_len - start)
    out = np.empty(n, dtype=np.uint8)
    piece = 64 << 20
    for off in range(0, n, piece):
        m = min(piece, n - off)
        out[off:off + m] = synth.random_bytes(TEXT_SEED, start + off, m)
    synth.plant(out, start, total_len, patterns, TEXT_SEED, every=4096)
    return out, min(bytes_per_gpu, n)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.period = period
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this command
    (profiles/traffic.json: per-launch dram__bytes_read.sum + dram__bytes_write.sum, the kernel it was
    taken on and the commit).  Returns (bytes, provenance)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return d.get("dense_dram_bytes_per_launch"), {k: d.get(k) for k in ("kernel", "commit", "round", "source")}
    except Exception:
        return None, None


def traffic_mix_reference():
    """What plain kernels reach with the dense kernel's 1:4 read:write mix on these GPUs (tools/micro/hbm_mix.cu,
    committed run): context for `frac`, whose denominator is the 1:1 copy bandwidth."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_hbm_mix.json")))
        return {"plain_kernel_read_1GiB_write_4GiB_GBps": d.get("mix_rows_GBps"), "cudaMemset_4GiB_GBps": d.get("memset_GBps"),
                "copy_kernel_GBps": d.get("copy_GBps"), "source": "profiles/r2_hbm_mix.json (tools/micro/hbm_mix.cu)"}
    except Exception:
        return None


def cpu_reference(pfile, text, threads, reps, budget_s=None):
    """Times the reference CPU_OMP matcher (oracle/_ref when built, else the oracle port)."""
    import oracle
    if oracle.ref_available():
        cls, kind = oracle.RefOracle, "reference"
    else:
        cls, kind = oracle.Oracle, "port"
    cls.set_threads(threads)
    m = cls(pfile)
    best = None
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        m.match(text, omp=threads > 1)
        dt = time.perf_counter() - t0
        times.append(dt)
        best = dt if best is None else min(best, dt)
        if budget_s is not None and sum(times) > budget_s:
            break
    return kind, cls.threads(), best, times


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on host cores."""
    if rank != 0:
        return
    patterns = synth.patterns_c2(N_PATTERNS)
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as td:
        pfile = synth.write_pattern_file(os.path.join(td, "c2.pat"), patterns)
        # size the per-step sample so that (warmup + steps) passes end within ~2 minutes
        probe_n = 32 << 20
        probe, _ = make_shard(0, world, probe_n, patterns)
        kind, threads, best, _ = cpu_reference(pfile, probe[:probe_n], cores, 1)
        rate = probe_n / best  # B/s
        want = int(rate * 120.0 / max(1, args.steps + args.warmup))
        sample_n = max(16 << 20, min(args.bytes, want, GIB))
        sample_n -= sample_n % 4096
        text, _ = make_shard(0, world, sample_n, patterns)
        text = text[:sample_n]
        import oracle
        cls = oracle.RefOracle if oracle.ref_available() else oracle.Oracle
        cls.set_threads(cores)
        m = cls(pfile)
        for _ in range(args.warmup):
            m.match(text, omp=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m.match(text, omp=True)
        dt = time.perf_counter() - t0
    value = sample_n * args.steps / dt / 1e9
    sample = "first %.0f MiB of the rank-0 shard per step, PFAC_CPU_OMP_timeDriven, %d OpenMP threads" % (
        sample_n / (1 << 20), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(args.bytes, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


C5_PATTERNS = 10000
C5_TOTAL = 32 * GIB


def run_c5(pf_factory, rank, local_rank, world, dev, dist, torch, barrier, max_over_ranks, steps, total_len, peak):
    """BASELINE configs[4]: `total_len` bytes of ASCII-weighted text regenerated on the GPUs from the
    counter hash (no host buffer, no H2D), 10,000 Snort-like patterns, contiguous shards with a
    (maxPatternLen-1)-byte tail halo; per step every rank runs the fused reduce kernel, which publishes
    its match count into every rank's mailbox over NVLink peer memory and takes the exclusive prefix
    itself (PFAC_comm), then the runs are placed into one global list on rank 0 by P2P stores.  Device
    timed, max over ranks.  Parity (outside the timed region): the first 512 MiB of every shard against
    the oracle, global 64-bit positions, offsets = exclusive scan of the counts."""
    from pfac_b200 import PFACComm
    from pfac_b200.sharding import shard_bounds
    from tests import configs
    cfg = dict(configs.CONFIGS["c5"])
    pats = cfg["patterns"]()
    tmp = tempfile.mkdtemp(prefix="pfac_c5_")
    pfile = synth.write_pattern_file(os.path.join(tmp, "c5_rank%d.pat" % rank), pats)
    pf = pf_factory()
    pf.readPatternFromFile(pfile)
    info = pf.tableInfo(reduce=True)
    maxlen = info["max_pattern_len"]
    start, owned, total = shard_bounds(total_len, world, rank, maxlen)
    t0 = time.perf_counter()
    d_in = configs.device_text(cfg, start, total, total_len, pats, dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    cap = max(owned // 32, 1 << 20)           # ~0.9 % of the positions match; 3 % holds them
    d_id = torch.empty(cap, dtype=torch.int32, device=dev)
    d_pos = torch.empty(cap, dtype=torch.int64, device=dev)
    d_scan = torch.zeros(3, dtype=torch.int64, device=dev)
    # the global list lives on rank 0 (every rank maps it); sized after a first pass tells the total
    comm0 = PFACComm.from_torch(0, device=dev) if world > 1 else PFACComm(0, 1, 0)
    off, total_m, count = pf.matchShardFromDeviceReduce64Global(comm0, d_in, owned, total, start, d_id, d_pos)
    barrier()
    comm0.destroy()
    comm = (PFACComm.from_torch(total_m + 64 if rank == 0 else 0, device=dev) if world > 1
            else PFACComm(0, 1, total_m + 64))

    def step(gather):
        pf.matchShardFromDeviceReduce64Global(comm, d_in, owned, total, start, d_id, d_pos, d_scan=d_scan, sync=False)
        if gather:
            pf.gatherRuns(comm, 0, d_id, d_pos, d_scan=d_scan, sync=False)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(gather):
        for _ in range(3):
            step(gather)
        barrier()
        ev[0].record()
        for _ in range(steps):
            step(gather)
        ev[1].record()
        torch.cuda.synchronize()
        return max_over_ranks(ev[0].elapsed_time(ev[1]) / steps)

    ms = timed(False)          # the config: sharded reduce + global offset scan
    ms_gather = timed(True)    # + the optional second step: every run stored into one list on rank 0
    # the same kernel with nobody to wait for (a one-rank comm) and no placement: what the cross-GPU
    # part of the step costs is the difference
    solo = PFACComm(0, 1, 0)
    d_scan1 = torch.zeros(3, dtype=torch.int64, device=dev)
    for _ in range(2):
        pf.matchShardFromDeviceReduce64Global(solo, d_in, owned, total, start, d_id, d_pos, d_scan=d_scan1, sync=False)
    barrier()
    ev[0].record()
    for _ in range(steps):
        pf.matchShardFromDeviceReduce64Global(solo, d_in, owned, total, start, d_id, d_pos, d_scan=d_scan1, sync=False)
    ev[1].record()
    torch.cuda.synchronize()
    ms_plain = max_over_ranks(ev[0].elapsed_time(ev[1]) / steps)
    solo.destroy()
    barrier()
    scan = [int(x) for x in d_scan.cpu().tolist()]
    assert scan == [off, total_m, count], (scan, off, total_m, count)
    # ---- parity, outside the timed region
    from tests.helpers import CheckerOracle
    check = min(owned, 512 << 20)
    CheckerOracle.set_threads(max(1, (os.cpu_count() or 1) // world))
    orc = CheckerOracle(pfile)
    mism, k = 0, 0
    t0 = time.perf_counter()
    for c0, c1, want in configs.ChunkedCheck(orc, d_in, owned, maxlen - 1, limit=check):
        ids, pos = configs.nonzero_pairs(want, c0 + start)
        g_ids = d_id[k:k + ids.size].cpu().numpy()
        g_pos = d_pos[k:k + ids.size].cpu().numpy()
        mism += int(g_ids.size != ids.size) + int((g_ids != ids[:g_ids.size]).sum()) + int((g_pos != pos[:g_pos.size]).sum())
        k += ids.size
    oracle_s = time.perf_counter() - t0
    # offsets: exclusive scan of the counts, checked against an NCCL all-gather of the same counts
    counts = torch.tensor([count], dtype=torch.int64, device=dev)
    if world > 1:
        allc = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, counts)
    else:
        allc = counts
    allc = [int(x) for x in allc.cpu().tolist()]
    ok_scan = off == sum(allc[:rank]) and total_m == sum(allc)
    # the global list on rank 0: this rank's run sits at its offset; ascending over rank boundaries
    ok_list = True
    if rank == 0:
        kk = min(count, 1 << 22)
        ids0, pos0 = comm.read_global_list(kk)
        ok_list = bool(np.array_equal(ids0, d_id[:kk].cpu().numpy()) and np.array_equal(pos0, d_pos[:kk].cpu().numpy()))
        lo = max(count - 2048, 0)
        seam_ids, seam_pos = comm.read_global_list(min(total_m - lo, 4096), first=lo)
        ok_list = ok_list and bool(np.all(seam_pos[1:] > seam_pos[:-1]))
        tail_ids, tail_pos = comm.read_global_list(min(total_m, 4096), first=max(total_m - 4096, 0))
        ok_list = ok_list and bool(np.all(tail_pos[1:] > tail_pos[:-1])) and int(tail_pos[-1]) < total_len
    flags = torch.tensor([mism, 0 if (ok_scan and ok_list) else 1], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(flags)
    bit_exact = int(flags[0].item()) == 0 and int(flags[1].item()) == 0
    barrier()
    comm.destroy()
    pf.destroy()
    algo = owned + 12 * count
    return {
        "workload": "C5: %d Snort-like patterns over %.0f GiB ASCII-weighted planted text regenerated on the GPUs, "
                    "%d contiguous shards + %d-byte tail halo, PFAC_matchShardFromDeviceReduce64Global (fused match + "
                    "compaction + in-kernel cross-GPU count exchange and exclusive scan)"
                    % (len(pats), total_len / GIB, world, maxlen - 1),
        "value": total_len / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": steps,
        "per_gpu_GBps": owned / (ms * 1e-3) / 1e9, "bytes_per_gpu": owned, "total_bytes": total_len,
        "kernel_only_ms_per_call": ms_plain,
        "count_scan_ms_per_step": max(ms - ms_plain, 0.0),
        "cross_gpu": "count exchange + exclusive scan inside the reduce kernel over NVLink peer memory; no NCCL call, "
                     "no host round trip in the step",
        "with_gather": {"value": total_len / (ms_gather * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_gather,
                        "gather_ms_per_step": max(ms_gather - ms, 0.0), "gather_bytes_into_rank0": 12 * (total_m - count) if rank == 0 else None,
                        "what": "PFAC_commGatherRuns after every step: every rank's run stored into ONE list on rank 0 "
                                "at its scanned offset by P2P stores (12 B per match over NVLink into one GPU)"},
        "matches_total": total_m, "matches_rank0": count if rank == 0 else None, "states": info["num_states"],
        "roofline": {"bound": "hbm", "kernel": "pfac_reduce_kernel<1, 8, %d>" % ({0: 0, 1: 2, 2: 3, 3: 5}[info["hashed_filter"]]),
                     "algorithmic_bytes_per_launch": int(algo), "achieved": algo / (ms_plain * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": algo / (ms_plain * 1e-3) / 1e9 / peak,
                     "note": "rank 0, N + 12 M bytes per launch; issue-bound, not HBM-bound"},
        "gen_s": gen_s, "oracle": orc.kind, "oracle_checked_bytes_per_rank": check, "oracle_s": oracle_s,
        "bit_exact": bit_exact,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bytes", type=int, default=GIB, help="input bytes per GPU")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--reduce-steps", type=int, default=10)
    ap.add_argument("--c5-steps", type=int, default=5)
    ap.add_argument("--c5-bytes", type=int, default=C5_TOTAL, help="total bytes of the C5 leg (all ranks)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-reduce", action="store_true")
    ap.add_argument("--skip-c5", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # NCCL prints its version banner / debug lines to stdout: keep stdout for the one JSON line by
    # pointing fd 1 at stderr while the ranks run and printing the line to the saved stdout
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    # the host copy pool is per process: share the box's cores among the ranks of this job
    os.environ.setdefault("PFAC_B200_COPY_THREADS", str(max(2, min(8, (os.cpu_count() or 8) // max(world, 1)))))
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from pfac_b200 import PFAC, PFACComm
    from pfac_b200.api import kernel_launch_count

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the library has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    patterns = synth.patterns_c2(N_PATTERNS)
    tmpdir = tempfile.mkdtemp(prefix="pfac_bench_")
    pfile = synth.write_pattern_file(os.path.join(tmpdir, "c2_rank%d.pat" % rank), patterns)
    shard, owned = make_shard(rank, world, args.bytes, patterns)
    total = shard.size

    pf = PFAC()
    pf.readPatternFromFile(pfile)
    info = pf.tableInfo()
    d_in = torch.from_numpy(shard).to(dev)
    d_out = torch.empty(owned, dtype=torch.int32, device=dev)

    def step():
        pf.matchShardFromDevice(d_in, owned, total, d_out)

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = kernel_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    launches = kernel_launch_count() - launches0
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    worst_ms = max_over_ranks(total_ms)
    ms_per_step = worst_ms / args.steps
    value = owned * world / (ms_per_step * 1e-3) / 1e9
    sampler.join(timeout=1.0)
    clocks = sampler.result()

    # result digest (cheap proof that the timed kernel did the work)
    n_matches = int((d_out != 0).sum().item())

    # ---- roofline of the dominant kernel (this rank) -------------------------------------------
    peak, peak_src = measured_peak()
    avg_launch_ms = float(np.mean(per_launch_ms))
    algo_bytes = 5 * owned  # 1 B read + 4 B written per position
    achieved = algo_bytes / (avg_launch_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic()
    filt = {1: 2, 2: 3, 3: 5}[info["hashed_filter"]] if info["hashed_filter"] else (1 if info["has_chk2"] else 0)
    kernel_name = "pfac_dense_kernel<%d, %d>" % (info["code_bits"], filt)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kernel_name,
                "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": avg_launch_ms,
                "median_launch_ms": float(np.median(per_launch_ms)), "best_launch_ms": float(np.min(per_launch_ms)),
                "frac_of_8TBps_spec": achieved / 8000.0,
                "traffic_mix_reference": traffic_mix_reference()}

    # ---- end to end through the host-buffer API ---------------------------------------------------
    e2e = None
    if not args.skip_e2e:
        h_in = torch.empty(owned, dtype=torch.uint8, pin_memory=True)
        h_in.numpy()[:] = shard[:owned]
        h_out = torch.empty(owned, dtype=torch.int32, pin_memory=True)

        def time_host_calls():
            pf.matchFromHost(h_in, h_out, size=owned)  # warm-up: allocates the pipeline buffers
            h_out.fill_(-1)                            # every timed call must rewrite the whole array
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                pf.matchFromHost(h_in, h_out, size=owned)
            mine = time.perf_counter() - t0
            return max_over_ranks(mine), mine

        dt, mine = time_host_calls()
        h2d_b, d2h_b = pf.lastHostTransfer()           # counted by the library around its cudaMemcpyAsync calls
        e2e = {"value": owned * world * args.e2e_steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
               "steps": args.e2e_steps, "per_rank_value": owned * args.e2e_steps / mine / 1e9,
               "sum_of_rank_rates": sum_over_ranks(owned * args.e2e_steps / mine / 1e9),
               "host_threads_per_rank": int(os.environ["PFAC_B200_COPY_THREADS"]),
               "api": "PFAC_matchFromHost (pinned host buffers; chunked H2D / fused match+compaction / D2H of the "
                      "(id, position) pairs; the host zero-fills and scatters the dense int32 array it returns)"}
        if world == 1:
            # the shard ends at the end of the stream, so the host call and the shard call agree
            assert bool((torch.from_numpy(h_out.numpy()).to(dev) == d_out).all().item()), "e2e result differs"
        # the same call with the dense array itself crossing PCIe (the reference's data movement)
        os.environ["PFAC_B200_HOST_RESULT"] = "dense"
        dt, _ = time_host_calls()
        del os.environ["PFAC_B200_HOST_RESULT"]
        h2d_b, d2h_b = pf.lastHostTransfer()
        e2e["dense_d2h"] = {"value": owned * world * args.e2e_steps / dt / 1e9, "unit": UNIT,
                            "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
                            "api": "PFAC_matchFromHost with PFAC_B200_HOST_RESULT=dense"}
        if world == 1:
            assert bool((torch.from_numpy(h_out.numpy()).to(dev) == d_out).all().item()), "e2e (dense D2H) result differs"
        del h_out
        # the same host-buffer input through the reduced API: 8 bytes per match come back instead of
        # 4 bytes per input byte (extra information; the contract's e2e is the dense call above)
        if owned < 2 ** 31:
            r_id = np.empty(owned, dtype=np.int32)    # pageable: only the first M entries are touched
            r_pos = np.empty(owned, dtype=np.int32)
            ids, pos = pf.matchFromHostReduce(h_in, r_id, r_pos, size=owned)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                ids, pos = pf.matchFromHostReduce(h_in, r_id, r_pos, size=owned)
            dt = max_over_ranks(time.perf_counter() - t0)
            e2e["reduce_api"] = {"value": owned * world * args.e2e_steps / dt / 1e9, "unit": UNIT,
                                 "h2d_bytes_per_step": int(owned), "d2h_bytes_per_step": int(ids.size) * 8,
                                 "api": "PFAC_matchFromHostReduce", "matches": int(ids.size)}
            del r_id, r_pos
        del h_in
        pf.releaseHostBuffers()

    # ---- fused reduce path on the same shard, the cross-GPU count scan inside the kernel -----------------
    reduce_info = None
    if not args.skip_reduce:
        cap = max(owned // 16, 1 << 20)
        d_id = torch.empty(cap, dtype=torch.int32, device=dev)
        d_pos = torch.empty(cap, dtype=torch.int64, device=dev)
        d_scan = torch.zeros(3, dtype=torch.int64, device=dev)
        base = rank * args.bytes
        comm = PFACComm.from_torch(0, device=dev) if world > 1 else PFACComm(0, 1, 0)
        for _ in range(3):  # warm-up: workspace allocation
            my_off, total_m, m = pf.matchShardFromDeviceReduce64Global(comm, d_in, owned, total, base, d_id, d_pos)
        barrier()
        # (a) synchronous call, the exclusive offsets returned to the host: one stream sync per step
        t0 = time.perf_counter()
        for _ in range(args.reduce_steps):
            my_off, total_m, m = pf.matchShardFromDeviceReduce64Global(comm, d_in, owned, total, base, d_id, d_pos)
        dt = max_over_ranks(time.perf_counter() - t0)
        # (b) the same enqueued without a host sync (scan kept on the device), device-timed
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reduce_steps):
            pf.matchShardFromDeviceReduce64Global(comm, d_in, owned, total, base, d_id, d_pos, d_scan=d_scan, sync=False)
        e1.record()
        torch.cuda.synchronize()
        ms_global = max_over_ranks(e0.elapsed_time(e1) / args.reduce_steps)
        # (c) the kernel without the exchange (plain shard call, synchronous), for the cost of the scan
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.reduce_steps):
            m2 = pf.matchShardFromDeviceReduce64Cap(d_in, owned, total, base, d_id, d_pos)
        t_call = max_over_ranks(time.perf_counter() - t0)
        barrier()
        comm.destroy()
        assert m == n_matches and m2 == m, "reduce count %d != dense non-zeros %d" % (m, n_matches)
        assert [int(x) for x in d_scan.cpu().tolist()] == [my_off, total_m, m]
        # spot-check order and content against the dense result of the timed kernel
        pos_local = d_pos[:m] - base
        assert bool((pos_local[1:] > pos_local[:-1]).all().item()), "positions not ascending"
        assert bool((d_out[pos_local] == d_id[:m]).all().item()), "reduce ids differ from dense result"
        reduce_info = {"value": owned * world * args.reduce_steps / dt / 1e9, "unit": UNIT,
                       "api": "PFAC_matchShardFromDeviceReduce64Global (fused match + compaction + in-kernel count "
                              "exchange and exclusive scan over peer memory; synchronous form: offsets returned to the host)",
                       "device_timed_value": owned * world / (ms_global * 1e-3) / 1e9,
                       "device_timed_ms_per_step": ms_global,
                       "call_only_value": owned * world * args.reduce_steps / t_call / 1e9,
                       "ms_per_call": t_call / args.reduce_steps * 1e3,
                       "matches_total": total_m, "rank0_offset": my_off if rank == 0 else None,
                       "count_scan_ms_per_step": max(dt / args.reduce_steps * 1e3 - t_call / args.reduce_steps * 1e3, 0.0),
                       "algorithmic_bytes_per_step": int(owned + 12 * m), "steps": args.reduce_steps,
                       "kernel": "pfac_reduce_kernel<1, 8, %d>" % ({0: 0, 1: 2, 2: 3, 3: 5}[pf.tableInfo(reduce=True)["hashed_filter"]]),
                       "roofline_frac": (owned + 12 * m) / (ms_global * 1e-3) / 1e9 / peak}
        del d_id, d_pos

    # ---- reference CPU path beside it (rank 0, N=1 only) ---------------------------------------------
    cpu = None
    if world == 1 and rank == 0 and not args.skip_cpu:
        cores = os.cpu_count() or 1
        sample_n = min(owned, GIB)
        kind, threads, best, times = cpu_reference(pfile, shard[:sample_n], cores, 3, budget_s=30.0)
        cpu = {"value": sample_n / best / 1e9, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "first %.0f MiB of the shard, best of %d passes of PFAC_CPU_OMP_timeDriven, %d threads"
                         % (sample_n / (1 << 20), len(times), threads)}
        one_n = min(owned, 64 << 20)   # the reference's single-thread PFAC_CPU
        _, _, best1, _ = cpu_reference(pfile, shard[:one_n], 1, 1)
        cpu["single_thread"] = {"value": one_n / best1 / 1e9, "unit": UNIT, "cores": 1,
                                "sample": "first %.0f MiB of the shard, one pass of PFAC_CPU_timeDriven" % (one_n / (1 << 20))}

    del d_in, d_out
    torch.cuda.empty_cache()
    pf.destroy()

    # ---- BASELINE configs[4]: the sharded 32 GiB run with the global offset scan ------------------------
    c5 = None
    if not args.skip_c5:
        c5 = run_c5(PFAC, rank, local_rank, world, dev, dist, torch, barrier, max_over_ranks, args.c5_steps,
                    args.c5_bytes, peak)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.bytes, world),
            "table": {"states": info["num_states"], "hot_depth": info["hot_depth"], "hot_buckets": info["hot_buckets"],
                      "first_stage": ("pair filter: one lookup per two start positions, keyed by the three bytes they "
                                      "share, %d of %d bits set" % (info["hfilt_bits_set"], 32 * info["hfilt_words"])
                                      if info["hashed_filter"] == 3 else
                                      "hashed 4-gram filter, %d bit(s) per lookup, %d of %d bits set"
                                      % (info["hashed_filter"], info["hfilt_bits_set"], 32 * info["hfilt_words"])
                                      if info["hashed_filter"] else
                                      "exact 2-gram set, %d of 65536 bits set" % info["pre2_bits_set"]),
                      "device_bytes": info["device_bytes"], "matches_per_gpu": n_matches},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "reduce": reduce_info, "c5": c5,
            "gpu_launches": int(launches), "clocks": clocks,
            "gbps_reference_unit": value * 8.0,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
