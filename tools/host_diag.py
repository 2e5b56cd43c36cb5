#!/usr/bin/env python
"""Host-side facts that bound PFAC_matchFromHost on the GPU box: cores, NUMA layout, where each GPU
hangs, zero-fill bandwidth of the copy pool by thread count, and pinned H2D bandwidth.

    python tools/host_diag.py            (writes gpurun_out/host_diag.json)
"""
import glob
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read(path):
    try:
        return open(path).read().strip()
    except Exception:
        return None


def main():
    out = {"cpu_count": os.cpu_count(), "affinity": sorted(os.sched_getaffinity(0))}
    out["nodes"] = {os.path.basename(p): {"cpulist": read(p + "/cpulist"),
                                          "meminfo": (read(p + "/meminfo") or "").splitlines()[:2]}
                    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))}
    out["cpuinfo_model"] = [l for l in (read("/proc/cpuinfo") or "").splitlines() if "model name" in l][:1]
    out["cgroup_cpuset"] = read("/sys/fs/cgroup/cpuset.cpus.effective") or read("/sys/fs/cgroup/cpuset/cpuset.cpus")
    out["cgroup_cpu_max"] = read("/sys/fs/cgroup/cpu.max")
    try:
        out["nvidia_smi_topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True,
                                                timeout=30).stdout
    except Exception as e:
        out["nvidia_smi_topo"] = repr(e)
    import torch
    out["gpus"] = torch.cuda.device_count()
    gp = []
    for i in range(torch.cuda.device_count()):
        pr = torch.cuda.get_device_properties(i)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        gp.append({"index": i, "bdf": bdf, "numa_node": read("/sys/bus/pci/devices/%s/numa_node" % bdf),
                   "local_cpulist": read("/sys/bus/pci/devices/%s/local_cpulist" % bdf),
                   "l2": pr.L2_cache_size})
    out["gpu_pci"] = gp

    # zero-fill bandwidth by thread count (one subprocess per setting: the pool is per process)
    fills = {}
    for th in (4, 8, 12, 16, 24, 32):
        if th > (os.cpu_count() or 1):
            continue
        env = dict(os.environ, PFAC_B200_COPY_THREADS=str(th))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "host_zero_bw.py")], env=env,
                           capture_output=True, text=True, timeout=300)
        fills[th] = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
    out["zero_fill"] = fills

    # pinned H2D / D2H bandwidth, 1 GiB
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.zero_()
    d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        best = 0
        for _ in range(3):
            t = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = max(best, n / (time.perf_counter() - t) / 1e9)
        out["pinned_%s_GBps" % name] = best
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "host_diag.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
