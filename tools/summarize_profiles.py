#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py r1

Inputs (gpurun_out/): launches_<r>.csv (ncu --metrics gpu__time_duration.sum --csv),
traffic_<r>.csv (ncu --metrics dram bytes ... --csv), prof_dense_<r>.ncu-rep and
prof_reduce_<r>.ncu-rep (ncu --set full).  ncu is run here only to read the reports (-i).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size']


def parse_log_csv(path):
    rows = list(csv.reader(open(path)))
    hdr, out = None, []
    for r in rows:
        if r and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            out.append(dict(zip(hdr, r)))
    return out


def to_bytes(v):
    x = float(v[0].replace(',', ''))
    return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[v[1]]


def main(tag):
    os.makedirs(OUT, exist_ok=True)
    # ---- launch list
    agg = collections.OrderedDict()
    for d in parse_log_csv(os.path.join(SRC, "launches_%s.csv" % tag)):
        short = d['Kernel Name'].split('(')[0].replace('void ', '').replace('pfac::<unnamed>::', 'pfac::')
        v = float(d['Metric Value'].replace(',', ''))
        v *= {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(d['Metric Unit'], 1)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = ["# ncu launch list (%s)" % tag,
             "# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 80 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-e2e",
             "# (first 80 launches; per-launch times are cold-cache and serialised: compare shares)",
             "kernel,launches,total_ns,avg_ns,share"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append("%s,%d,%.0f,%.0f,%.4f" % (k[:90], n, t, t / n, t / tot))
    open(os.path.join(OUT, "%s_launches.csv" % tag), "w").write("\n".join(lines) + "\n")
    # ---- traffic
    by = collections.OrderedDict()
    for d in parse_log_csv(os.path.join(SRC, "traffic_%s.csv" % tag)):
        by.setdefault((d['ID'], d['Kernel Name'].split('(')[0].replace('void ', '')), {})[d['Metric Name']] = (
            d['Metric Value'], d['Metric Unit'])
    tl = ["# per-launch DRAM traffic etc. on the full 1 GiB bench launch (%s)" % tag,
          "# command: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,"
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum "
          "--clock-control none -k regex:pfac_ -s 3 -c 4 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --reduce-steps 1"]
    dense_b = red_b = None
    for (i, k), m in by.items():
        tl.append("%s,%s," % (i, k) + ",".join("%s=%s %s" % (n, v[0], v[1]) for n, v in m.items()))
        b = to_bytes(m['dram__bytes_read.sum']) + to_bytes(m['dram__bytes_write.sum'])
        if 'dense' in k and dense_b is None:
            dense_b = b
        if 'reduce' in k:
            red_b = b
    open(os.path.join(OUT, "%s_traffic.csv" % tag), "w").write("\n".join(tl) + "\n")
    json.dump({"dense_dram_bytes_per_launch": dense_b, "reduce_dram_bytes_per_launch": red_b,
               "source": "profiles/%s_traffic.csv (ncu, 1 GiB C2 workload)" % tag},
              open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    # ---- full captures
    for k, cmd in (("dense", "python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --skip-reduce --bytes 268435456"),
                   ("reduce", "python tools/reduce_stress.py 256")):
        rep = os.path.join(SRC, "prof_%s_%s.ncu-rep" % (k, tag))
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, r = rows[0], rows[1], rows[2]
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        out = ["# ncu --set full, %s kernel (%s)" % (k, tag), "",
               "Command: `ncu --set full --clock-control none --import-source on -k regex:pfac_%s -c 1 %s`" % (k, cmd),
               "(256 MiB of the C2 workload)", "", "## " + d['Kernel Name'], "", "| metric | value |", "|---|---|"]
        for m in KEYS:
            if m in d:
                out.append("| `%s` | %s %s |" % (m, d[m], u[m]))
        st = {x: float(v.replace(',', '')) for x, v in d.items()
              if x.startswith('smsp__pcsamp_warps_issue_stalled_') and not x.endswith('not_issued')}
        tots = sum(st.values()) or 1
        out.append("")
        out.append("warp-stall samples: " + ", ".join(
            "%s %.1f%%" % (x[33:], 100 * v / tots) for x, v in sorted(st.items(), key=lambda z: -z[1])[:8]))
        open(os.path.join(OUT, "%s_%s_full.md" % (tag, k)), "w").write("\n".join(out) + "\n")
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
