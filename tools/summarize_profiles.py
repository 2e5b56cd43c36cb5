#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py r2 <commit>

Inputs (gpurun_out/): launches_<r>.csv (ncu --metrics gpu__time_duration.sum --csv on bench.py),
traffic_<r>.csv (ncu --metrics dram bytes ... --csv on bench.py).  ncu is not run here.
Outputs: profiles/<r>_launches.csv (per kernel: launches, total and mean device time, share),
profiles/<r>_traffic.csv (per launch), profiles/traffic.json (what bench.py's roofline.traffic reads).
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")


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


def short(name):
    return name.split('(')[0].replace('void ', '').replace('pfac::<unnamed>::', 'pfac::').strip()


def main(tag, commit):
    os.makedirs(OUT, exist_ok=True)
    agg = collections.OrderedDict()
    for d in parse_log_csv(os.path.join(SRC, "launches_%s.csv" % tag)):
        v = float(d['Metric Value'].replace(',', ''))
        v *= {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(d['Metric Unit'], 1)
        a = agg.setdefault(short(d['Kernel Name']), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, "%s_launches.csv" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 5 --warmup 3 "
                "--skip-cpu --skip-e2e --c5-steps 2 (commit %s); per-launch times are cold-cache and serialised: "
                "compare shares\n" % commit)
        f.write("kernel,launches,total_ms,mean_ms,share_of_all_device_time\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.4f,%.4f,%.4f\n" % (k, n, t / 1e6, t / n / 1e6, t / tot))
    # ---- traffic per launch
    per = collections.OrderedDict()
    for d in parse_log_csv(os.path.join(SRC, "traffic_%s.csv" % tag)):
        key = (d['ID'], short(d['Kernel Name']))
        v = float(d['Metric Value'].replace(',', ''))
        unit = d['Metric Unit']
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1, 'us': 1e3, 'ms': 1e6}.get(unit, 1)
        per.setdefault(key, {})[d['Metric Name']] = v
    cols = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'gpu__time_duration.sum']
    dense, reduce_ = [], []
    with open(os.path.join(OUT, "%s_traffic.csv" % tag), "w") as f:
        f.write("# ncu per-launch metrics of bench.py's kernels on the 1 GiB C2 shard (commit %s)\n" % commit)
        f.write("id,kernel," + ",".join(cols) + "\n")
        for (i, k), m in per.items():
            f.write("%s,%s,%s\n" % (i, k, ",".join("%.0f" % m.get(c, float('nan')) for c in cols)))
            b = m.get(cols[0], 0) + m.get(cols[1], 0)
            (dense if 'dense' in k else reduce_).append((k, b, m.get('gpu__time_duration.sum', 0)))
    out = {"round": tag, "commit": commit, "source": "profiles/%s_traffic.csv (ncu --metrics dram__bytes_read.sum,"
           "dram__bytes_write.sum per launch of bench.py on the 1 GiB C2 shard)" % tag}
    if dense:
        out["kernel"] = dense[0][0]
        out["dense_dram_bytes_per_launch"] = sum(b for _, b, _ in dense) / len(dense)
        out["dense_ncu_ms_per_launch"] = sum(t for _, _, t in dense) / len(dense) / 1e6
    if reduce_:
        out["reduce_kernel"] = reduce_[0][0]
        out["reduce_dram_bytes_per_launch"] = sum(b for _, b, _ in reduce_) / len(reduce_)
        out["reduce_ncu_ms_per_launch"] = sum(t for _, _, t in reduce_) / len(reduce_) / 1e6
    json.dump(out, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?")
