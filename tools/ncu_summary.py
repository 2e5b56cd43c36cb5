#!/usr/bin/env python
"""Key metrics + warp-stall breakdown of an ncu report (ncu -i, no GPU needed).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--md profiles/x.md "title / command"]
"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum', 'lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum',
        'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
        'launch__block_size']


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    return [(dict(zip(hdr, r)), dict(zip(hdr, units))) for r in rows[2:]]


def main():
    path = sys.argv[1]
    lines = []
    for vals, units in load(path):
        lines.append("## %s" % vals.get("Kernel Name", "?"))
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for k in KEYS:
            if k in vals:
                lines.append("| `%s` | %s %s |" % (k, vals[k], units.get(k, "")))
        stalls = []
        for k, v in vals.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") \
                    and "not_issued" not in k:
                try:
                    stalls.append((float(v.replace(",", "")), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1.0
        stalls.sort(reverse=True)
        lines.append("")
        lines.append("warp-stall samples: " + ", ".join("%s %.1f%%" % (n, 100 * s / tot) for s, n in stalls[:9]))
        lines.append("")
    text = "\n".join(lines)
    if "--md" in sys.argv:
        i = sys.argv.index("--md")
        with open(sys.argv[i + 1], "w") as f:
            f.write("# %s\n\n" % sys.argv[i + 2] + text + "\n")
    print(text)


if __name__ == "__main__":
    main()
