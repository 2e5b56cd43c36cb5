#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel: joins the SASS page of an ncu
report with the line table of the library's cubin (nvdisasm --print-line-info; build with -lineinfo).
The library must be the build that was profiled.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep <kernel substring> [top N] [--lib path.so]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(lib, kernel_sub):
    td = tempfile.mkdtemp(prefix="pfac_cub_")
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    best = None
    for f in os.listdir(td):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, f)], capture_output=True, text=True).stdout
        cur, line, table = None, 0, {}
        for l in dis.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
            if m:
                cur = m.group(1)
                table[cur] = []
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
            if m:
                line = int(m.group(2))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
            if m and cur:
                table[cur].append((int(m.group(1), 16), line, m.group(2).strip()))
        for k, v in table.items():
            if kernel_sub in k and v and (best is None or len(v) > len(best)):
                best = v
    return best


def main():
    path, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
    lib = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else os.path.join(ROOT, "pfac_b200", "lib", "libpfac.so")
    table = line_table(lib, ksub)
    if not table:
        raise SystemExit("kernel not found in " + lib)
    off2line = {o: ln for o, ln, _ in table}
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, recs, base = None, [], None
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
            d = dict(zip(hdr, r))
            a = int(d["Address"], 16)
            if base is None:
                base = a
            d["_off"] = a - base
            recs.append(d)

    def num(d, k):
        try:
            return float((d.get(k) or "0").replace(",", ""))
        except ValueError:
            return 0.0
    agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for d in recs:
        ln = off2line.get(d["_off"], -1)
        a = agg[ln]
        a[0] += num(d, "Instructions Executed")
        a[1] += num(d, "# Samples")
        for c in stall_cols:
            v = num(d, c)
            if v:
                a[2][c[6:]] += v
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    src = open(os.path.join(ROOT, "pfac_b200", "csrc", "pfac_kernels.cu")).read().splitlines()
    print("kernel %s: %.0f warp instructions, %.0f samples, %d SASS instructions" % (ksub, tot_i, tot_s, len(recs)))
    print("%5s %7s %7s  %-34s %s" % ("line", "inst%", "smpl%", "top stalls", "source"))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        st = " ".join("%s:%d" % (k, 100 * v / max(a[1], 1)) for k, v in a[2].most_common(3))
        text = src[ln - 1].strip()[:90] if 0 < ln <= len(src) else "?"
        print("%5d %6.2f%% %6.2f%%  %-34s %s" % (ln, 100 * a[0] / tot_i, 100 * a[1] / tot_s, st, text))


if __name__ == "__main__":
    main()
