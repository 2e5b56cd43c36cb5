#!/usr/bin/env python
"""Append the result lines a gpurun call brought back (gpurun_out/configs.jsonl, host_path.jsonl) to the
tracked, append-only profiles/r2_configs.jsonl, tagged with the commit they were measured on.

    python tools/keep_results.py [note]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(ROOT, "profiles", "r2_configs.jsonl")
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
dirty = bool(subprocess.run(["git", "status", "--porcelain", "--untracked-files=no"], cwd=ROOT, capture_output=True, text=True).stdout.strip())
seen = set(open(dst).read().splitlines()) if os.path.exists(dst) else set()
n = 0
with open(dst, "a") as out:
    for name in ("configs.jsonl", "host_path.jsonl"):
        src = os.path.join(ROOT, "gpurun_out", name)
        if not os.path.exists(src):
            continue
        for line in open(src):
            line = line.strip()
            if not line:
                continue
            d = json.loads(line)
            d["round"] = 2
            d["commit"] = commit + ("+" if dirty else "")
            if len(sys.argv) > 1:
                d["note"] = sys.argv[1]
            s = json.dumps(d)
            if s not in seen:
                out.write(s + "\n")
                seen.add(s)
                n += 1
print("appended", n, "lines to", dst)
