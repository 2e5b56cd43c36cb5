#!/bin/bash
# A/B runs of library variants on one GPU box: tools/ab_variants.sh "<variant names>" "<run_configs args>" ...
# Variants are build/variants/libpfac_<name>.so (built here with extra -D flags; build/ travels with the
# gpurun snapshot); "v1" = the base library with PFAC_B200_DENSE=v1.  Output: gpurun_out/ab.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
variants="$1"; shift
cp pfac_b200/lib/libpfac.so /tmp/libpfac_keep.so
for v in $variants; do
  for args in "$@"; do
    if [ "$v" = "v1" ]; then
      cp build/variants/libpfac_base.so pfac_b200/lib/libpfac.so
      out=$(PFAC_B200_DENSE=v1 timeout 300 python tests/run_configs.py $args 2>&1 | tail -1)
    else
      cp build/variants/libpfac_$v.so pfac_b200/lib/libpfac.so
      out=$(timeout 300 python tests/run_configs.py $args 2>&1 | tail -1)
    fi
    echo "$v | $args | $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("ms %.4f GB/s %.1f" % (d["ms_per_step"], d["input_GBps_all_ranks"]), {k:d[k] for k in ("bit_exact","mismatches") if k in d})
except Exception as e: print("FAILED", e)')" | tee -a gpurun_out/ab.log
  done
done
cp /tmp/libpfac_keep.so pfac_b200/lib/libpfac.so
