set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/host_path_bench.py 2>&1 | tail -1
PFAC_B200_STAGE_CHUNK_MB=32 timeout 120 python tools/host_path_bench.py --skip-pinned 2>&1 | tail -1
PFAC_B200_STAGE_CHUNK_MB=2 timeout 120 python tools/host_path_bench.py --skip-pinned 2>&1 | tail -1
PFAC_B200_COPY_THREADS=16 timeout 120 python tools/host_path_bench.py --skip-pinned 2>&1 | tail -1
PFAC_B200_COPY_THREADS=4 timeout 120 python tools/host_path_bench.py --skip-pinned 2>&1 | tail -1
PFAC_B200_COPY_THREADS=1 timeout 120 python tools/host_path_bench.py --skip-pinned 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfac_dense -c 1 -s 3 -o gpurun_out/c3_dense -f python tests/run_configs.py --config c3 --bytes 536870912 --check-bytes 0 --steps 1 > gpurun_out/c3_ncu.log 2>&1; tail -2 gpurun_out/c3_ncu.log
