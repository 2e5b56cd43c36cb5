cd $GRAFT_REPO_ROOT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --skip-cpu --skip-e2e > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:pfac_ -s 3 -c 4 --csv --log-file gpurun_out/traffic_r1.csv python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --reduce-steps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfac_dense -c 1 -f -o gpurun_out/prof_dense_r1 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --skip-reduce --bytes 268435456 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfac_reduce -c 1 -f -o gpurun_out/prof_reduce_r1 python tools/reduce_stress.py 256 > /dev/null 2>&1
ls -la gpurun_out/*.csv gpurun_out/prof_*_r1.ncu-rep
timeout 400 python tests/run_configs.py --config c3 --steps 5 2>&1 | tail -1 | cut -c1-100
PFAC_B200_COPY_STREAM=1 timeout 120 python tools/host_path_bench.py --skip-pinned 2>&1 | tail -1
timeout 120 python tools/host_path_bench.py --skip-pinned 2>&1 | tail -1
