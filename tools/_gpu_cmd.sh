cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/run_configs.py --config c5 --check-bytes 536870912 --steps 5 > gpurun_out/c5_n8.log 2> gpurun_out/c5_n8.err
grep -h '"rank": 0' gpurun_out/c5_n8.log | cut -c1-100
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 5 --skip-cpu --e2e-steps 1 2> gpurun_out/bench8.err | tail -1 > gpurun_out/bench8.json
cut -c1-300 gpurun_out/bench8.json
