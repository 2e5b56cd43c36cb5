timeout 300 python bench.py 2>gpurun_out/bench.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('bench', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'reduce', d['reduce']['call_only_value'], d['reduce']['ms_per_call'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfac_reduce -c 1 -s 2 -o gpurun_out/c2_reduce_h -f python tools/reduce_stress.py 256 > gpurun_out/c2r_ncu.log 2>&1; tail -2 gpurun_out/c2r_ncu.log
