import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from pfac_b200 import PFAC
from workloads import synth
import bench
pats = synth.patterns_c2(1000)
pfile = synth.write_pattern_file('/tmp/c2.pat', pats)
dev = torch.device('cuda:0')
pf = PFAC(); pf.readPatternFromFile(pfile)
for mb in [int(x) for x in sys.argv[1:]]:
    n = mb << 20
    shard, owned = bench.make_shard(0, 1, n, pats)
    d_in = torch.from_numpy(shard).to(dev)
    cap = max(n // 16, 1 << 20)
    d_id = torch.empty(cap, dtype=torch.int32, device=dev); d_pos = torch.empty(cap, dtype=torch.int64, device=dev)
    d_out = torch.empty(n, dtype=torch.int32, device=dev)
    pf.matchFromDevice(d_in, n, d_out); torch.cuda.synchronize()
    nz = int((d_out != 0).sum())
    print(mb, 'MiB dense nonzeros', nz, flush=True)
    for rep in range(3):
        t0 = time.perf_counter()
        m = pf.matchFromDeviceReduce64(d_in, n, d_id, d_pos)
        dt = time.perf_counter() - t0
        print('  reduce', m, 'ms', round(dt*1e3, 3), 'GB/s', round(n/dt/1e9, 1), flush=True)
    assert m == nz
    pos = torch.nonzero(d_out).flatten()
    assert torch.equal(pos, d_pos[:m]) and torch.equal(d_out[pos], d_id[:m])
print('ok')
