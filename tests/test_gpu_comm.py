"""The in-library cross-GPU step (PFAC_comm): the reduce kernel publishes its match count into every
rank's mailbox over peer memory and scans the counts itself; PFAC_commGatherRuns places the runs in
one list.  World 1 runs on any box; the 2-GPU cases (one process with peer access; buffers on a GPU
other than the handle's, as reference test/UVA.cpp:137) are skipped on a single-GPU box; the
one-process-per-GPU form (CUDA IPC + torchrun) is tests/test_gpu_configs.py::test_c5_two_ranks."""
import os

import numpy as np
import pytest

from tests.helpers import CheckerOracle
from workloads import synth

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def _workload(tmp_path, n):
    pats = synth.patterns_snort_like(2000, seed=synth.SEED_BASE + 9)
    pfile = synth.write_pattern_file(str(tmp_path / "p.pat"), pats)
    text = synth.make_text("ascii", synth.SEED_BASE + 9, 0, n, n, pats, every=1024)
    orc = CheckerOracle(pfile)
    want_ids, want_pos = orc.reduce(orc.match(text))
    return pfile, text, orc, want_ids, want_pos


def test_world_one_scan_and_gather(cuda, tmp_path):
    from pfac_b200 import PFAC, PFACComm
    n = (3 << 20) + 77
    pfile, text, orc, want_ids, want_pos = _workload(tmp_path, n)
    pf = PFAC()
    pf.readPatternFromFile(pfile)
    comm = PFACComm(0, 1, list_capacity=want_ids.size + 10)
    d_in = torch.from_numpy(text).to(cuda)
    d_id = torch.full((n,), -7, dtype=torch.int32, device=cuda)
    d_pos = torch.full((n,), -7, dtype=torch.int64, device=cuda)
    for rep in range(3):   # epochs advance, parities alternate
        off, total, count = pf.matchShardFromDeviceReduce64Global(comm, d_in, n, n, 1000, d_id, d_pos)
        assert (off, total, count) == (0, want_ids.size, want_ids.size)
        # asynchronous form: the scan stays on the device
        d_scan = torch.zeros(3, dtype=torch.int64, device=cuda)
        pf.matchShardFromDeviceReduce64Global(comm, d_in, n, n, 1000, d_id, d_pos, d_scan=d_scan, sync=False)
        pf.gatherRuns(comm, 0, d_id, d_pos, d_scan=d_scan, sync=True)
        assert d_scan.cpu().tolist() == [0, want_ids.size, want_ids.size]
        ids, pos = comm.read_global_list(want_ids.size)
        assert np.array_equal(ids, want_ids) and np.array_equal(pos, want_pos + 1000)
    # an empty shard still takes part in the exchange
    assert pf.matchShardFromDeviceReduce64Global(comm, d_in, 0, 0, 0, d_id, d_pos) == (0, 0, 0)
    comm.destroy()
    pf.destroy()


def test_two_gpus_one_process_peer_mailboxes(cuda, tmp_path):
    """Two handles, two devices, one host thread: both kernels are enqueued without a host sync and
    wait for each other's count on the device; then both runs land in rank 0's list."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pfac_b200 import PFAC, PFACComm
    from pfac_b200.sharding import shard_bounds
    n = (6 << 20) + 123
    pfile, text, orc, want_ids, want_pos = _workload(tmp_path, n)
    comms = PFACComm.local([0, 1], list_capacity=want_ids.size + 10)
    pfs, bufs = [], []
    for r in range(2):
        torch.cuda.set_device(r)
        pf = PFAC()
        pf.readPatternFromFile(pfile)
        pfs.append(pf)
        s, owned, total = shard_bounds(n, 2, r, orc.max_pattern_len)
        dev = torch.device("cuda", r)
        bufs.append((s, owned, total, torch.from_numpy(text[s:s + total]).to(dev),
                     torch.empty(max(owned, 1), dtype=torch.int32, device=dev),
                     torch.empty(max(owned, 1), dtype=torch.int64, device=dev),
                     torch.zeros(3, dtype=torch.int64, device=dev)))
    for rep in range(2):
        for r in (1, 0):   # launch order must not matter
            torch.cuda.set_device(r)
            s, owned, total, d_in, d_id, d_pos, d_scan = bufs[r]
            pfs[r].matchShardFromDeviceReduce64Global(comms[r], d_in, owned, total, s, d_id, d_pos, d_scan=d_scan, sync=False)
        for r in (0, 1):
            torch.cuda.set_device(r)
            s, owned, total, d_in, d_id, d_pos, d_scan = bufs[r]
            pfs[r].gatherRuns(comms[r], 0, d_id, d_pos, d_scan=d_scan, sync=False)
        for r in (0, 1):
            torch.cuda.synchronize(r)
        c0 = int(((want_pos < bufs[1][0])).sum())
        assert bufs[0][6].cpu().tolist() == [0, want_ids.size, c0]
        assert bufs[1][6].cpu().tolist() == [c0, want_ids.size, want_ids.size - c0]
        ids, pos = comms[0].read_global_list(want_ids.size)
        assert np.array_equal(ids, want_ids) and np.array_equal(pos, want_pos)
    torch.cuda.set_device(0)
    for c in comms:
        c.destroy()
    for pf in pfs:
        pf.destroy()


def test_buffers_on_a_peer_gpu(cuda, tmp_path):
    """Reference test/UVA.cpp:137: the handle lives on GPU 0, input and output buffers on GPU 1 (peer
    access enabled by the caller); dense and reduce results equal the oracle's."""
    if torch.cuda.device_count() < 2 or not torch.cuda.can_device_access_peer(0, 1):
        pytest.skip("needs 2 GPUs with peer access")
    from pfac_b200 import PFAC
    n = (2 << 20) + 5
    pfile, text, orc, want_ids, want_pos = _workload(tmp_path, n)
    torch.cuda.set_device(0)
    pf = PFAC()
    pf.readPatternFromFile(pfile)
    dev1 = torch.device("cuda:1")
    d_in = torch.from_numpy(text).to(dev1)
    _ = d_in[:16].to(cuda)   # a cross-device copy makes torch enable peer access 0 <-> 1
    d_out = torch.full((n,), -7, dtype=torch.int32, device=dev1)
    pf.matchFromDevice(d_in, n, d_out)
    torch.cuda.synchronize(0)
    assert np.array_equal(d_out.cpu().numpy(), orc.match(text))
    d_pos = torch.full((n,), -7, dtype=torch.int32, device=dev1)
    m = pf.matchFromDeviceReduce(d_in, n, d_out, d_pos)
    assert m == want_ids.size
    assert np.array_equal(d_out[:m].cpu().numpy(), want_ids)
    assert np.array_equal(d_pos[:m].cpu().numpy().astype(np.int64), want_pos)
    pf.destroy()


def test_reduce_with_stated_capacity(cuda, tmp_path):
    """PFAC_matchShardFromDeviceReduce64Cap: nothing is stored past the stated capacity, the full count
    still comes back, and a capacity that was too small is an error status."""
    from pfac_b200 import PFAC, PFACError, Status
    n = (1 << 20) + 9
    pfile, text, orc, want_ids, want_pos = _workload(tmp_path, n)
    pf = PFAC()
    pf.readPatternFromFile(pfile)
    d_in = torch.from_numpy(text).to(cuda)
    m = want_ids.size
    d_id = torch.full((m + 7,), -7, dtype=torch.int32, device=cuda)
    d_pos = torch.full((m + 7,), -7, dtype=torch.int64, device=cuda)
    assert pf.matchShardFromDeviceReduce64Cap(d_in, n, n, 5, d_id, d_pos) == m
    assert np.array_equal(d_id[:m].cpu().numpy(), want_ids) and np.array_equal(d_pos[:m].cpu().numpy(), want_pos + 5)
    assert d_id[m:].eq(-7).all() and d_pos[m:].eq(-7).all()
    small = m // 3
    d_id.fill_(-7)
    d_pos.fill_(-7)
    with pytest.raises(PFACError) as e:
        pf.matchShardFromDeviceReduce64Cap(d_in, n, n, 5, d_id[:small], d_pos[:small])
    assert e.value.status == Status.INVALID_PARAMETER
    assert np.array_equal(d_id[:small].cpu().numpy(), want_ids[:small])      # the first `capacity` entries are there
    assert d_id[small:].eq(-7).all() and d_pos[small:].eq(-7).all()          # and nothing beyond them
    pf.destroy()


def test_multi_gpu_example_program(cuda, tmp_path):
    """examples/multi_gpu_scan.cpp: a C++ caller of the public headers only.  One process, every visible
    GPU, in-kernel count scan, one global list on GPU 0; the program compares that list element by
    element with a single-GPU run (as reference test/omp_PFAC.cpp:397-439 does) and exits non-zero on a
    difference.  Here its printed total must also be the oracle's."""
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from pfac_b200 import library_path
    libdir = os.path.dirname(library_path())
    exe = str(tmp_path / "multi_gpu_scan")
    subprocess.run(["/usr/bin/g++", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(ROOT, "examples", "multi_gpu_scan.cpp"), "-L", libdir, "-lpfac",
                    "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    n = (5 << 20) + 321
    pfile, text, orc, want_ids, want_pos = _workload(tmp_path, n)
    text.tofile(str(tmp_path / "text.bin"))
    r = subprocess.run([exe, pfile, str(tmp_path / "text.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    gpus = torch.cuda.device_count()
    assert "number of matched = %d on %d GPU(s); single-GPU check: identical" % (want_ids.size, gpus) in r.stdout, r.stdout
    first = ["At position %4d, match pattern %d" % (want_pos[i], want_ids[i]) for i in range(min(10, want_ids.size))]
    assert [l for l in r.stdout.splitlines() if l.startswith("At position")] == first
