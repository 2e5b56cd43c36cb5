"""Multi-GPU host logic on CPU: shard bounds + halo sufficiency, 64-bit global positions and the
count all-gather / exclusive scan, with the oracle standing in for the per-shard kernel.
Includes a world_size-2 gloo run of the real torch.distributed code path."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import Oracle
from workloads import synth
from pfac_b200.sharding import exclusive_offsets, shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_align():
    for total in (0, 1, 4095, 4096, 4097, 1_000_003, 1 << 26):
        for world in (1, 2, 3, 8):
            covered = 0
            for r in range(world):
                s, owned, tot = shard_bounds(total, world, r, 33)
                assert s == covered or owned == 0
                assert s % 4096 == 0 or owned == 0
                assert owned <= tot <= owned + 32 and s + tot <= total
                if s + owned < total:
                    assert tot == min(owned + 32, total - s)  # full halo unless the stream ends
                covered += owned
            assert covered == total


def test_exclusive_offsets():
    assert exclusive_offsets([3, 0, 5, 1]) == ([0, 3, 3, 8], 9)
    assert exclusive_offsets([]) == ([], 0)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_reduce_equals_whole_stream(tmp_path, world):
    """Emulates k ranks in one process: per-shard dense (oracle shard form) -> local reduce with
    pos_base -> exclusive scan of counts -> global list == reduce of the whole stream."""
    pats = synth.patterns_snort_like(600, seed=5)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    o = Oracle(pfile)
    n = 300_017
    text = synth.make_text("ascii", 55, 0, n, n, pats, 512)
    want_ids, want_pos = o.reduce(o.match(text))
    parts, counts = [], []
    for r in range(world):
        s, owned, tot = shard_bounds(n, world, r, o.max_pattern_len, align=4096)
        # each rank regenerates only its own bytes: the counter-based generator must agree
        mine = synth.make_text("ascii", 55, s, tot, n, pats, 512)
        assert np.array_equal(mine, text[s:s + tot])
        dense = o.match_shard(mine, owned)
        ids, pos = o.reduce(dense)
        parts.append((ids, pos + s))
        counts.append(ids.size)
    offs, total = exclusive_offsets(counts)
    assert total == want_ids.size
    g_ids = np.empty(total, dtype=np.int32)
    g_pos = np.empty(total, dtype=np.int64)
    for (ids, pos), off in zip(parts, offs):
        g_ids[off:off + ids.size] = ids
        g_pos[off:off + pos.size] = pos
    assert np.array_equal(g_ids, want_ids) and np.array_equal(g_pos, want_pos)


def test_halo_of_maxlen_minus_one_is_necessary_and_sufficient(tmp_path):
    pats = [b"ABCDEFGH", b"CD", b"H"]
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    o = Oracle(pfile)
    text = np.frombuffer(b"xxABCDEFGHxx", dtype=np.uint8)
    want = o.match(text)
    cut = 3  # the long pattern starts at the last owned position
    full = o.match_shard(text[:cut + o.max_pattern_len - 1], cut)
    assert np.array_equal(full, want[:cut])
    short = o.match_shard(text[:cut + o.max_pattern_len - 2], cut)
    assert not np.array_equal(short, want[:cut])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from oracle import Oracle
from workloads import synth
from pfac_b200.sharding import shard_bounds, allgather_count_offsets, place_runs
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
pats = synth.patterns_c2(300, seed=9, min_len=2, max_len=20)
pfile = os.path.join(%(tmp)r, "p%%d.txt" %% rank)
synth.write_pattern_file(pfile, pats)
o = Oracle(pfile)
n = 200_003
s, owned, tot = shard_bounds(n, world, rank, o.max_pattern_len)
mine = synth.make_text("random", 77, s, tot, n, pats, 256)
ids, pos = o.reduce(o.match_shard(mine, owned))
pos = pos + s
off, total, counts = allgather_count_offsets(ids.size)
# place every rank's run at its scanned offset in one global list (gather for the check)
g_ids = torch.zeros(total, dtype=torch.int32); g_pos = torch.zeros(total, dtype=torch.int64)
g_ids[off:off + ids.size] = torch.from_numpy(ids); g_pos[off:off + pos.size] = torch.from_numpy(pos)
dist.all_reduce(g_ids); dist.all_reduce(g_pos)
# the same list delivered point to point to the last rank (place_runs)
cap = ids.size + 5
t_ids = torch.full((cap,), -1, dtype=torch.int32); t_ids[:ids.size] = torch.from_numpy(ids)
t_pos = torch.full((cap,), -1, dtype=torch.int64); t_pos[:pos.size] = torch.from_numpy(pos)
p_ids, p_pos = place_runs(t_ids, t_pos, counts, dst=world - 1)
if rank == world - 1:
    assert torch.equal(p_ids, g_ids) and torch.equal(p_pos, g_pos)
else:
    assert p_ids is None and p_pos is None
if rank == 0:
    text = synth.make_text("random", 77, 0, n, n, pats, 256)
    w_ids, w_pos = o.reduce(o.match(text))
    assert total == w_ids.size and sum(counts) == total
    assert np.array_equal(g_ids.numpy(), w_ids) and np.array_equal(g_pos.numpy(), w_pos)
    print("OK", total, counts)
dist.destroy_process_group()
"""


def test_gloo_world2_count_scan(tmp_path):
    import subprocess
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "tmp": str(tmp_path)})
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "OK" in outs[0]
