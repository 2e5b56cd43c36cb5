"""Host-side shard arithmetic for the multi-GPU path.

Every start position is independent (reference PFAC_CPU.cpp:76), so an N-byte stream is cut
into contiguous shards; rank g owns start positions [s_g, e_g) and additionally holds the
next H = maxPatternLen-1 bytes (tail halo; real bytes, or fewer when the stream ends).  The
only exchange is the per-rank match count: an exclusive scan of the counts gives each rank's offset
into the global (ID, position) list.  The library does that scan inside the reduce kernel over peer
memory (PFAC_comm, include/PFAC_ext.h); the torch.distributed forms below (an all-gather of one int64
per rank + a local scan; send/recv placement of the runs) are the host-side cross-check used by the
tests (gloo on CPU) and by tests/run_configs.py.
This is what the reference's test/omp_PFAC.cpp:316-383 does by hand with host threads
(there with maxPatternLen+1 bytes of overlap and no reduced output).
"""
from typing import List, Tuple


def shard_bounds(total_len: int, world: int, rank: int, max_pattern_len: int,
                 align: int = 4096) -> Tuple[int, int, int]:
    """(start, n_owned, n_total) for `rank`: shard starts are multiples of `align` (keeps the
    device pointers 16-byte aligned for the TMA path), n_total = owned + tail halo."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-total_len // world)               # ceil
    per = -(-per // align) * align if per else 0
    start = min(rank * per, total_len)
    end = min(start + per, total_len)
    halo = max(max_pattern_len - 1, 0)
    n_total = min(end + halo, total_len) - start
    return start, end - start, n_total


def exclusive_offsets(counts: List[int]) -> Tuple[List[int], int]:
    """Exclusive scan of per-rank match counts -> (offsets, total)."""
    offs, run = [], 0
    for c in counts:
        offs.append(run)
        run += int(c)
    return offs, run


def allgather_count_offsets(count: int, device=None, group=None) -> Tuple[int, int, List[int]]:
    """The cross-GPU step: all-gather of one int64 per rank (NCCL over NVLink on GPUs, gloo in
    the CPU tests) + local exclusive scan.  Returns (my_offset, total, all_counts)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([int(count)], dtype=torch.int64, device=device)
    allc = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine, group=group)
    counts = [int(x) for x in allc.cpu().tolist()]
    offs, total = exclusive_offsets(counts)
    return offs[rank], total, counts


def place_runs(ids, pos, counts: List[int], dst: int = 0, group=None):
    """Optional second step of SURVEY.md 8(e): one global (ID, position) list on rank `dst`.  Every
    rank sends its run (the first counts[rank] entries of `ids` int32 / `pos` int64, positions
    already global) point to point; `dst` receives each run directly at its scanned offset, so no
    rank holds more than its own run plus, on `dst`, the final list.  NCCL send/recv over NVLink on
    GPUs (tensors on the device), gloo in the CPU tests.  Returns (ids, pos) of length sum(counts)
    on `dst`, (None, None) elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    offs, total = exclusive_offsets(counts)
    n = int(counts[rank])
    if rank != dst:
        if n:
            dist.send(ids[:n].contiguous(), dst, group=group)
            dist.send(pos[:n].contiguous(), dst, group=group)
        return None, None
    g_ids = torch.empty(total, dtype=torch.int32, device=ids.device)
    g_pos = torch.empty(total, dtype=torch.int64, device=pos.device)
    for r in range(world):
        c, o = int(counts[r]), offs[r]
        if c == 0:
            continue
        if r == rank:
            g_ids[o:o + c] = ids[:c]
            g_pos[o:o + c] = pos[:c]
        else:
            dist.recv(g_ids[o:o + c], r, group=group)   # contiguous slices: received in place
            dist.recv(g_pos[o:o + c], r, group=group)
    return g_ids, g_pos
