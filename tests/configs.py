"""The BASELINE.json configs as workloads, shared by tests/test_gpu_configs.py,
tests/run_configs.py and bench.py: pattern set, text law, seed, size, planting period, API.

Text is regenerated on the GPU (workloads/devgen, byte-equal to the numpy definition in
workloads/synth) and checked against the oracle in chunks the reference's `int input_size` can hold.
"""
import os

import numpy as np

from workloads import synth

GIB = 1 << 30

CONFIGS = {
    # 1,000 patterns len 4-32 (255-symbol alphabet), planted random text, dense int32 result
    "c2": dict(patterns=lambda: synth.patterns_c2(1000), kind="random", seed=synth.SEED_BASE + 2,
               bytes=GIB, every=4096, api="dense"),
    # 20,000 Snort-like patterns, ASCII-weighted text: tables exceed shared memory, 64-bit indexing
    "c3": dict(patterns=lambda: synth.patterns_snort_like(20000), kind="ascii", seed=synth.SEED_BASE + 3,
               bytes=4 * GIB, every=2048, api="dense"),
    # DNA, 5,000 patterns len 8-24, natural match density 0.6 %, reduce
    "c4": dict(patterns=lambda: synth.patterns_dna(5000), kind="dna", seed=synth.SEED_BASE + 4,
               bytes=2_000_000_000, every=0, api="reduce"),
    # + 64 patterns of length 4-6: every tenth position matches
    "c4dense": dict(patterns=lambda: synth.patterns_dna(5000, short=64), kind="dna", seed=synth.SEED_BASE + 4,
                    bytes=2_000_000_000, every=0, api="reduce"),
    # 10,000 Snort-like patterns over 32 GiB sharded across the ranks, reduce + global offset scan
    "c5": dict(patterns=lambda: synth.patterns_snort_like(10000, seed=synth.SEED_BASE + 5), kind="ascii",
               seed=synth.SEED_BASE + 5, bytes=32 * GIB, every=2048, api="reduce64"),
}


def _patterns_an():
    """a, aa, ..., a^32 on top of the C2 dictionary: over the text a^n every position walks 32 steps and
    reports a^32 (the reference's own worst-case micro-benchmark, doc/PFAC_hash_draft.pdf p.9 Table 4)."""
    return [b"a" * k for k in range(1, 33)] + synth.patterns_c2(1000)


# adversarial texts (the reference publishes regular AND worst-case throughput):
#   wprefix  the C2 dictionary over a text made of concatenated pattern prefixes (half to all-but-one
#            byte of a random pattern each): every piece passes the first stage and walks deep, few match
#   wan      the text a^n against a^1..a^32: every position matches after the longest possible walk
CONFIGS.update({
    "wprefix_dense": dict(patterns=lambda: synth.patterns_c2(1000), kind="pool:prefix", seed=synth.SEED_BASE + 6,
                          bytes=GIB, every=0, api="dense"),
    "wprefix_reduce": dict(patterns=lambda: synth.patterns_c2(1000), kind="pool:prefix", seed=synth.SEED_BASE + 6,
                           bytes=GIB, every=0, api="reduce"),
    "wan_dense": dict(patterns=_patterns_an, kind="pool:an", seed=synth.SEED_BASE + 7, bytes=GIB, every=0, api="dense"),
    "wan_reduce": dict(patterns=_patterns_an, kind="pool:an", seed=synth.SEED_BASE + 7, bytes=GIB, every=0, api="reduce"),
})


def adversarial_pool(kind, pats, seed, size=4 << 20):
    """`size` bytes of adversarial text (the stream is this pool repeated)."""
    if kind == "pool:an":
        return np.full(size, ord("a"), dtype=np.uint8)
    rng = np.random.Generator(np.random.PCG64(seed))
    pieces, have = [], 0
    while have < size:
        p = pats[int(rng.integers(0, len(pats)))]
        cut = int(rng.integers(max(len(p) // 2, 1), max(len(p), 2)))
        pieces.append(p[:cut])
        have += cut
    return np.frombuffer(b"".join(pieces)[:size], dtype=np.uint8).copy()


def device_text(cfg, start, n, total_len, pats, device, out=None):
    """Stream bytes [start, start + n) of the config's text as a uint8 tensor on `device`."""
    if cfg["kind"].startswith("pool:"):
        import torch
        pool = adversarial_pool(cfg["kind"], pats, cfg["seed"])
        d_pool = torch.from_numpy(pool).to(device)
        reps = (start % pool.size + n + pool.size - 1) // pool.size
        return d_pool.repeat(reps)[start % pool.size:start % pool.size + n].contiguous()
    from workloads import devgen
    return devgen.make_text(cfg["kind"], cfg["seed"], start, n, total_len, pats, cfg["every"], device=device, out=out)


def host_text(cfg, start, n, total_len, pats):
    """The numpy definition (slow: minutes per GiB); for small windows and CPU-only callers."""
    return synth.make_text(cfg["kind"], cfg["seed"], start, n, total_len, pats, cfg["every"])


class ChunkedCheck:
    """Oracle results of a device-resident shard, one chunk at a time.

    for c0, c1, want in ChunkedCheck(orc, d_text, owned, halo): want = dense oracle result of
    positions [c0, c1) of the shard (int32 numpy), computed from text[c0 : c1 + halo]."""

    def __init__(self, orc, d_text, owned, halo, chunk=256 << 20, limit=None):
        self.orc, self.d_text, self.owned, self.halo, self.chunk = orc, d_text, owned, halo, chunk
        self.limit = owned if limit is None else min(limit, owned)

    def __iter__(self):
        total = int(self.d_text.numel())
        for c0 in range(0, self.limit, self.chunk):
            c1 = min(c0 + self.chunk, self.limit)
            seg = self.d_text[c0:min(c1 + self.halo, total)].cpu().numpy()
            yield c0, c1, self.orc.match_shard(seg, c1 - c0)


def append_result(res, name="configs.jsonl"):
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", name), "a") as f:
        f.write(json.dumps(res) + "\n")


def nonzero_pairs(dense, base=0):
    pos = np.flatnonzero(dense)
    return dense[pos].astype(np.int32), pos.astype(np.int64) + base
