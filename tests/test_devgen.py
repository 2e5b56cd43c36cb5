"""The device workload generator (workloads/devgen) must be byte-equal to its numpy definition
(workloads/synth): the at-size parity tests and bench.py's C5 leg regenerate their text on the GPU."""
import numpy as np
import pytest

from workloads import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,every", [("random", 4096), ("ascii", 2048), ("dna", 0), ("dna", 4096)])
def test_device_text_equals_numpy(kind, every):
    import torch
    from workloads import devgen
    pats = {"random": synth.patterns_c2(300), "ascii": synth.patterns_snort_like(500),
            "dna": synth.patterns_dna(200)}[kind]
    seed = synth.SEED_BASE + 7
    total = (5 << 20) + 12345
    # windows: stream start, an odd offset across a 1 MiB boundary, the end of the stream
    for start, n in ((0, (2 << 20) + 3), ((1 << 20) - 70001, 150003), (total - (1 << 20) - 5, (1 << 20) + 5),
                     (4097, 1), (total - 1, 1)):
        want = synth.make_text(kind, seed, start, n, total, pats, every)
        got = devgen.make_text(kind, seed, start, n, total, pats, every, device="cuda:0")
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), want), (kind, start, n)


def test_device_text_far_offset():
    """Positions beyond 2**32 (C3 / C5 shards) use the same 64-bit counter."""
    import torch
    from workloads import devgen
    pats = synth.patterns_snort_like(300)
    seed = synth.SEED_BASE + 5
    total = 32 << 30
    start = (5 << 30) - 4096 - 17
    n = (1 << 20) + 8192 + 17
    want = synth.make_text("ascii", seed, start, n, total, pats, 2048)
    got = devgen.make_text("ascii", seed, start, n, total, pats, 2048, device="cuda:0")
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), want)
