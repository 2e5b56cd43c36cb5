"""GPU parity tests: the CUDA path, called through the C ABI (libpfac.so), against the oracle.

Bit-exact (integer results).  The oracle is oracle/libpfac_oracle.so (plain-C restatement of the
reference CPU matcher); /root/reference is never read here.
"""
import os

import numpy as np
import pytest

from pfac_b200 import synth

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def _oracle(path):
    from oracle import Oracle
    return Oracle(path)


def _dev_match(pf, text, cuda, owned=None):
    d_in = torch.from_numpy(text).to(cuda)
    n = text.size if owned is None else owned
    d_out = torch.full((max(n, 1),), -7, dtype=torch.int32, device=cuda)
    if owned is None:
        pf.matchFromDevice(d_in, n, d_out)
    else:
        pf.matchShardFromDevice(d_in, owned, text.size, d_out)
    torch.cuda.synchronize()
    return d_out[:n].cpu().numpy()


def _dev_reduce(pf, text, cuda, pos64=False, alias=None):
    d_in = torch.from_numpy(text).to(cuda)
    n = text.size
    d_id = torch.full((max(n, 1),), -7, dtype=torch.int32, device=cuda)
    if pos64:
        d_pos = torch.full((max(n, 1),), -7, dtype=torch.int64, device=cuda)
        m = pf.matchFromDeviceReduce64(d_in, n, d_id, d_pos)
    else:
        d_pos = torch.full((max(n, 1),), -7, dtype=torch.int32, device=cuda)
        m = pf.matchFromDeviceReduce(d_in, n, d_id, d_pos, alias=alias)
    return m, d_id.cpu().numpy(), d_pos.cpu().numpy()


def test_readme_example_all_entry_points(cuda, golden_dir, tmp_path):
    """BASELINE config 1 / goldens G1, G2, G3 (reference README.md:114-120; user guide p.21,27,29)."""
    from pfac_b200 import PFAC, PerfMode
    pat = os.path.join(golden_dir, "example_pattern")
    text = np.fromfile(os.path.join(golden_dir, "example_input"), dtype=np.uint8)
    g1 = np.array([1, 3, 4, 0, 4, 0, 2, 0, 0, 0], dtype=np.int32)
    with PFAC() as pf:
        pf.readPatternFromFile(pat)
        assert np.array_equal(pf.matchFromHost(text), g1)
        assert np.array_equal(_dev_match(pf, text, cuda), g1)
        for mode in (PerfMode.TIME_DRIVEN, PerfMode.SPACE_DRIVEN):
            pf.setPerfMode(mode)
            assert np.array_equal(_dev_match(pf, text, cuda), g1)
            for alias in (None, "reduceOnDevice", "reduceInplaceOnDevice"):
                m, ids, pos = _dev_reduce(pf, text, cuda, alias=alias)
                assert m == 5
                assert ids[:5].tolist() == [1, 3, 4, 4, 2]
                assert pos[:5].tolist() == [0, 1, 2, 4, 6]
                assert (ids[5:] == -7).all() and (pos[5:] == -7).all()  # nothing written past M
            ids, pos = pf.matchFromHostReduce(text)
            assert ids.tolist() == [1, 3, 4, 4, 2] and pos.tolist() == [0, 1, 2, 4, 6]
        dump = tmp_path / "table.txt"
        pf.dumpTransitionTable(str(dump))
        assert dump.read_bytes() == open(os.path.join(golden_dir, "example_pattern.dump"), "rb").read()


def test_second_machine(cuda, golden_dir):
    """Golden G4 (PFAC_hash_draft.pdf p.1 machine; vector from the reference CPU path)."""
    from pfac_b200 import PFAC
    text = np.fromfile(os.path.join(golden_dir, "example_input2"), dtype=np.uint8)
    with PFAC() as pf:
        pf.readPatternFromFile(os.path.join(golden_dir, "example_pattern2"))
        got = _dev_match(pf, text, cuda)
    assert got.tolist() == [4, 3, 0, 4, 5, 0, 0, 1, 7, 9, 1, 8, 9, 1, 0]


CASES = [
    # name, pattern generator, text kind, n, plant every
    ("c2_small", lambda: synth.patterns_c2(200, seed=11), "random", 300_001, 512),
    ("c2_1k", lambda: synth.patterns_c2(1000), "random", (1 << 22) + 13, 4096),
    ("snort_small", lambda: synth.patterns_snort_like(1500, seed=12), "ascii", (1 << 21) + 5, 1024),
    ("dna_small", lambda: synth.patterns_dna(400, seed=13, short=8), "dna", 1_000_003, 0),
]


@pytest.mark.parametrize("name,gen,kind,n,every", CASES, ids=[c[0] for c in CASES])
def test_dense_and_reduce_vs_oracle(cuda, tmp_path, name, gen, kind, n, every):
    from pfac_b200 import PFAC, PerfMode
    pats = gen()
    pfile = synth.write_pattern_file(str(tmp_path / "pat.txt"), pats)
    text = synth.make_text(kind, 1000 + len(pats), 0, n, n, pats, every)
    orc = _oracle(pfile)
    want = orc.match(text)
    want_ids, want_pos = orc.reduce(want)
    assert want_ids.size > 0, "workload must contain matches"
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        for mode in (PerfMode.TIME_DRIVEN, PerfMode.SPACE_DRIVEN):
            pf.setPerfMode(mode)
            got = _dev_match(pf, text, cuda)
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, "%s mode %s: first mismatch at %d got %d want %d (%d total)" % (
                name, mode, bad[0], got[bad[0]], want[bad[0]], bad.size)
            m, ids, pos = _dev_reduce(pf, text, cuda)
            assert m == want_ids.size
            assert np.array_equal(ids[:m], want_ids)
            assert np.array_equal(pos[:m].astype(np.int64), want_pos)
            m64, ids64, pos64 = _dev_reduce(pf, text, cuda, pos64=True)
            assert m64 == m and np.array_equal(ids64[:m], want_ids) and np.array_equal(pos64[:m], want_pos)
        assert np.array_equal(pf.matchFromHost(text), want)
        hid, hpos = pf.matchFromHostReduce(text)
        assert np.array_equal(hid, want_ids) and np.array_equal(hpos.astype(np.int64), want_pos)
