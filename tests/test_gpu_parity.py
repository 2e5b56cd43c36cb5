"""GPU parity tests: the CUDA path, called through the C ABI (libpfac.so), against the oracle.

Bit-exact (integer results).  The checker is the reference's own CPU matcher compiled from the
reference sources (oracle/_ref/libpfac_ref.so, built where /root/reference exists and shipped as a
binary) when present, else oracle/libpfac_oracle.so (the plain-C restatement); /root/reference is
never read here.
"""
import os

import numpy as np
import pytest

from workloads import synth

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def _oracle(path):
    from tests.helpers import CheckerOracle
    return CheckerOracle(path)


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
    ("snort_10k", lambda: synth.patterns_snort_like(10000, seed=14), "ascii", (1 << 23) + 777, 2048),
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


def _check_all(pf, orc, text, cuda, n_owned=None):
    """dense (+ shard form) and both reduce widths against the oracle on one text."""
    n = text.size if n_owned is None else n_owned
    want = orc.match_shard(text, n) if n_owned is not None else orc.match(text)
    got = _dev_match(pf, text, cuda, owned=n_owned)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, "dense mismatch at %d: got %d want %d (n=%d)" % (bad[0], got[bad[0]], want[bad[0]], n)
    wid, wpos = orc.reduce(want)
    if n_owned is None:
        m, ids, pos = _dev_reduce(pf, text, cuda)
        assert m == wid.size and np.array_equal(ids[:m], wid) and np.array_equal(pos[:m].astype(np.int64), wpos)
    d_in = torch.from_numpy(text).to(cuda)
    d_id = torch.full((max(n, 1),), -7, dtype=torch.int32, device=cuda)
    d_pos = torch.full((max(n, 1),), -7, dtype=torch.int64, device=cuda)
    base = 5_000_000_000  # positions beyond 2^32 must survive
    m = pf.matchShardFromDeviceReduce64(d_in, n, text.size, base, d_id, d_pos)
    assert m == wid.size
    assert np.array_equal(d_id[:m].cpu().numpy(), wid)
    assert np.array_equal(d_pos[:m].cpu().numpy(), wpos + base)
    assert (d_id[m:] == -7).all() and (d_pos[m:] == -7).all()


def test_edge_sizes_tail_tiles_and_shards(cuda, tmp_path):
    """Empty / tiny / ragged inputs, tile-boundary sizes, patterns cut off by the end of the input,
    and the shard form (owned positions + tail halo)."""
    from pfac_b200 import PFAC
    pats = synth.patterns_c2(300, seed=21, min_len=1, max_len=40, prefix_pairs=60)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        for n in [1, 2, 3, 15, 16, 17, 39, 40, 511, 512, 513, 527, 528, 1023, 1024, 4095, 4096, 4097,
                  16383, 16384, 16385, 70001, 262144 + 17]:
            text = synth.make_text("random", 300 + n, 0, n, n, pats, 96)
            _check_all(pf, orc, text, cuda)
        text = synth.make_text("random", 77, 0, 100_000, 100_000, pats, 64)
        for owned in [1, 511, 512, 4096, 50_000, 99_961, 99_999, 100_000]:
            _check_all(pf, orc, text, cuda, n_owned=owned)
        # size 0: SUCCESS, nothing written (reference PFAC.cpp:860-862)
        d_in = torch.zeros(16, dtype=torch.uint8, device=cuda)
        d_out = torch.full((16,), -7, dtype=torch.int32, device=cuda)
        pf.matchFromDevice(d_in, 0, d_out)
        assert pf.matchFromDeviceReduce(d_in, 0, d_out, d_out) == 0
        torch.cuda.synchronize()
        assert (d_out == -7).all()


@pytest.mark.parametrize("policy", ["hash", "exact"])
def test_first_stage_filters_agree(cuda, tmp_path, monkeypatch, policy):
    """The hashed 4-gram first stage (large byte-alphabet dictionaries, pfac_table.h hfilt) and the exact
    2-gram stage must both give the oracle's result: 1-, 2- and 3-byte patterns, matches in the last
    bytes of the input, ragged sizes, the shard form, SPACE_DRIVEN (no room for the filter)."""
    from pfac_b200 import PFAC, PerfMode
    monkeypatch.setenv("PFAC_B200_FILTER", policy)
    pats = synth.patterns_snort_like(2500, seed=71)
    pats += [b"q", b"zq", b"~z", b"xyz", b"e", b"th", b"the", b"them", b"\xff", b"\x00\x01\x02"]
    pats = list(dict.fromkeys(pats))
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        assert bool(pf.tableInfo()["hashed_filter"]) == (policy == "hash")
        for n in [1, 2, 3, 4, 5, 17, 511, 512, 513, 515, 1024, 4099, 70001, 300_000]:
            text = synth.make_text("ascii", 700 + n, 0, n, n, pats, 80)
            _check_all(pf, orc, text, cuda)
            for tail in (b"q", b"zq", b"xyz", b"the", b"them", b"\xff"):
                if len(tail) <= n:
                    text[n - len(tail):] = np.frombuffer(tail, dtype=np.uint8)
                    _check_all(pf, orc, text, cuda)
        text = synth.make_text("ascii", 7171, 0, 200_000, 200_000, pats, 64)
        for owned in [1, 511, 512, 513, 100_000, 199_990, 199_999, 200_000]:
            _check_all(pf, orc, text, cuda, n_owned=owned)
        pf.setPerfMode(PerfMode.SPACE_DRIVEN)
        assert pf.tableInfo()["hashed_filter"] == 0
        _check_all(pf, orc, text, cuda)


def test_pair_filter_in_both_kernels(cuda, tmp_path, monkeypatch):
    """Sparse dictionaries without 1- and 2-byte patterns: the reduce kernel's layout uses the pair filter (one
    lookup per two start positions), the dense kernel's the per-position filter; PFAC_B200_FILTER=hash puts the
    pair filter into the dense kernel too.  3-byte patterns, matches at either parity and in the last bytes."""
    from pfac_b200 import PFAC
    pats = synth.patterns_c2(700, seed=91) + [b"xyz", b"abc", b"bcd", b"zzz", b"zzzz", b"\xff\xff\xff", b"\x00\x01\x02"]
    pats = list(dict.fromkeys(pats))
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    for policy, dense_stage in ((None, 1), ("hash", 3), ("nopair", 1)):
        if policy:
            monkeypatch.setenv("PFAC_B200_FILTER", policy)
        else:
            monkeypatch.delenv("PFAC_B200_FILTER", raising=False)
        with PFAC() as pf:
            pf.readPatternFromFile(pfile)
            assert pf.tableInfo()["hashed_filter"] == dense_stage
            assert pf.tableInfo(reduce=True)["hashed_filter"] == (1 if policy == "nopair" else 3)
            for n in [3, 4, 5, 6, 7, 511, 512, 513, 1535, 1536, 1537, 1539, 70_001, 400_003]:
                text = synth.make_text("random", 900 + n, 0, n, n, pats, 61)   # odd period: both parities
                _check_all(pf, orc, text, cuda)
                for tail in (b"xyz", b"zzzz", b"abc", b"\xff\xff\xff"):
                    for back in (0, 1):
                        if len(tail) + back <= n:
                            t = text.copy()
                            t[n - back - len(tail):n - back] = np.frombuffer(tail, dtype=np.uint8)
                            _check_all(pf, orc, t, cuda)
            text = synth.make_text("random", 9191, 0, 300_001, 300_001, pats, 37)
            for owned in [1, 2, 1535, 1536, 1537, 150_001, 299_999, 300_000, 300_001]:
                _check_all(pf, orc, text, cuda, n_owned=owned)


def test_dense_dictionary_with_many_matches_per_tile(cuda, tmp_path):
    """A large byte dictionary (row-indexed two-bit filter, zeros of the dense result by bulk stores) on a text
    with more matches per 1,536-position tile than a warp parks (64): the dense kernel's wait-and-patch path,
    the reduce kernel's spill ring."""
    from pfac_b200 import PFAC
    pats = synth.patterns_snort_like(12000, seed=15) + [b"e", b"t", b"a", b"o", b" ", b"n", b"th", b"he "]
    pats = list(dict.fromkeys(pats))
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    n = (1 << 21) + 1234
    text = synth.make_text("ascii", 515, 0, n, n, pats, 48)
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        info = pf.tableInfo()
        assert info["hashed_filter"] == 2 and info["hfilt_words"] == 16384, info
        _check_all(pf, orc, text, cuda)
        want = orc.match(text)
        assert (want > 0).mean() * 1536 > 80
        _check_all(pf, orc, text[:700_001].copy(), cuda, n_owned=699_000)


@pytest.mark.parametrize("policy", ["auto", "exact"])
def test_random_dictionaries_vs_oracle(cuda, tmp_path, monkeypatch, policy):
    """Random alphabets (all three symbol codings), pattern lengths 1..40 with shared prefixes, texts
    with foreign bytes, a match ending at the last byte: dense, reduce and shard forms against the
    oracle, for both first stages."""
    from pfac_b200 import PFAC
    if policy != "auto":
        monkeypatch.setenv("PFAC_B200_FILTER", policy)
    rng = np.random.default_rng(4242)
    with PFAC() as pf:
        for trial in range(14):
            asize = int(rng.choice([2, 4, 5, 16, 17, 64, 255]))
            alpha = rng.choice(np.setdiff1d(np.arange(256), [10]), size=asize, replace=False).astype(np.uint8)
            pats = []
            for _ in range(int(rng.integers(1, 400))):
                body = alpha[rng.integers(0, asize, size=int(rng.integers(1, 41)))].tobytes()
                if pats and rng.random() < 0.4:
                    body = (pats[int(rng.integers(0, len(pats)))] + body)[:60]
                pats.append(body)
            pats = list(dict.fromkeys(pats))
            pfile = synth.write_pattern_file(str(tmp_path / ("p%d.txt" % trial)), pats)
            orc = _oracle(pfile)
            pf.readPatternFromFile(pfile)
            n = int(rng.choice([1, 7, 513, 4097, 50_001, 131_072]))
            text = alpha[rng.integers(0, asize, size=n)].copy()
            k = max(1, n // 40)
            text[rng.integers(0, n, size=k)] = rng.integers(0, 256, size=k).astype(np.uint8)
            for p in pats[:50]:
                if len(p) <= n:
                    at = int(rng.integers(0, n - len(p) + 1))
                    text[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
            tail = pats[int(rng.integers(0, len(pats)))]
            if len(tail) <= n:
                text[n - len(tail):] = np.frombuffer(tail, dtype=np.uint8)
            _check_all(pf, orc, text, cuda)
            if n > 600:
                _check_all(pf, orc, text, cuda, n_owned=n - 37)


def test_unaligned_device_pointers(cuda, tmp_path):
    """Any pointer alignment is accepted (the reference needs 4-byte aligned input and reads up to
    3 bytes past the end, PFAC.cpp:838-841; this library does neither)."""
    from pfac_b200 import PFAC
    pats = synth.patterns_snort_like(400, seed=22)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    n = 150_001
    text = synth.make_text("ascii", 901, 0, n, n, pats, 200)
    want = orc.match(text)
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        for in_off, out_off in [(1, 0), (3, 1), (16, 2), (5, 3), (0, 1)]:
            buf = torch.zeros(n + 64, dtype=torch.uint8, device=cuda)
            buf[in_off:in_off + n] = torch.from_numpy(text).to(cuda)
            obuf = torch.full((n + 8,), -7, dtype=torch.int32, device=cuda)
            pf.matchFromDevice(buf[in_off:], n, obuf[out_off:])
            torch.cuda.synchronize()
            got = obuf.cpu().numpy()
            assert np.array_equal(got[out_off:out_off + n], want), (in_off, out_off)
            assert (got[:out_off] == -7).all() and (got[out_off + n:] == -7).all()


def test_high_match_density_compaction(cuda, tmp_path):
    """More matches per tile than a warp can park (reduce kernel's spill ring, dense kernel's wait-and-patch path), several
    rounds of CTA tiles, 1-byte and 2-byte patterns (every occurrence of a byte matches)."""
    from pfac_b200 import PFAC
    pats = [b"AC", b"GT", b"TT", b"CA", b"G", b"ACGTAC", b"TTTT", b"CAT", b"GATTACA", b"AA", b"TG", b"CC"]
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    n = 148 * 16384 * 2 + 12_345   # > 2 rounds of CTA tiles
    text = synth.dna_bytes(4242, 0, n)
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        _check_all(pf, orc, text, cuda)
        want = orc.match(text)
        assert (want > 0).mean() > 0.5


def test_patterns_longer_than_the_staged_halo(cuda, tmp_path):
    """Walks that run past the shared-memory halo (512 bytes) continue from global memory."""
    from pfac_b200 import PFAC
    rng = np.random.default_rng(5)
    longp = [bytes(rng.integers(32, 127, size=L, dtype=np.uint8).tolist()) for L in (600, 900, 1500, 2300)]
    pats = longp + [longp[0][:300], longp[1][:513], b"xyz", longp[2][:1499]]
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    n = 400_000
    text = synth.make_text("ascii", 99, 0, n, n, pats, 3000)
    # near misses: a long pattern with its last byte changed, and one cut off by the end
    t2 = np.frombuffer(longp[3], dtype=np.uint8).copy()
    t2[-1] ^= 1
    text[1000:1000 + t2.size] = t2
    text[n - 1400:n] = np.frombuffer(longp[2][:1400], dtype=np.uint8)
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        _check_all(pf, orc, text, cuda)
        assert set(np.unique(orc.match(text))) >= {1, 2, 3, 4, 5, 6}


def test_two_handles_and_pattern_reload(cuda, golden_dir, tmp_path):
    """Reference SimpleMultiGPU_pthread.cpp pattern: independent handles with different pattern sets;
    plus re-reading patterns into a live handle (PFAC.cpp:663-666)."""
    from pfac_b200 import PFAC
    t1 = np.fromfile(os.path.join(golden_dir, "example_input"), dtype=np.uint8)
    t2 = np.fromfile(os.path.join(golden_dir, "example_input2"), dtype=np.uint8)
    with PFAC() as a, PFAC() as b:
        a.readPatternFromFile(os.path.join(golden_dir, "example_pattern"))
        b.readPatternFromFile(os.path.join(golden_dir, "example_pattern2"))
        assert _dev_match(a, t1, cuda).tolist() == [1, 3, 4, 0, 4, 0, 2, 0, 0, 0]
        assert _dev_match(b, t2, cuda).tolist() == [4, 3, 0, 4, 5, 0, 0, 1, 7, 9, 1, 8, 9, 1, 0]
        a.readPatternFromFile(os.path.join(golden_dir, "example_pattern2"))
        assert _dev_match(a, t2, cuda).tolist() == [4, 3, 0, 4, 5, 0, 0, 1, 7, 9, 1, 8, 9, 1, 0]
        a.readPatternFromMemory(b"ED\nAB\n")
        assert _dev_match(a, t1, cuda).tolist() == [2, 0, 1, 0, 1, 0, 2, 0, 0, 0]


def test_errors_with_a_live_handle(cuda, tmp_path):
    from pfac_b200 import PFAC, PFACError, Status
    with PFAC() as pf:
        d = torch.zeros(32, dtype=torch.int32, device=cuda)
        with pytest.raises(PFACError) as e:
            pf.matchFromDevice(d, 8, d)
        assert e.value.status == Status.PATTERNS_NOT_READY
        with pytest.raises(PFACError) as e:
            pf.readPatternFromFile(str(tmp_path / "missing"))
        assert e.value.status == Status.FILE_OPEN_ERROR
        pf.readPatternFromMemory(b"AB\n")
        for call in (lambda: pf.matchFromDevice(None, 8, d), lambda: pf.matchFromDevice(d, 8, None)):
            with pytest.raises(PFACError) as e:
                call()
            assert e.value.status == Status.INVALID_PARAMETER
        with pytest.raises(PFACError) as e:
            pf.setPlatform(7)
        assert e.value.status == Status.INVALID_PARAMETER
        pf.setPlatform(1)      # CPU / CPU_OMP are accepted; work still runs on the GPU
        pf.setTextureMode(1)
        assert pf.matchFromHost(np.frombuffer(b"xxABxx", dtype=np.uint8)).tolist() == [0, 0, 1, 0, 0, 0]


def test_reduce_repeated_large_is_stable(cuda, tmp_path):
    """Many rounds of CTA tiles (ring-slot reuse in the reduce kernel), repeated calls: the list
    must equal the compaction of the dense result every time."""
    from pfac_b200 import PFAC
    pats = synth.patterns_c2(1000)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    n = (96 << 20) + 4099
    text = synth.make_text("random", 31337, 0, n, n, pats, 1024)
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        d_in = torch.from_numpy(text).to(cuda)
        d_out = torch.empty(n, dtype=torch.int32, device=cuda)
        pf.matchFromDevice(d_in, n, d_out)
        torch.cuda.synchronize()
        want_pos = torch.nonzero(d_out).flatten()
        want_id = d_out[want_pos]
        assert want_pos.numel() > 90_000
        d_id = torch.empty(n // 8, dtype=torch.int32, device=cuda)
        d_pos = torch.empty(n // 8, dtype=torch.int64, device=cuda)
        for rep in range(12):
            m = pf.matchFromDeviceReduce64(d_in, n, d_id, d_pos)
            assert m == want_pos.numel()
            assert torch.equal(d_pos[:m], want_pos), "rep %d: positions differ" % rep
            assert torch.equal(d_id[:m], want_id), "rep %d: ids differ" % rep
    # and the dense result itself against the oracle on a prefix
    orc = _oracle(pfile)
    k = 4 << 20
    assert np.array_equal(d_out[:k].cpu().numpy(), orc.match_shard(text[:k + 64], k))


def _noisy(text, seed, junk=b"Nn\x00\xff-"):
    rng = np.random.default_rng(seed)
    t = text.copy()
    idx = rng.integers(0, t.size, size=max(t.size // 700, 8))
    t[idx] = np.frombuffer(junk, dtype=np.uint8)[rng.integers(0, len(junk), size=idx.size)]
    return t


def test_small_alphabets_symbol_coded_prefilter(cuda, tmp_path):
    """2-bit x 8 (DNA) and 4-bit x 4 (hex) prefilter indices; bytes outside the alphabet, patterns
    shorter than K and input ends inside a K-gram go through the generic path."""
    from pfac_b200 import PFAC
    rng = np.random.default_rng(23)
    hexa = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)
    hexp = list({hexa[rng.integers(0, 16, size=int(rng.integers(1, 14)))].tobytes() for _ in range(900)})
    cases = [
        ("dna", synth.patterns_dna(2000, seed=31, min_len=3, max_len=24, short=12), "dna", None),
        ("dna_long_only", synth.patterns_dna(5000), "dna", None),
        ("hex", hexp, None, hexa),
    ]
    for name, pats, kind, alpha in cases:
        pfile = synth.write_pattern_file(str(tmp_path / (name + ".txt")), pats)
        orc = _oracle(pfile)
        with PFAC() as pf:
            pf.readPatternFromFile(pfile)
            info = pf.tableInfo()
            assert info["code_bits"] == (2 if kind == "dna" else 4), info
            for n in [7, 8, 9, 33, 511, 512, 513, 519, 520, 70_001, 1_500_003]:
                if kind == "dna":
                    text = synth.dna_bytes(900 + n, 0, n)
                else:
                    text = alpha[np.random.default_rng(n).integers(0, 16, size=n)].copy()
                if n > 1000:
                    synth.plant(text, 0, n, pats, 1234 + n, every=700)
                    text = _noisy(text, n)
                _check_all(pf, orc, text, cuda)
            text = _noisy(synth.dna_bytes(5, 0, 300_000), 5) if kind == "dna" else _noisy(
                alpha[np.random.default_rng(5).integers(0, 16, size=300_000)].copy(), 5)
            for owned in [1, 100_000, 299_990, 299_999, 300_000]:
                _check_all(pf, orc, text, cuda, n_owned=owned)


def test_host_apis_across_pipeline_chunks(cuda, tmp_path, monkeypatch):
    """PFAC_matchFromHost / PFAC_matchFromHostReduce split the input into chunks with a tail halo and
    run a two-stream H2D / kernel / D2H pipeline: matches that straddle chunk edges, pageable and pinned
    buffers, and the running output offset of the reduce variant."""
    from pfac_b200 import PFAC
    monkeypatch.setenv("PFAC_B200_HOST_CHUNK_MB", "4")   # force many chunks
    pats = synth.patterns_snort_like(3000, seed=41)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    n = (4 << 20) * 5 + 123_457
    text = synth.make_text("ascii", 4141, 0, n, n, pats, 1500)
    synth.plant(text, 0, n, pats, 77, every=1 << 30, boundary=4 << 20)  # straddle every chunk edge
    want = orc.match(text)
    wid, wpos = orc.reduce(want)
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        got = pf.matchFromHost(text)                                    # pageable numpy buffers
        assert np.array_equal(got, want)
        h_in = torch.from_numpy(text).pin_memory()
        h_out = torch.empty(n, dtype=torch.int32).pin_memory()
        pf.matchFromHost(h_in, h_out, size=n)                           # pinned buffers
        assert np.array_equal(h_out.numpy(), want)
        ids, pos = pf.matchFromHostReduce(text)
        assert np.array_equal(ids, wid) and np.array_equal(pos.astype(np.int64), wpos)
        for cut in (1, 4 << 20, (4 << 20) + 1, (8 << 20) - 1):
            ids, pos = pf.matchFromHostReduce(text[:cut])
            w = orc.match(text[:cut])
            wi, wp = orc.reduce(w)
            assert np.array_equal(ids, wi) and np.array_equal(pos.astype(np.int64), wp)


def test_pageable_and_pinned_host_buffers_agree(cuda, tmp_path, monkeypatch):
    """Pageable (malloc-style) buffers go through pinned staging + the host copy pool, pinned ones are
    DMA'd directly (pfac_api.cu hostDenseShard / hostReduceShard): every mix of the two, staging switched
    off, one copy thread, and sizes around the staged chunk give the oracle's result."""
    from pfac_b200 import PFAC
    monkeypatch.setenv("PFAC_B200_STAGE_CHUNK_MB", "1")
    pats = synth.patterns_snort_like(1500, seed=61)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    n = (1 << 20) * 7 + 4321
    text = synth.make_text("ascii", 6161, 0, n, n, pats, 900)
    synth.plant(text, 0, n, pats, 78, every=1 << 30, boundary=1 << 20)   # straddle every staged chunk edge
    want = orc.match(text)
    wid, wpos = orc.reduce(want)
    p_in = torch.from_numpy(text).pin_memory()
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        for pin_in in (False, True):
            for pin_out in (False, True):
                h_in = p_in if pin_in else text
                h_out = torch.zeros(n, dtype=torch.int32)
                h_out = h_out.pin_memory() if pin_out else h_out.numpy()
                pf.matchFromHost(h_in, h_out, size=n)
                got = h_out.numpy() if pin_out else h_out
                assert np.array_equal(got, want), (pin_in, pin_out)
            ids, pos = pf.matchFromHostReduce(h_in, size=n)
            assert np.array_equal(ids, wid) and np.array_equal(pos.astype(np.int64), wpos), pin_in
        for cut in (1, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, (2 << 20) + 7, 3 << 20):
            w = orc.match(text[:cut])
            assert np.array_equal(pf.matchFromHost(text[:cut].copy()), w), cut
            ids, pos = pf.matchFromHostReduce(text[:cut].copy())
            wi, wp = orc.reduce(w)
            assert np.array_equal(ids, wi) and np.array_equal(pos.astype(np.int64), wp), cut
        monkeypatch.setenv("PFAC_B200_STAGE", "0")
        assert np.array_equal(pf.matchFromHost(text), want)
        ids, pos = pf.matchFromHostReduce(text)
        assert np.array_equal(ids, wid) and np.array_equal(pos.astype(np.int64), wpos)
    monkeypatch.delenv("PFAC_B200_STAGE")
    monkeypatch.setenv("PFAC_B200_STAGE_CHUNK_MB", "3")
    with PFAC() as pf:                                                   # other chunk size, fresh buffers
        pf.readPatternFromFile(pfile)
        assert np.array_equal(pf.matchFromHost(text), want)
        ids, pos = pf.matchFromHostReduce(text)
        assert np.array_equal(ids, wid) and np.array_equal(pos.astype(np.int64), wpos)


@pytest.mark.parametrize("pinned", [False, True])
def test_host_result_sparse_and_dense_paths(cuda, tmp_path, monkeypatch, pinned):
    """PFAC_matchFromHost returns the dense array either by shipping it over PCIe or (default) by
    shipping the (id, position) pairs and letting the host zero-fill + scatter (pfac_api.cu
    hostDenseSparse).  A text whose first third is match-dense (more than one match per 16 positions:
    those chunks fall back to the dense kernel), the rest sparse; stale garbage in the output buffer;
    both modes and the transfer counters."""
    from pfac_b200 import PFAC
    monkeypatch.setenv("PFAC_B200_HOST_CHUNK_MB", "1")
    monkeypatch.setenv("PFAC_B200_STAGE_CHUNK_MB", "1")
    pats = synth.patterns_snort_like(800, seed=91) + [b"zz", b"z"]
    pats = list(dict.fromkeys(pats))
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    n = (1 << 20) * 6 + 777
    text = synth.make_text("ascii", 9191, 0, n, n, pats, 700)
    dense_part = (1 << 20) * 2 + 5000
    rng = np.random.default_rng(3)
    text[:dense_part][rng.random(dense_part) < 0.3] = ord("z")            # 30 % of the positions match
    want = orc.match(text)
    assert (want[:dense_part] > 0).mean() > 0.25 and (want[dense_part:] > 0).mean() < 0.05
    h_in = torch.from_numpy(text).pin_memory() if pinned else text
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        moved = {}
        for mode in ("sparse", "dense"):
            if mode == "dense":
                monkeypatch.setenv("PFAC_B200_HOST_RESULT", "dense")
            h_out = torch.full((n,), -5, dtype=torch.int32)
            h_out = h_out.pin_memory() if pinned else h_out.numpy()
            pf.matchFromHost(h_in, h_out, size=n)
            got = h_out.numpy() if pinned else h_out
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, (mode, int(bad[0]), int(got[bad[0]]), int(want[bad[0]]))
            moved[mode] = pf.lastHostTransfer()
            for cut in (1, 15, (1 << 20) - 1, (1 << 20) + 1):
                assert np.array_equal(pf.matchFromHost(text[:cut].copy()), orc.match(text[:cut])), (mode, cut)
        assert moved["dense"][1] == 4 * n and moved["dense"][0] >= n
        # sparse: the two dense chunks came back whole, the rest as pairs
        assert 4 * (2 << 20) <= moved["sparse"][1] < 4 * (3 << 20) + 8 * int((want > 0).sum())


def test_compiled_patterns_file(cuda, tmp_path, monkeypatch):
    """PFAC_saveCompiledPatterns / PFAC_loadCompiledPatterns: a second handle loaded from the file is
    indistinguishable from one that read the pattern file (table info, dump, results); a stored layout
    for another filter policy is recompiled from the stored automaton; a bad file changes nothing."""
    from pfac_b200 import PFAC, PFACError, Status
    pats = synth.patterns_snort_like(3000, seed=101)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    blob = str(tmp_path / "p.pfacb")
    text = synth.make_text("ascii", 1011, 0, 300_000, 300_000, pats, 90)
    with PFAC() as a, PFAC() as b:
        a.readPatternFromFile(pfile)
        a.saveCompiledPatterns(blob)
        with pytest.raises(PFACError) as e:
            b.saveCompiledPatterns(str(tmp_path / "none.pfacb"))
        assert e.value.status == Status.PATTERNS_NOT_READY
        b.loadCompiledPatterns(blob)
        assert a.tableInfo() == b.tableInfo()
        a.dumpTransitionTable(str(tmp_path / "a.txt"))
        b.dumpTransitionTable(str(tmp_path / "b.txt"))
        assert open(str(tmp_path / "a.txt"), "rb").read() == open(str(tmp_path / "b.txt"), "rb").read()
        _check_all(b, orc, text, cuda)
        # garbage: INVALID_PARAMETER, the loaded patterns stay
        bad = str(tmp_path / "bad.pfacb")
        open(bad, "wb").write(open(blob, "rb").read()[:-7])
        with pytest.raises(PFACError) as e:
            b.loadCompiledPatterns(bad)
        assert e.value.status == Status.INVALID_PARAMETER
        with pytest.raises(PFACError) as e:
            b.loadCompiledPatterns(str(tmp_path / "missing.pfacb"))
        assert e.value.status == Status.FILE_OPEN_ERROR
        _check_all(b, orc, text, cuda)
        # other policy than the file's: layouts are recompiled, results unchanged
        monkeypatch.setenv("PFAC_B200_FILTER", "exact")
        b.loadCompiledPatterns(blob)
        assert b.tableInfo()["hashed_filter"] == 0
        _check_all(b, orc, text, cuda)


def test_reference_test_programs_run_unchanged(cuda, tmp_path, golden_dir):
    """The drop-in check: the reference's own test programs (PFAC/test/simple_example.cpp,
    simple_example_reduce.cpp), compiled unchanged against this
    repo's include/PFAC.h and linked with -lpfac to this repo's libpfac.so (oracle/Makefile `progs`, built
    where /root/reference exists; the binaries travel in oracle/_ref/progs).  Their output must be what
    the oracle says."""
    import shutil
    import subprocess
    progs = os.path.join(ROOT, "oracle", "_ref", "progs")
    if not os.path.exists(os.path.join(progs, "simple_example")):
        pytest.skip("oracle/_ref/progs not built (needs /root/reference at build time)")
    # the programs hard-code ../test/pattern/... and ../test/data/... relative to the working directory
    (tmp_path / "test" / "pattern").mkdir(parents=True)
    (tmp_path / "test" / "data").mkdir(parents=True)
    (tmp_path / "bin").mkdir()
    for f in ("example_pattern", "example_pattern2"):
        shutil.copy(os.path.join(golden_dir, f), tmp_path / "test" / "pattern" / f)
    for f in ("example_input", "example_input2"):
        shutil.copy(os.path.join(golden_dir, f), tmp_path / "test" / "data" / f)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "pfac_b200", "lib") + ":" +
               os.environ.get("LD_LIBRARY_PATH", ""), OMP_NUM_THREADS="4")

    def run(name, *args):
        r = subprocess.run([os.path.join(progs, name)] + list(args), cwd=str(tmp_path / "bin"), env=env,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, name + ":\n" + r.stdout + r.stderr
        assert "Error" not in r.stdout, r.stdout
        return r.stdout

    orc = _oracle(os.path.join(golden_dir, "example_pattern"))
    text = np.fromfile(os.path.join(golden_dir, "example_input"), dtype=np.uint8)
    want = orc.match(text)
    lines = ["At position %4d, match pattern %d" % (i, want[i]) for i in np.flatnonzero(want)]
    out = run("simple_example")
    assert [l for l in out.splitlines() if l.startswith("At position")] == lines
    out = run("simple_example_reduce")
    assert "number of matched = %d" % len(lines) in out
    assert [l for l in out.splitlines() if l.startswith("At position")] == lines
    # SimpleMultiGPU_pthread (reference test/SimpleMultiGPU_pthread.cpp:188-197): two host threads, one
    # handle each (on one GPU both use device 0), two dictionaries at once; it writes match1/match2 and
    # table1/table2 into its working directory
    run("SimpleMultiGPU_pthread")
    for k, pat, inp in ((1, "example_pattern", "example_input"), (2, "example_pattern2", "example_input2")):
        o = _oracle(os.path.join(golden_dir, pat))
        w = o.match(np.fromfile(os.path.join(golden_dir, inp), dtype=np.uint8))
        got = open(str(tmp_path / "bin" / ("match%d" % k))).read().splitlines()
        assert got == ["At position %4d, match pattern %d" % (i, w[i]) for i in np.flatnonzero(w)], k
        o.dump(str(tmp_path / ("want_table%d" % k)))
        assert open(str(tmp_path / "bin" / ("table%d" % k)), "rb").read() == \
            open(str(tmp_path / ("want_table%d" % k)), "rb").read()
    # omp_PFAC is built and linked the same way (oracle/Makefile) but not run: it sizes its segments with
    # `((int)minGlobalMem) >> 3` (omp_PFAC.cpp:205), which overflows on a 180 GB device before the
    # library is ever called.


def test_reference_profiling_harness(cuda, tmp_path):
    """The reference's own benchmark program (test/profiling.cpp: its metric is Gbps = 8 * input_size /
    time, :296-322), compiled unchanged and linked to this library, on 128 MiB of the C2 workload:
    device mode (-TD: PFAC_matchFromDevice between CUDA events) and host mode (-TH: PFAC_matchFromHost
    on malloc'ed buffers), time-driven and space-driven.  Its match count must be the oracle's; its
    Gbps lines go to gpurun_out/ for RESULTS.md."""
    import re
    import subprocess
    prog = os.path.join(ROOT, "oracle", "_ref", "progs", "profiling")
    if not os.path.exists(prog):
        pytest.skip("oracle/_ref/progs not built (needs /root/reference at build time)")
    n = 128 << 20   # the program computes input_size * 8 in an int: 128 MiB is the most it can report
    pats = synth.patterns_c2(1000)
    pfile = synth.write_pattern_file(str(tmp_path / "c2.pat"), pats)
    text = synth.make_text("random", synth.SEED_BASE + 2, 0, n, n, pats, every=4096)
    text.tofile(str(tmp_path / "c2.txt"))
    want = int((_oracle(pfile).match(text) > 0).sum())
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "pfac_b200", "lib") + ":" +
               os.environ.get("LD_LIBRARY_PATH", ""))
    results = {}
    for label, args in (("device", ["-TD"]), ("host", ["-TH"]), ("device_space_driven", ["-TD", "-S"])):
        r = subprocess.run([prog, "-P", pfile, "-I", str(tmp_path / "c2.txt")] + args, env=env, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0 and "Error" not in r.stdout, r.stdout + r.stderr
        assert "The number of matched is %d" % want in r.stdout, r.stdout
        gbps = float(re.search(r"The throughput is ([0-9.]+) Gbps", r.stdout).group(1))
        assert gbps > 0
        results[label] = {"Gbps": gbps, "GBps": gbps / 8.0}
    from tests import configs
    configs.append_result({"test": "reference_profiling_harness", "bytes": n, "matches": want, "results": results,
                           "note": "one cold launch each (the program times a single call, table staging and "
                                   "first-touch included)"})


def test_multi_gpu_driver_one_process(cuda, tmp_path, monkeypatch):
    """PFAC_mgpu_* (the library-level replacement of reference test/omp_PFAC.cpp): shards + halo, one
    host thread per handle, runs placed at the exclusive scan of the per-GPU counts.  Uses every
    visible GPU; with a single GPU two handles share it, which exercises the same host logic.
    Like omp_PFAC.cpp:397-439 the result is also compared with the single-handle run."""
    from pfac_b200 import PFAC, PFACMultiGPU
    monkeypatch.setenv("PFAC_B200_HOST_CHUNK_MB", "2")
    pats = synth.patterns_snort_like(2000, seed=51)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    n = 9_000_017
    text = synth.make_text("ascii", 5151, 0, n, n, pats, 900)
    want = orc.match(text)
    wid, wpos = orc.reduce(want)
    ndev = torch.cuda.device_count()
    for devices in ([0, 0, 0], list(range(ndev)) if ndev > 1 else [0, 0]):
        with PFACMultiGPU(devices) as mg:
            mg.readPatternFromFile(pfile)
            assert np.array_equal(mg.matchFromHost(text), want), devices
            ids, pos = mg.matchFromHostReduce64(text)
            assert np.array_equal(ids, wid) and np.array_equal(pos, wpos), devices
            ids, pos = mg.matchFromHostReduce64(text[:5])
            w5 = orc.match(text[:5])
            assert ids.tolist() == w5[w5 > 0].tolist()
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)
        assert np.array_equal(pf.matchFromHost(text), want)
        pf_mem = pf.tableInfo()
        assert pf_mem["num_patterns"] == len(pats)


def test_threads_sharing_one_handle_and_private_streams(cuda, tmp_path):
    """Reference r1.2 supports several host threads on one handle (NOTICE; user guide r1.2 p.2).  Here:
    concurrent dense and reduce calls from four threads on one handle, and PFAC_setStream with a
    caller-owned stream, must all give the oracle's result."""
    import threading
    from pfac_b200 import PFAC
    pats = synth.patterns_c2(500, seed=61)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    orc = _oracle(pfile)
    texts = [synth.make_text("random", 700 + i, 0, 3_000_001 + 7 * i, 3_000_001 + 7 * i, pats, 600) for i in range(4)]
    wants = [orc.match(t) for t in texts]
    errors = []
    with PFAC() as pf:
        pf.readPatternFromFile(pfile)

        def work(i):
            try:
                torch.cuda.set_device(0)
                for rep in range(3):
                    got = _dev_match(pf, texts[i], cuda)
                    if not np.array_equal(got, wants[i]):
                        errors.append("dense thread %d rep %d" % (i, rep))
                    m, ids, pos = _dev_reduce(pf, texts[i], cuda, pos64=True)
                    wid, wpos = orc.reduce(wants[i])
                    if m != wid.size or not np.array_equal(ids[:m], wid) or not np.array_equal(pos[:m], wpos):
                        errors.append("reduce thread %d rep %d" % (i, rep))
            except Exception as e:  # noqa: BLE001
                errors.append("thread %d: %r" % (i, e))

        threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors

        # caller-owned stream: work is ordered on it, not on the legacy default stream
        stream = torch.cuda.Stream()
        pf.setStream(stream)
        with torch.cuda.stream(stream):
            d_in = torch.from_numpy(texts[0]).to(cuda, non_blocking=True)
            d_out = torch.empty(texts[0].size, dtype=torch.int32, device=cuda)
            pf.matchFromDevice(d_in, texts[0].size, d_out)
            host = d_out.to("cpu", non_blocking=False)
        stream.synchronize()
        assert np.array_equal(host.numpy(), wants[0])
        pf.setStream(None)
        assert np.array_equal(_dev_match(pf, texts[1], cuda), wants[1])


def test_patterns_from_arrays_with_newlines(cuda):
    """PFAC_readPatternFromArrays: patterns may contain 0x0A, which the reference's file grammar cannot
    express (user guide r1.2 p.26); checked against a brute-force restatement of the semantics."""
    from pfac_b200 import PFAC
    from tests.helpers import brute_force_match
    rng = np.random.default_rng(12)
    pats = [b"a\nb", b"\n", b"\n\n\x00", b"line1\nline2\n", b"xyz", b"xy", b"\x00\n\xff", b"q", b"GET /\r\n\r\n"]
    text = rng.integers(0, 256, size=40_000, dtype=np.uint8)
    for p in pats * 40:
        at = int(rng.integers(0, text.size - len(p)))
        text[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
    want = brute_force_match(pats, text)
    with PFAC() as pf:
        pf.readPatternFromArrays(pats)
        assert np.array_equal(_dev_match(pf, text, cuda), want)
        m, ids, pos = _dev_reduce(pf, text, cuda)
        nz = np.flatnonzero(want)
        assert m == nz.size and np.array_equal(pos[:m], nz) and np.array_equal(ids[:m], want[nz])
