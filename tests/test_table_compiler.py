"""Host logic of the product library, no GPU: the C++ table compiler (pfac_b200/csrc/
pfac_table.cpp) reached through the C ABI (PFAC_tableCompile*, include/PFAC_ext.h).

  * dump is byte-identical to the reference's PFAC_dumpTransitionTable (goldens),
  * state numbering / counts equal the oracle's,
  * the emitted device layout (root row, prefilter bitmap, hot/cold hash rows), walked by a
    Python restatement of the kernels' lookup sequence, reproduces the oracle's dense result.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import Oracle
from pfac_b200 import PFACError, Status, TableCompiler
from workloads import synth
from tests.helpers import brute_force_match, emulate_layout_walk, read_patterns


def _layout(tc):
    return tc.layout()


@pytest.mark.parametrize("fixture", ["example_pattern", "example_pattern2"])
def test_dump_matches_reference_golden(golden_dir, tmp_path, fixture):
    tc = TableCompiler(os.path.join(golden_dir, fixture))
    out = tmp_path / "t.txt"
    tc.dump(str(out))
    assert out.read_bytes() == open(os.path.join(golden_dir, fixture + ".dump"), "rb").read()


@pytest.mark.parametrize("case", ["c2", "snort", "dna"])
def test_dump_digest_matches_reference(golden_dir, tmp_path, case):
    g = np.load(os.path.join(golden_dir, "synth_%s.npz" % case))
    tc = TableCompiler(os.path.join(golden_dir, "synth_%s.pat" % case))
    out = tmp_path / "t.txt"
    tc.dump(str(out))
    assert hashlib.sha256(out.read_bytes()).hexdigest() == str(g["dump_sha256"])
    info = tc.info()
    assert info["num_states"] == int(g["num_states"])
    assert info["max_pattern_len"] == int(g["max_pattern_len"])


def test_info_readme_example(golden_dir):
    info = TableCompiler(os.path.join(golden_dir, "example_pattern")).info()
    assert (info["num_patterns"], info["num_states"], info["initial_state"], info["max_pattern_len"]) == (4, 11, 5, 4)
    assert info["num_leaves"] == 3 and info["num_edges"] == 9 and info["root_fanout"] == 3
    assert info["pre2_bits_set"] == 51
    assert info["code_bits"] == 4 and info["gram_len"] == 4  # alphabet {A,B,D,E,G}: 4-bit symbol codes
    # 4-grams over 5 symbols that can yield a result: AB??, ED?? (2 x 25) + BEDE


@pytest.mark.parametrize("hot_kb", [0, 1, 24, 512])
@pytest.mark.parametrize("case", ["c2", "snort", "dna"])
def test_layout_walk_equals_oracle(golden_dir, case, hot_kb):
    """Every hot/cold split must give the oracle's result (PFAC_SPACE_DRIVEN == hot budget 0)."""
    g = np.load(os.path.join(golden_dir, "synth_%s.npz" % case))
    pfile = os.path.join(golden_dir, "synth_%s.pat" % case)
    pats = read_patterns(pfile)
    n = 6000
    text = synth.make_text(str(g["kind"]), int(g["seed"]), 0, n, n, pats, 128)
    o = Oracle(pfile)
    want = o.match(text)
    assert (want > 0).sum() > 10
    tc = TableCompiler(pfile, hot_budget_bytes=hot_kb * 1024)
    info = tc.info()
    L = _layout(tc)
    assert L["hot"].shape[0] * 16 <= max(hot_kb * 1024, 0)
    if hot_kb == 0:
        assert info["hot_depth"] == 1 and info["hot_buckets"] == 0 and not info["chains_hot"]
    if hot_kb == 512:
        assert info["hot_depth"] == info["max_depth"] + 1 and info["chains_hot"]  # everything fits
        assert info["next2_hot"]
        if info["gram_len"] == 2:
            assert info["hot_buckets"] == info["hash_edges"]
        else:  # edges of depth < K serve only the generic path: always cold
            assert 0 < info["hot_buckets"] < info["hash_edges"]
            assert info["hot_buckets"] + info["cold_buckets"] // 2 == info["hash_edges"]
    if hot_kb == 0:
        assert not info["next2_hot"]
    assert info["num_chains"] > 0 and info["hash_edges"] < info["num_edges"]
    got = np.array([emulate_layout_walk(L, o.num_patterns, text, i) for i in range(n)], dtype=np.int32)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, "first mismatch at %d: got %d want %d" % (bad[0], got[bad[0]], want[bad[0]])


def test_layout_root_and_prefilter_against_dense_table(golden_dir):
    pfile = os.path.join(golden_dir, "synth_snort.pat")
    o = Oracle(pfile)
    T = o.dense_table()
    tc = TableCompiler(pfile)
    L = tc.layout()
    root, pre2, hot, cold = L["root"], L["pre2"], L["hot"], L["cold"]
    init, k = o.initial_state, o.num_patterns
    assert np.array_equal(root, T[init])
    assert tc.info()["code_bits"] == 8 and tc.info()["gram_len"] == 2
    nset = 0
    for c1 in range(256):
        for c0 in range(256):
            s = T[init, c0]
            idx = c0 | (c1 << 8)
            bit = (int(pre2[idx >> 5]) << (idx & 31) >> 31) & 1
            expect = s >= 0 and (s <= k or T[s, c1] >= 0)
            assert bool(bit) == bool(expect), (c0, c1)
            if bit:  # next2 / best2 are indexed by the rank of the bit, in idx order
                if (idx & 31) == 0:
                    assert int(L["rank2"][idx >> 5]) == nset
                v = int(L["next2"][nset])
                if T[s, c1] < 0:
                    assert v == 0xFFFFFFFF
                elif not (v & 0x80000000):
                    assert (v & 0x3FFFFFFF) == T[s, c1]
                nset += 1
    assert nset == tc.info()["pre2_bits_set"]
    # every hash entry is unique; chain compression accounts for the missing transitions
    info = tc.info()
    keys = np.concatenate([hot[:, 0], hot[:, 2], cold[:, 0], cold[:, 2]])
    keys = keys[keys != 0xFFFFFFFF]
    assert keys.size == np.unique(keys).size == info["hash_edges"]
    chain_len = int(L["chains"][:info["num_chains"], 1].sum())
    depth1 = int((L["next2"] != 0xFFFFFFFF).sum())  # (root child, c1) edges live in next2
    assert info["num_edges"] == info["root_fanout"] + depth1 + info["hash_edges"] + chain_len
    assert info["has_best2"] == 0  # K == 2: 1-byte patterns are told by the root row


def test_duplicates_prefixes_and_one_byte_patterns():
    """Last edge wins in the matching table (reference PFAC.cpp:376-381); a 1-byte pattern makes
    its byte always a candidate; a final state can have out-edges."""
    image = b"AB\nA\nABC\nB\n"
    o = Oracle(image=image)
    tc = TableCompiler(image=image)
    L = _layout(tc)
    text = np.frombuffer(b"ABCABXBA", dtype=np.uint8)
    got = [emulate_layout_walk(L, o.num_patterns, text, i) for i in range(text.size)]
    assert got == o.match(text).tolist() == [3, 4, 0, 1, 4, 0, 4, 2]


def test_parser_errors_and_quirks(tmp_path):
    with pytest.raises(PFACError) as e:
        TableCompiler(image=b"AB\n\nCD\n")
    assert e.value.status == Status.INVALID_PARAMETER
    with pytest.raises(PFACError) as e:
        TableCompiler(pattern_file=str(tmp_path / "does_not_exist"))
    assert e.value.status == Status.FILE_OPEN_ERROR
    assert TableCompiler(image=b"AB\nCD").info()["num_patterns"] == 1     # unterminated tail dropped
    assert TableCompiler(image=b"AB\n\n\n").info()["num_patterns"] == 1   # trailing blank lines ok
    info = TableCompiler(image=b"").info()
    assert info["num_patterns"] == 0 and info["pre2_bits_set"] == 0


def test_state_numbering_matches_oracle_on_binary_patterns(tmp_path):
    pats = synth.patterns_c2(500, seed=77, min_len=1, max_len=12, prefix_pairs=100)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    o = Oracle(pfile)
    tc = TableCompiler(pfile)
    a, b = tmp_path / "a.txt", tmp_path / "b.txt"
    o.dump(str(a))
    tc.dump(str(b))
    assert a.read_bytes() == b.read_bytes()


def _small_alphabet_cases():
    rng = np.random.default_rng(17)
    dna = synth.patterns_dna(300, seed=5, min_len=3, max_len=14, short=10)          # < K-long patterns too
    hexa = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)
    hexp = list({hexa[rng.integers(0, 16, size=int(rng.integers(1, 11)))].tobytes() for _ in range(400)})
    return [("dna", dna, np.frombuffer(b"ACGT", dtype=np.uint8), 2, 8),
            ("hex", hexp, hexa, 4, 4)]


@pytest.mark.parametrize("hot_kb", [0, 512])
@pytest.mark.parametrize("which", [0, 1])
def test_symbol_coded_prefilter_small_alphabets(which, hot_kb):
    """Alphabets of <= 4 / <= 16 symbols use 2-bit x 8 / 4-bit x 4 prefilter indices.  Text with bytes
    outside the alphabet, patterns shorter than K and walks cut off by the end of the input must take
    the generic path and still give the oracle's result."""
    name, pats, alpha, bits, k = _small_alphabet_cases()[which]
    image = synth.pattern_file_image(pats)
    o = Oracle(image=image)
    tc = TableCompiler(image=image, hot_budget_bytes=hot_kb * 1024)
    info = tc.info()
    assert (info["code_bits"], info["gram_len"]) == (bits, k)
    assert info["has_best2"] == 1
    L = tc.layout()
    rng = np.random.default_rng(3)
    n = 5000
    text = alpha[rng.integers(0, alpha.size, size=n)].copy()
    text[rng.integers(0, n, size=60)] = ord("N")          # bytes that occur in no pattern
    text[rng.integers(0, n, size=20)] = 0
    for p in pats[:40]:                                    # plant some, including at the very end
        at = int(rng.integers(0, n - len(p)))
        text[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
    longest = max(pats, key=len)
    text[n - len(longest) + 1:] = np.frombuffer(longest[:-1], dtype=np.uint8)
    want = o.match(text)
    got = np.array([emulate_layout_walk(L, o.num_patterns, text, i) for i in range(n)], dtype=np.int32)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, "%s: first mismatch at %d: got %d want %d" % (name, bad[0], got[bad[0]], want[bad[0]])
    assert (want > 0).sum() > 100


def test_patterns_from_arrays_any_byte():
    """PFAC_tableCompileArrays: explicit lengths, so 0x0A (and everything else) may occur in a pattern.
    For newline-free sets the result is the file form's, including the dump."""
    rng = np.random.default_rng(9)
    pats = [b"a\nb", b"\n", b"\n\n\x00", b"line1\nline2\n", b"xyz", b"xy", b"\x00\n\xff", b"q"]
    tc = TableCompiler(patterns=pats)
    L = tc.layout()
    text = np.frombuffer(b"a\nbxyz\n\n\x00line1\nline2\nq\x00\n\xffxy\n", dtype=np.uint8)
    want = brute_force_match(pats, text)
    got = np.array([emulate_layout_walk(L, len(pats), text, i) for i in range(text.size)], dtype=np.int32)
    assert np.array_equal(got, want) and set(want) >= {1, 2, 3, 4, 5, 7, 8}
    big = rng.integers(0, 256, size=3000, dtype=np.uint8)
    for p in pats * 5:
        at = int(rng.integers(0, big.size - len(p)))
        big[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
    want = brute_force_match(pats, big)
    got = np.array([emulate_layout_walk(L, len(pats), big, i) for i in range(big.size)], dtype=np.int32)
    assert np.array_equal(got, want)
    # same machine as the file form when no pattern holds a newline
    plain = synth.patterns_c2(120, seed=3, min_len=1, max_len=9, prefix_pairs=20)
    a = TableCompiler(patterns=plain)
    b = TableCompiler(image=synth.pattern_file_image(plain))
    assert a.info() == b.info()
    for k, v in a.layout().items():
        assert np.array_equal(v, b.layout()[k]), k
    with pytest.raises(PFACError) as e:
        TableCompiler(patterns=[b"ab", b""])
    assert e.value.status == Status.INVALID_PARAMETER


@pytest.mark.parametrize("policy", ["hash", "auto"])
def test_hashed_four_gram_filter_never_hides_a_match(monkeypatch, policy):
    """Large byte-alphabet dictionaries get a hashed 4-gram filter as the per-position test
    (pfac_table.h hfilt).  It may pass too much, never too little: patterns of 1, 2 and 3 bytes,
    matches at the very end of the input whatever bytes lie behind it, duplicates and prefixes."""
    monkeypatch.setenv("PFAC_B200_FILTER", policy)
    rng = np.random.default_rng(17)
    n_pat = 300 if policy == "hash" else 9000      # sparse table: one bit per lookup; dense: two
    pats = synth.patterns_snort_like(n_pat, seed=23)
    pats += [b"q", b"zq", b"~z", b"xyz", b"\x00\x01\x02", b"e", b"th", b"the", b"them", b"\xff"]
    pats = list(dict.fromkeys(pats))
    tc = TableCompiler(patterns=pats, hot_budget_bytes=64 * 1024)
    info = tc.info()
    assert info["hashed_filter"] == (1 if policy == "hash" else 2)   # 310 / 9,010 patterns: sparse / dense table
    L = tc.layout()
    assert L["hfilt"].size == info["hfilt_words"] == (8192 if policy == "hash" else 16384)
    assert 0 < info["hfilt_bits_set"] < 32 * info["hfilt_words"] * 0.6
    n = 20000
    text = synth.make_text("ascii", 99, 0, n, n, pats, 64)
    for p in (b"q", b"zq", b"xyz", b"th", b"the", b"them", b"\xff", b"\x00\x01\x02"):   # shorts at the end
        text[n - len(p):] = np.frombuffer(p, dtype=np.uint8)
        want = brute_force_match(pats, text[n - 8:])
        for pad in (0, 0xA5, 0xFF):
            got = [emulate_layout_walk(L, len(pats), text[n - 8:], i, pad=pad) for i in range(8)]
            assert got == want.tolist(), (p, pad)
    want = brute_force_match(pats, text)
    got = np.array([emulate_layout_walk(L, len(pats), text, i, pad=0x5A) for i in range(n)], dtype=np.int32)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, "first mismatch at %d: got %d want %d" % (bad[0], got[bad[0]], want[bad[0]])
    assert (want > 0).sum() > 300
    # the filter does reject: most positions never reach the exact tables
    t = text.astype(np.uint64)
    x = t[:-3] | (t[1:-2] << 8) | (t[2:-1] << 16) | (t[3:] << 24)
    if info["hashed_filter"] == 2:
        w = L["hfilt"][(x & (info["hfilt_words"] - 1)).astype(np.int64)].astype(np.uint64)
    else:
        w = L["hfilt"][((((x * 0x9E3779B1) & 0xFFFFFFFF) >> 2) & 8191).astype(np.int64)].astype(np.uint64)
    passed = ((w << (((x * 0x85EBCA6B) >> 32) & 31)) >> 31) & 1
    if info["hashed_filter"] == 2:
        passed &= ((w << (((x * 0xC2B2AE35) >> 32) & 31)) >> 31) & 1
    assert passed.mean() < 0.5
    assert np.all(passed[np.flatnonzero(want[:-3] > 0)] == 1)
    # exact policy: no filter, same results
    monkeypatch.setenv("PFAC_B200_FILTER", "exact")
    tc2 = TableCompiler(patterns=pats, hot_budget_bytes=64 * 1024)
    assert tc2.info()["hashed_filter"] == 0 and tc2.layout()["hfilt"].size == 0
    # no room for the filter: falls back to the exact stage
    monkeypatch.setenv("PFAC_B200_FILTER", policy)
    assert TableCompiler(patterns=pats, hot_budget_bytes=16 * 1024).info()["hashed_filter"] == 0


def test_row_indexed_filter_at_32_kb(monkeypatch):
    """Dense dictionaries get the row-indexed two-bit filter with 16,384 words when the budget holds 64 KB and with
    8,192 words (word = c0 | (c1 & 31) << 8) when it holds only 32 KB: same answers either way."""
    monkeypatch.delenv("PFAC_B200_FILTER", raising=False)
    pats = synth.patterns_snort_like(9000, seed=23) + [b"q", b"zq", b"xyz", b"e", b"th", b"\xff"]
    pats = list(dict.fromkeys(pats))
    n = 12000
    text = synth.make_text("ascii", 199, 0, n, n, pats, 64)
    for p in (b"q", b"zq", b"xyz", b"\xff"):
        text[n - len(p):] = np.frombuffer(p, dtype=np.uint8)
    want = brute_force_match(pats, text)
    for budget, words in ((48 * 1024, 8192), (80 * 1024, 16384)):
        tc = TableCompiler(patterns=pats, hot_budget_bytes=budget)
        info = tc.info()
        assert info["hashed_filter"] == 2 and info["hfilt_words"] == words, info
        L = tc.layout()
        got = np.array([emulate_layout_walk(L, len(pats), text, i, pad=0xC3) for i in range(n)], dtype=np.int32)
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, "budget %d: first mismatch at %d" % (budget, bad[0])
        # patterns shorter than the gram fill words of their own first byte and no others (a 3-byte pattern sets the
        # bits of its 256 continuations in one word: all of them)
        hf = L["hfilt"]
        full = np.flatnonzero(hf == 0xFFFFFFFF)
        assert full.size and set((full & 0xFF).tolist()) <= {p[0] for p in pats if len(p) < 4}


def test_pair_filter_never_hides_a_match(monkeypatch):
    """Sparse byte dictionaries whose patterns all have three bytes or more get the pair filter: one lookup
    for the start positions q and q+1, keyed by the three text bytes they share (pfac_table.cpp).  It may pass
    too much, never too little: 3-byte patterns, matches at either parity, adjacent and overlapping matches,
    matches at the very end of the input whatever bytes lie behind it."""
    monkeypatch.delenv("PFAC_B200_FILTER", raising=False)
    pats = synth.patterns_c2(600) + [b"xyz", b"abc", b"bcd", b"\x00\x01\x02", b"\xff\xff\xff", b"abcd", b"zzzz", b"zzz"]
    pats = list(dict.fromkeys(pats))
    tc = TableCompiler(patterns=pats, hot_budget_bytes=64 * 1024)
    info = tc.info()
    assert info["hashed_filter"] == 3 and info["hfilt_words"] == 8192
    assert 0 < info["hfilt_bits_set"] <= 8192 * 32 // 64
    L = tc.layout()
    rng = np.random.default_rng(5)
    n = 30000
    text = synth.make_text("random", 77, 0, n, n, pats, 97)     # odd period: both parities get planted patterns
    for at, p in ((100, b"abcd"), (103, b"xyz"), (201, b"abc"), (202, b"bcd"), (300, b"zzzzzzz"), (401, b"\xff\xff\xff\xff")):
        text[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
    for p in (b"xyz", b"abcd", b"zzz", b"\x00\x01\x02"):        # at the very end, both parities of the start
        for shift in (0, 1):
            t = text[:n - shift].copy()
            t[len(t) - len(p):] = np.frombuffer(p, dtype=np.uint8)
            want = brute_force_match(pats, t[len(t) - 8:])
            for pad in (0, 0xA5, 0xFF):
                got = [emulate_layout_walk(L, len(pats), t[len(t) - 8:], i, pad=pad) for i in range(8)]
                assert got == want.tolist(), (p, shift, pad)
    want = brute_force_match(pats, text)
    got = np.array([emulate_layout_walk(L, len(pats), text, i, pad=0x5A) for i in range(n)], dtype=np.int32)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, "first mismatch at %d: got %d want %d" % (bad[0], got[bad[0]], want[bad[0]])
    assert (want > 0).sum() > 250
    # the filter does reject: few pairs pass
    t = text.astype(np.uint64)
    y = t[1:n - 3:2] | (t[2:n - 2:2] << 8) | (t[3:n - 1:2] << 16)
    h = (y * 0x9E3779B1) & 0xFFFFFF
    w = L["hfilt"][((h >> 2) & 8191).astype(np.int64)].astype(np.uint64)
    assert (((w << (h >> 19)) >> 31) & 1).mean() < 0.05
    # a 1- or 2-byte pattern rules the pair filter out; so does the nopair policy
    assert TableCompiler(patterns=pats + [b"q"], hot_budget_bytes=64 * 1024).info()["hashed_filter"] == 1
    assert TableCompiler(patterns=pats + [b"qr"], hot_budget_bytes=64 * 1024).info()["hashed_filter"] == 1
    monkeypatch.setenv("PFAC_B200_FILTER", "nopair")
    assert TableCompiler(patterns=pats, hot_budget_bytes=64 * 1024).info()["hashed_filter"] == 1


@pytest.mark.parametrize("seed", range(6))
def test_random_dictionaries_both_first_stages(monkeypatch, seed):
    """Property check of the compiled layout against the brute-force semantics on random dictionaries:
    random alphabets (2..256 symbols, so all three symbol codings occur), pattern lengths 1..12 with
    shared prefixes, texts drawn from the same alphabet plus foreign bytes; exact and hashed first
    stage; every budget from nothing to everything in shared memory."""
    rng = np.random.default_rng(1000 + seed)
    for trial in range(8):
        asize = int(rng.choice([2, 3, 4, 5, 9, 16, 17, 40, 256]))
        alpha = rng.choice(256, size=asize, replace=False).astype(np.uint8)
        npat = int(rng.integers(1, 60))
        pats = []
        for _ in range(npat):
            ln = int(rng.integers(1, 13))
            body = alpha[rng.integers(0, asize, size=ln)].tobytes()
            if pats and rng.random() < 0.4:
                body = (pats[int(rng.integers(0, len(pats)))] + body)[:14]
            pats.append(body)
        pats = list(dict.fromkeys(pats))
        n = 700
        text = alpha[rng.integers(0, asize, size=n)].copy()
        text[rng.integers(0, n, size=25)] = rng.integers(0, 256, size=25).astype(np.uint8)
        for p in pats[:12]:
            at = int(rng.integers(0, n - len(p) + 1))
            text[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
        tail = pats[int(rng.integers(0, len(pats)))]
        text[n - len(tail):] = np.frombuffer(tail, dtype=np.uint8)        # a match ending at the last byte
        want = brute_force_match(pats, text)
        for policy in ("exact", "hash"):
            monkeypatch.setenv("PFAC_B200_FILTER", policy)
            for budget in (0, 40 * 1024, 512 * 1024):
                tc = TableCompiler(patterns=pats, hot_budget_bytes=budget)
                L = tc.layout()
                info = tc.info()
                hashable = info["code_bits"] == 8 or (info["code_bits"] == 2 and info["code_shift"] >= 0)
                assert bool(info["hashed_filter"]) == (policy == "hash" and hashable and budget >= 32768)
                got = np.array([emulate_layout_walk(L, len(pats), text, i, pad=int(rng.integers(0, 256)))
                                for i in range(n)], dtype=np.int32)
                bad = np.flatnonzero(got != want)
                assert bad.size == 0, (seed, trial, policy, budget, int(bad[0]), int(got[bad[0]]), int(want[bad[0]]))


@pytest.mark.parametrize("seed", range(4))
def test_random_dictionaries_pair_filter(monkeypatch, seed):
    """The same property check for dictionaries the pair filter takes (byte alphabets, every pattern three bytes or
    longer, many 3-byte patterns, shared prefixes, overlapping and adjacent matches at both parities)."""
    monkeypatch.delenv("PFAC_B200_FILTER", raising=False)
    rng = np.random.default_rng(3000 + seed)
    took = 0
    for trial in range(8):
        asize = int(rng.choice([17, 40, 256]))
        alpha = rng.choice(256, size=asize, replace=False).astype(np.uint8)
        pats = []
        for _ in range(int(rng.integers(1, 80))):
            ln = int(rng.integers(3, 13))
            body = alpha[rng.integers(0, asize, size=ln)].tobytes()
            if pats and rng.random() < 0.4:
                body = (pats[int(rng.integers(0, len(pats)))] + body)[:14]
            pats.append(body)
        pats = list(dict.fromkeys(pats))
        n = 701 + trial                                   # odd and even lengths
        text = alpha[rng.integers(0, asize, size=n)].copy()
        for p in pats[:20]:
            at = int(rng.integers(0, n - len(p) + 1))
            text[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
        tail = pats[int(rng.integers(0, len(pats)))]
        text[n - len(tail):] = np.frombuffer(tail, dtype=np.uint8)
        want = brute_force_match(pats, text)
        tc = TableCompiler(patterns=pats, hot_budget_bytes=64 * 1024)
        info = tc.info()
        assert info["code_bits"] == 8 and info["hashed_filter"] in (1, 3)
        took += info["hashed_filter"] == 3
        L = tc.layout()
        for pad in (0, 0xFF, int(rng.integers(0, 256))):
            got = np.array([emulate_layout_walk(L, len(pats), text, i, pad=pad) for i in range(n)], dtype=np.int32)
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, (seed, trial, pad, int(bad[0]), int(got[bad[0]]), int(want[bad[0]]))
    assert took >= 6      # the pair filter is what these dictionaries get


def test_compiled_table_file_round_trip(tmp_path, golden_dir):
    """PFAC_tableSave / PFAC_tableLoad: the automaton and the layout come back bit for bit (info, every
    layout array, the text dump); truncated, corrupted and foreign files are rejected, a missing
    file is FILE_OPEN_ERROR."""
    for name, budget in (("synth_snort.pat", 64 * 1024), ("synth_dna.pat", 24 * 1024), ("example_pattern", 0)):
        src = os.path.join(golden_dir, name)
        a = TableCompiler(src, hot_budget_bytes=budget)
        f = str(tmp_path / (name + ".pfacb"))
        a.save(f)
        b = TableCompiler(compiled_file=f)
        assert a.info() == b.info()
        la, lb = a.layout(), b.layout()
        assert la.keys() == lb.keys()
        for k in la:
            assert np.array_equal(la[k], lb[k]), k
        a.dump(str(tmp_path / "a.txt"))
        b.dump(str(tmp_path / "b.txt"))
        assert open(str(tmp_path / "a.txt"), "rb").read() == open(str(tmp_path / "b.txt"), "rb").read()
    blob = open(f, "rb").read()
    bad = str(tmp_path / "bad.pfacb")
    for mutate in (lambda x: x[:-1], lambda x: x[:40], lambda x: x + b"\0",
                   lambda x: x[:100] + bytes([x[100] ^ 1]) + x[101:], lambda x: b"NOTPFAC0" + x[8:],
                   lambda x: x[:8] + b"\xff\xff\xff\xff" + x[12:], lambda x: b""):
        open(bad, "wb").write(mutate(blob))
        with pytest.raises(PFACError) as e:
            TableCompiler(compiled_file=bad)
        assert e.value.status == Status.INVALID_PARAMETER
    with pytest.raises(PFACError) as e:
        TableCompiler(compiled_file=str(tmp_path / "missing.pfacb"))
    assert e.value.status == Status.FILE_OPEN_ERROR


def _fnv1a64(b):
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & ((1 << 64) - 1)
    return h


def test_compiled_table_file_content_is_validated(tmp_path, golden_dir):
    """A file whose checksum is right but whose content is not (a hostile or mismatched image) must be
    rejected at load, not crash the dump / a recompile / the kernels: patch fields and recompute the FNV."""
    import struct
    src = os.path.join(golden_dir, "example_pattern")
    a = TableCompiler(src, hot_budget_bytes=0)
    f = str(tmp_path / "ok.pfacb")
    a.save(f)
    blob = bytearray(open(f, "rb").read())
    head, payload = blob[:40], blob[40:]
    (image_len,) = struct.unpack_from("<Q", payload, 0)
    ints = 8 + image_len          # numPatterns, numFinal, initialState, numStates, maxPatternLen, numLeaves
    assert struct.unpack_from("<i", payload, ints)[0] == a.info()["num_patterns"]

    def reject(mut):
        pl = bytearray(payload)
        mut(pl)
        out = bytes(head[:32]) + struct.pack("<Q", _fnv1a64(pl)) + bytes(pl)
        bad = str(tmp_path / "bad.pfacb")
        open(bad, "wb").write(out)
        with pytest.raises(PFACError) as e:
            TableCompiler(compiled_file=bad)
        assert e.value.status == Status.INVALID_PARAMETER

    # the unmodified payload with a recomputed checksum still loads
    out = bytes(head[:32]) + struct.pack("<Q", _fnv1a64(payload)) + bytes(payload)
    good = str(tmp_path / "good.pfacb")
    open(good, "wb").write(out)
    assert TableCompiler(compiled_file=good).info() == a.info()
    reject(lambda pl: struct.pack_into("<i", pl, ints + 4, 50_000_000))     # numFinal
    reject(lambda pl: struct.pack_into("<i", pl, ints + 8, 3))              # initialState != numFinal + 1
    reject(lambda pl: struct.pack_into("<i", pl, ints + 12, 1 << 30))       # numStates
    # an edge target beyond the automaton: the flattened edge list is the last vector of the machine
    # (per-state counts, then {ch, next} pairs); find the pair (ord('A'), next) of the first edge
    edge = bytes(payload).find(struct.pack("<i", ord("A")), ints + 24)
    assert edge > 0
    reject(lambda pl: struct.pack_into("<i", pl, edge + 4, 1 << 28))


def test_saturated_hashed_filter_falls_back_to_exact_stage():
    """Dozens of 1-byte patterns fill the hashed filter (256 words each): the compiler then keeps the
    exact 2-gram stage; results are the brute-force ones either way."""
    rng = np.random.default_rng(5)
    pats = [bytes([c]) for c in range(0x30, 0x30 + 40)] + synth.patterns_snort_like(200, seed=7)
    pats = list(dict.fromkeys(pats))
    tc = TableCompiler(patterns=pats, hot_budget_bytes=64 * 1024)
    assert tc.info()["hashed_filter"] == 0 and tc.info()["has_chk2"] == 1
    L = tc.layout()
    text = rng.integers(0, 256, size=4000, dtype=np.uint8)
    want = brute_force_match(pats, text)
    got = np.array([emulate_layout_walk(L, len(pats), text, i) for i in range(text.size)], dtype=np.int32)
    assert np.array_equal(got, want) and (want > 0).mean() > 0.1


def test_parser_fuzz_against_the_oracle(tmp_path):
    """Random pattern-file images over {a, b, c, CR, NUL, 0xFF, LF}: blank lines, missing trailing
    newline, duplicates, prefixes.  Whenever the image is valid for the reference (no pattern after
    a blank line) the dump -- state numbering, edge order, pattern IDs and lengths -- and the
    matches equal the oracle's (and the brute-force semantics when no pattern occurs twice);
    otherwise INVALID_PARAMETER (the reference aborts on an assert)."""
    rng = np.random.default_rng(2024)
    sym = np.frombuffer(b"abc\r\x00\xff\n\n", dtype=np.uint8)        # LF twice: plenty of short lines
    checked = rejected = 0
    for trial in range(300):
        image = sym[rng.integers(0, sym.size, size=int(rng.integers(0, 40)))].tobytes()
        lines = image.split(b"\n")[:-1]                              # complete lines only
        first_blank = next((i for i, l in enumerate(lines) if not l), None)
        invalid = first_blank is not None and any(lines[first_blank:])
        if invalid:
            with pytest.raises(PFACError) as e:
                TableCompiler(image=image)
            assert e.value.status == Status.INVALID_PARAMETER, image
            rejected += 1
            continue
        tc = TableCompiler(image=image)
        pats = [l for l in lines if l]
        assert tc.info()["num_patterns"] == len(pats), image
        if not pats:
            continue
        o = Oracle(image=image)
        a, b = tmp_path / "a.txt", tmp_path / "b.txt"
        o.dump(str(a))
        tc.dump(str(b))
        assert a.read_bytes() == b.read_bytes(), image
        text = sym[rng.integers(0, sym.size - 2, size=200)].copy()
        L = tc.layout()
        got = np.array([emulate_layout_walk(L, len(pats), text, i) for i in range(text.size)], dtype=np.int32)
        assert np.array_equal(got, o.match(text)), image
        if len(set(pats)) == len(pats):
            assert np.array_equal(got, brute_force_match(pats, text)), image
        # (with duplicates the reference's construction -- first edge for building, last edge for
        # matching -- makes the later copy win AND hides patterns that extend the duplicated one;
        # the oracle and the layout reproduce that, the brute-force restatement does not)
        checked += 1
    assert checked > 100 and rejected > 30


def test_dna_hashed_first_stage_keeps_every_match(golden_dir):
    """2-bit alphabets with an arithmetic symbol code (ACGT: bits 1-2 of the byte) get the hashed 10-mer
    first stage: no false negatives on DNA text with foreign bytes (N, n, newline, 0x00, 0xFF) and with
    patterns shorter than the gram (all their continuations are set), whatever pads the end."""
    from oracle import Oracle
    rng = np.random.default_rng(77)
    pats = synth.patterns_dna(400, short=12)      # lengths 8..24 plus a dozen of length 4..6
    tc = TableCompiler(patterns=pats, hot_budget_bytes=64 * 1024)
    info = tc.info()
    assert info["code_bits"] == 2 and info["code_shift"] == 1 and info["hashed_filter"] == 2
    assert 0 < info["hfilt_bits_set"] < 262144 // 2
    L = tc.layout()
    n = 6000
    text = synth.dna_bytes(5, 0, n)
    text[rng.integers(0, n, size=60)] = np.frombuffer(b"Nn\n\x00\xff-", dtype=np.uint8)[rng.integers(0, 6, size=60)]
    for p in pats[:40]:
        at = int(rng.integers(0, n - len(p) + 1))
        text[at:at + len(p)] = np.frombuffer(p, dtype=np.uint8)
    text[n - 5:] = np.frombuffer(pats[-1][:5].ljust(5, b"A"), dtype=np.uint8)
    want = brute_force_match(pats, text)
    assert (want > 0).sum() > 40
    for pad in (0, ord("A"), ord("T"), 0xFF):
        got = np.array([emulate_layout_walk(L, len(pats), text, i, pad=pad) for i in range(n)], dtype=np.int32)
        assert np.array_equal(got, want), pad
    # lower-case alphabet: same bits; an alphabet whose bytes no two bits tell apart keeps the exact stage
    low = [p.lower() for p in pats[:50]]
    assert TableCompiler(patterns=low, hot_budget_bytes=64 * 1024).info()["code_shift"] == 1
    odd = [bytes([0x10, 0x11, 0x12, 0x13, 0x30][c % 5] for c in p) for p in pats[:50]]   # 5 symbols -> 4-bit codes
    assert TableCompiler(patterns=odd, hot_budget_bytes=64 * 1024).info()["hashed_filter"] == 0
