"""Pins the oracle (oracle/pfac_oracle.c) before anything trusts it:
  * doc goldens G1-G4,
  * committed fixtures generated from the reference's own CPU path (tests/golden/make_golden.py),
  * the reference itself (oracle/_ref) on randomized planted cases, when it is built
    (in the build container; /root/reference is absent on the GPU box).
CPU only.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from oracle import Oracle, RefOracle, reduce_dense
from workloads import synth

G1_DENSE = [1, 3, 4, 0, 4, 0, 2, 0, 0, 0]          # reference README.md:114-120 (+ trailing '\n')
G2_POS, G2_IDS = [0, 1, 2, 4, 6], [1, 3, 4, 4, 2]  # user guide r1.2 p.29
G4_DENSE = [4, 3, 0, 4, 5, 0, 0, 1, 7, 9, 1, 8, 9, 1, 0]


def test_g1_g2_readme_example(golden_dir):
    o = Oracle(os.path.join(golden_dir, "example_pattern"))
    text = np.fromfile(os.path.join(golden_dir, "example_input"), dtype=np.uint8)
    assert text.tobytes() == b"ABEDEDABG\n"
    for omp in (False, True):
        assert o.match(text, omp=omp).tolist() == G1_DENSE
    ids, pos = o.reduce(o.match(text))
    assert ids.tolist() == G2_IDS and pos.tolist() == G2_POS
    assert (o.num_patterns, o.num_states, o.initial_state, o.max_pattern_len) == (4, 11, 5, 4)


def test_g3_transition_table_dump(golden_dir, tmp_path):
    """User guide r1.2 p.21: 11 states, initial 5, the nine edges, the output table."""
    o = Oracle(os.path.join(golden_dir, "example_pattern"))
    out = tmp_path / "dump.txt"
    o.dump(str(out))
    text = out.read_text()
    assert text == open(os.path.join(golden_dir, "example_pattern.dump")).read()
    assert "# Transition table: number of states = 11, initial state = 5\n" in text
    for edge in ["(   1,   G) -> 2 ", "(   5,   A) -> 6 ", "(   5,   B) -> 7 ", "(   5,   E) -> 10 ",
                 "(   6,   B) -> 1 ", "(   7,   E) -> 8 ", "(   8,   D) -> 9 ", "(   9,   E) -> 3 ",
                 "(  10,   D) -> 4 "]:
        assert edge + "\n" in text
    assert '    1     1     2    "AB"\n' in text and '    3     3     4    "BEDE"\n' in text


def test_g4_second_machine(golden_dir, tmp_path):
    """PFAC_hash_draft.pdf p.1: F=10, 14 states (+ unused 0), initial 11, (11,h)->2 (2,e)->3 (2,i)->12."""
    o = Oracle(os.path.join(golden_dir, "example_pattern2"))
    assert (o.num_patterns, o.num_states, o.initial_state) == (10, 15, 11)
    T = o.dense_table()
    assert T[11, ord("h")] == 2 and T[2, ord("e")] == 3 and T[2, ord("i")] == 12
    leaves = [s for s in range(1, 11) if (T[s] == -1).all()]
    assert leaves == [4, 5, 7, 8, 9]
    text = np.fromfile(os.path.join(golden_dir, "example_input2"), dtype=np.uint8)
    assert o.match(text).tolist() == G4_DENSE
    out = tmp_path / "dump2.txt"
    o.dump(str(out))
    assert out.read_text() == open(os.path.join(golden_dir, "example_pattern2.dump")).read()


@pytest.mark.parametrize("fixture", ["example_pattern", "example_pattern2"])
def test_reference_generated_json(golden_dir, fixture):
    g = json.load(open(os.path.join(golden_dir, fixture + ".dense.json")))
    o = Oracle(os.path.join(golden_dir, fixture))
    text = np.fromfile(os.path.join(golden_dir, g["input"]), dtype=np.uint8)
    dense = o.match(text)
    assert dense.tolist() == g["dense"]
    ids, pos = o.reduce(dense)
    assert ids.tolist() == g["ids"] and pos.tolist() == g["pos"]
    assert (o.num_states, o.initial_state, o.max_pattern_len) == (
        g["num_states"], g["initial_state"], g["max_pattern_len"])


@pytest.mark.parametrize("case", ["c2", "snort", "dna"])
def test_reference_generated_synthetic(golden_dir, tmp_path, case):
    """Committed digests/lists came from the reference's PFAC_CPU / PFAC_dumpTransitionTable."""
    g = np.load(os.path.join(golden_dir, "synth_%s.npz" % case))
    pfile = os.path.join(golden_dir, "synth_%s.pat" % case)
    n, seed, every, kind = int(g["n"]), int(g["seed"]), int(g["every"]), str(g["kind"])
    from tests.helpers import read_patterns
    pats = read_patterns(pfile)
    text = synth.make_text(kind, seed, 0, n, n, pats, every)
    assert hashlib.sha256(text.tobytes()).hexdigest() == str(g["text_sha256"]), "generator drifted"
    o = Oracle(pfile)
    assert o.num_states == int(g["num_states"]) and o.max_pattern_len == int(g["max_pattern_len"])
    dense = o.match(text)
    assert hashlib.sha256(dense.tobytes()).hexdigest() == str(g["dense_sha256"])
    ids, pos = o.reduce(dense)
    assert np.array_equal(ids, g["ids"]) and np.array_equal(pos, g["pos"])
    i2, p2 = reduce_dense(dense)
    assert np.array_equal(i2, ids) and np.array_equal(p2, pos)
    out = tmp_path / "d.txt"
    o.dump(str(out))
    assert hashlib.sha256(out.read_bytes()).hexdigest() == str(g["dump_sha256"])


def test_parser_quirks(tmp_path):
    """reference PFAC_reorder_Table.cpp:176-193: last line without newline is dropped, '\\r' and
    0x00/0x80-0xFF are pattern bytes, a blank line before a pattern is an error (the reference
    asserts), trailing blank lines are harmless."""
    o = Oracle(image=b"AB\nCD")
    assert o.num_patterns == 1
    o = Oracle(image=b"A\r\n\x00\xff\n")
    assert o.num_patterns == 2 and o.match(b"A\r\x00\xff").tolist() == [1, 0, 2, 0]
    assert Oracle(image=b"AB\n\n\n").num_patterns == 1
    with pytest.raises(ValueError):
        Oracle(image=b"AB\n\nCD\n")
    with pytest.raises(ValueError):
        Oracle(image=b"\nAB\n")


def test_signed_char_order_and_numbering():
    """pattern_cmp_functor compares plain char: 0x80..0xFF sort before 0x00..0x7F
    (PFAC_reorder_Table.cpp:56-60), which fixes the internal state numbering."""
    o = Oracle(image=b"zAq\n\x01Aq\n\xffAq\n")
    T = o.dense_table()
    init = o.initial_state  # 4; internal states from 5 in sorted first-visit order
    assert T[init, 0xFF] == 5 and T[init, 0x01] == 7 and T[init, ord("z")] == 9


def test_longest_match_prefix_and_truncation():
    o = Oracle(image=b"AB\nABCD\nB\n")
    assert o.match(b"ABCD").tolist() == [2, 3, 0, 0]
    assert o.match(b"ABC").tolist() == [1, 3, 0]      # ABCD cut off by the end: not reported
    assert o.match(b"").tolist() == []
    assert o.match_shard(np.frombuffer(b"ABCDAB", dtype=np.uint8), 2).tolist() == [2, 3]


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("seed", range(6))
def test_against_reference_build_random(tmp_path, seed):
    rng = np.random.default_rng(seed)
    kind = ["random", "ascii", "dna"][seed % 3]
    if kind == "dna":
        pats = synth.patterns_dna(int(rng.integers(5, 200)), seed=seed, min_len=3, max_len=12, short=3)
    elif kind == "ascii":
        pats = synth.patterns_snort_like(int(rng.integers(5, 400)), seed=seed)
    else:
        pats = synth.patterns_c2(int(rng.integers(60, 300)), seed=seed, min_len=1, max_len=9, prefix_pairs=20)
    pfile = synth.write_pattern_file(str(tmp_path / "p.txt"), pats)
    n = int(rng.integers(1, 60000))
    text = synth.make_text(kind, seed, 0, n, n, pats, 256)
    o, r = Oracle(pfile), RefOracle(pfile)
    assert (o.num_states, o.initial_state, o.max_pattern_len) == (r.num_states, r.initial_state, r.max_pattern_len)
    assert np.array_equal(o.dense_table(), r.dense_table())
    want = r.match(text, omp=False)
    assert np.array_equal(want, r.match(text, omp=True))
    assert np.array_equal(o.match(text, omp=False), want)
    assert np.array_equal(o.match(text, omp=True), want)
    o.dump(str(tmp_path / "o.txt"))
    r.dump(str(tmp_path / "r.txt"))
    assert (tmp_path / "o.txt").read_bytes() == (tmp_path / "r.txt").read_bytes()


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
def test_parser_fuzz_against_reference_build(tmp_path):
    """Random small pattern files over {a, b, c, CR, NUL, 0xFF, LF} (missing trailing newline, trailing
    blank lines, prefixes; no pattern after a blank line -- the reference aborts there -- and no
    duplicates -- its std::sort order is unspecified for them): table, dump and matches of the
    restatement equal the reference build's."""
    rng = np.random.default_rng(77)
    sym = np.frombuffer(b"abc\r\x00\xff\n\n", dtype=np.uint8)
    done = 0
    for trial in range(400):
        image = sym[rng.integers(0, sym.size, size=int(rng.integers(1, 40)))].tobytes()
        lines = image.split(b"\n")[:-1]
        first_blank = next((i for i, l in enumerate(lines) if not l), None)
        pats = [l for l in lines if l]
        if (first_blank is not None and any(lines[first_blank:])) or not pats or len(set(pats)) != len(pats):
            continue
        pfile = tmp_path / ("f%d.txt" % trial)
        pfile.write_bytes(image)
        o, r = Oracle(str(pfile)), RefOracle(str(pfile))
        assert (o.num_patterns, o.num_states, o.max_pattern_len) == (r.num_patterns, r.num_states, r.max_pattern_len)
        assert np.array_equal(o.dense_table(), r.dense_table()), image
        text = sym[rng.integers(0, sym.size - 2, size=300)].copy()
        assert np.array_equal(o.match(text), r.match(text, omp=False)), image
        o.dump(str(tmp_path / "o.txt"))
        r.dump(str(tmp_path / "r.txt"))
        assert (tmp_path / "o.txt").read_bytes() == (tmp_path / "r.txt").read_bytes(), image
        done += 1
    assert done > 100
