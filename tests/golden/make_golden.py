"""Regenerates tests/golden/ from the reference, IN THE BUILD CONTAINER ONLY
(needs /root/reference and oracle/_ref/libpfac_ref.so: `make -C oracle`).

    python tests/golden/make_golden.py

Writes
  example_pattern, example_input, example_pattern2, example_input2
      the reference's four data fixtures (PFAC/test/pattern, PFAC/test/data; 15/10/34/15 bytes)
  <fixture>.dump, <fixture>.dense.json
      PFAC_dumpTransitionTable text and PFAC_CPU result from the reference's own CPU path
  synth_<case>.pat + synth_<case>.npz
      small synthetic pattern sets (pfac_b200/synth.py) with the reference's dense result
      digest, transition-table dump digest and reduced (id, position) lists on the synthetic
      text named in the npz.
The committed outputs are what the GPU box (which has no /root/reference) checks against.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import RefOracle  # noqa: E402
from workloads import synth   # noqa: E402

REF = "/root/reference/PFAC/test"

SYNTH_CASES = {
    # name: (pattern generator, text kind, text seed, n, plant every)
    "c2": (lambda: synth.patterns_c2(300, seed=101), "random", 9001, 200_003, 512),
    "snort": (lambda: synth.patterns_snort_like(800, seed=102), "ascii", 9002, 150_001, 1024),
    "dna": (lambda: synth.patterns_dna(300, seed=103, short=6), "dna", 9003, 120_007, 0),
}


def main():
    for pat, inp in (("example_pattern", "example_input"), ("example_pattern2", "example_input2")):
        shutil.copyfile(os.path.join(REF, "pattern", pat), os.path.join(HERE, pat))
        shutil.copyfile(os.path.join(REF, "data", inp), os.path.join(HERE, inp))
        ref = RefOracle(os.path.join(HERE, pat))
        ref.dump(os.path.join(HERE, pat + ".dump"))
        text = np.fromfile(os.path.join(HERE, inp), dtype=np.uint8)
        dense = ref.match(text, omp=False)
        ids, pos = ref.reduce(dense)
        with open(os.path.join(HERE, pat + ".dense.json"), "w") as f:
            json.dump({"input": inp, "dense": dense.tolist(), "ids": ids.tolist(), "pos": pos.tolist(),
                       "num_states": ref.num_states, "initial_state": ref.initial_state,
                       "max_pattern_len": ref.max_pattern_len}, f)
    for name, (gen, kind, seed, n, every) in SYNTH_CASES.items():
        pats = gen()
        pfile = synth.write_pattern_file(os.path.join(HERE, "synth_%s.pat" % name), pats)
        text = synth.make_text(kind, seed, 0, n, n, pats, every)
        ref = RefOracle(pfile)
        dense = ref.match(text, omp=False)
        assert np.array_equal(dense, ref.match(text, omp=True))
        ids, pos = ref.reduce(dense)
        tmp_dump = os.path.join(HERE, "_tmp.dump")
        ref.dump(tmp_dump)
        dump_sha = hashlib.sha256(open(tmp_dump, "rb").read()).hexdigest()
        os.remove(tmp_dump)
        np.savez_compressed(os.path.join(HERE, "synth_%s.npz" % name),
                            kind=kind, seed=seed, n=n, every=every, ids=ids, pos=pos,
                            text_sha256=hashlib.sha256(text.tobytes()).hexdigest(),
                            dense_sha256=hashlib.sha256(dense.tobytes()).hexdigest(),
                            dump_sha256=dump_sha,
                            num_states=ref.num_states, max_pattern_len=ref.max_pattern_len)
        print(name, "patterns", len(pats), "states", ref.num_states, "matches", ids.size)


if __name__ == "__main__":
    main()
