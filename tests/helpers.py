"""Shared test helpers (CPU side)."""
import numpy as np


def read_patterns(path):
    """Patterns of a reference-grammar file, in ID order (ID = index + 1)."""
    data = open(path, "rb").read()
    parts = data.split(b"\n")
    return [p for p in parts[:-1] if p]  # the unterminated tail is dropped by the parser


def emulate_layout_walk(L, num_final, text, start, n_total=None, pad=0):
    """Python restatement of the lookup sequence the CUDA kernels run on the compiled device
    layout (pfac_b200/csrc/pfac_kernels.cu): symbol codes -> K-gram prefilter bit -> rank ->
    next2/best2 (or, for K-grams with a byte outside the alphabet / cut off by the end of the
    input, the generic path from the root row) -> bucketed hash rows (hot for K <= depth <
    hot_depth, else cold) -> chain records with tail compare.  With the hashed 4-gram first stage
    (L["hfilt"]) a position must pass that filter before anything else; `pad` stands for whatever
    bytes the kernel finds past the end of the input (they must never hide a match).
    Test-only; checks the table compiler without a GPU.  L = TableCompiler.layout()."""
    n_total = len(text) if n_total is None else n_total
    avail = n_total - start
    if avail <= 0:
        return 0
    K, B = L["gram_len"], L["code_bits"]
    lut = L["lut"]
    syms = [int(text[start + i]) if i < avail else 0 for i in range(K)]
    fast = B == 8 or (avail >= K and all(not (int(lut[c]) & 0x80) for c in syms))
    if L["hfilt"].size and B == 2:
        # hashed 10-mer first stage of 2-bit alphabets: arithmetic codes of whatever bytes are there
        sh = L["code_shift"]
        assert sh >= 0 and L["hfilt_k"] == 2
        x = 0
        for i in range(10):
            c = int(text[start + i]) if i < avail else pad
            x |= ((c >> sh) & 3) << (2 * i)
        w = int(L["hfilt"][(((x * 0x9E3779B1) & 0xFFFFFFFF) >> 19) & 8191])
        for m in (0x85EBCA6B, 0xC2B2AE35):
            if not ((w << (((x * m) & 0xFFFFFFFF) >> 27)) >> 31) & 1:
                return 0
    elif L["hfilt"].size and L["hfilt_k"] == 3:
        # pair filter: the positions q and q+1 (q even) are decided together by the three bytes they share
        assert B == 8
        q = start & ~1
        y = sum((int(text[q + 1 + i]) if q + 1 + i < n_total else pad) << (8 * i) for i in range(3))
        h = (y * 0x9E3779B1) & 0xFFFFFF
        w = int(L["hfilt"][(h >> 2) & 8191])
        if not ((w << (h >> 19)) >> 31) & 1:
            return 0
    elif L["hfilt"].size:
        assert B == 8
        x = sum((int(text[start + i]) if i < avail else pad) << (8 * i) for i in range(4))
        if L["hfilt_k"] == 2:   # dense tables: row-indexed, the word is picked by c0 and the low bits of c1 themselves
            w = int(L["hfilt"][x & (L["hfilt"].size - 1)])
        else:
            w = int(L["hfilt"][(((x * 0x9E3779B1) & 0xFFFFFFFF) >> 2) & 8191])   # word picked by a hash of (c0,c1)
        for m in (0x85EBCA6B, 0xC2B2AE35)[:L["hfilt_k"]]:                     # one or two bits by all four bytes
            if not ((w << (((x * m) >> 32) & 31)) >> 31) & 1:
                return 0
    if fast:
        idx = 0
        for i, c in enumerate(syms):
            idx |= (int(lut[c]) & 0x7F if B != 8 else c) << (B * i)
        word = int(L["pre2"][idx >> 5])
        b = idx & 31
        if not ((word << b) >> 31) & 1:           # bit 31-(idx&31) of the word
            return 0
        rank = int(L["rank2"][idx >> 5]) + (bin(word >> (32 - b)).count("1") if b else 0)
        if B == 8:   # one symbol inside K-1: the root row tells whether c0 alone is a pattern
            r0 = int(L["root"][syms[0]])
            best = r0 if 0 <= r0 <= num_final else 0
        else:
            best = int(L["best2"][rank]) if L["best2"].size else 0
        if avail < K:
            return best                            # b == 8 only: the second symbol was padding
        if L["chk2"].size and avail > K:           # second stage: can the next byte lead anywhere?
            if not (int(L["chk2"][rank]) >> (int(text[start + K]) & 15)) & 1:
                return 0                           # (entries with a result of their own are all ones)
        v = int(L["next2"][rank])                  # the walk after its K-th symbol
        d = K - 1
    else:
        s0 = int(L["root"][syms[0]])
        if s0 < 0:
            return 0
        best, v, d = 0, s0, 0
    while True:
        if v == 0xFFFFFFFF:
            break
        if v & 0x80000000:
            off, ln, end, inline4 = (int(x) for x in L["chains"][v & 0x7FFFFF])
            assert ((v >> 23) & 0xFF) == (inline4 & 0xFF), "chain reference must carry the first tail byte"
            if d + 1 >= avail or int(text[start + d + 1]) != ((v >> 23) & 0xFF):
                break  # the kernels stop here without touching the chain record
            if d + 1 + ln > avail:
                break  # cut off by the end of the input: nothing more can be reported
            tail = bytes(L["tails"][off:off + ln])
            assert inline4 == int.from_bytes(tail[:4].ljust(4, b"\0"), "little")
            if bytes(text[start + d + 1:start + d + 1 + ln]) != tail:
                break
            s = end & 0x7FFFFFFF
            d += 1 + ln
            if s <= num_final:
                best = s
            if end & 0x80000000:
                break
        else:
            s = v & 0x3FFFFFFF
            if s <= num_final:
                best = s
            d += 1
            if v & 0x40000000:
                break  # plain leaf: no out-edges
        if d >= avail:
            break
        key = ((s << 8) | int(text[start + d])) & 0xFFFFFFFF
        tab = L["hot"] if K <= d < L["hot_depth"] else L["cold"]
        v = probe(tab, L["mul"], key)
        if v < 0:
            break
    return best


def probe(tab, mul, key):
    nb = tab.shape[0]
    if nb == 0:
        return -1
    b = (((key * mul) & 0xFFFFFFFF) * nb) >> 32
    for _ in range(nb + 1):
        k0, v0, k1, v1 = (int(x) for x in tab[b])
        if k0 == key:
            return v0
        if k1 == key:
            return v1
        if k1 == 0xFFFFFFFF:
            return -1  # trap
        b = 0 if b + 1 == nb else b + 1
    raise AssertionError("probe did not terminate: table has no empty slot")


def brute_force_match(patterns, text):
    """Reference semantics restated without the file grammar (needed for patterns that contain 0x0A):
    result[i] = ID (index + 1) of the longest pattern that is a prefix of text[i:], else 0; of two
    identical patterns the later ID wins."""
    text = bytes(text)
    by_len = {}
    for i, p in enumerate(patterns):
        by_len.setdefault(len(p), {})[bytes(p)] = i + 1
    lens = sorted(by_len, reverse=True)
    out = np.zeros(len(text), dtype=np.int32)
    for i in range(len(text)):
        for L in lens:
            if i + L <= len(text):
                pid = by_len[L].get(text[i:i + L])
                if pid:
                    out[i] = pid
                    break
    return out


class CheckerOracle:
    """The checker the GPU tests compare with: the reference's own CPU matcher (oracle/_ref, compiled
    from /root/reference by oracle/Makefile; it travels to the GPU box as a built .so) when it is
    there, else the plain-C restatement (oracle/pfac_oracle.c) -- the two are pinned against each
    other in tests/test_oracle.py.  Same surface as oracle.Oracle."""

    def __init__(self, pattern_file):
        import oracle
        self.kind = "reference" if oracle.ref_available() else "port"
        port = oracle.Oracle(pattern_file)   # raises on files the reference would assert-abort on
        self._o = oracle.RefOracle(pattern_file) if oracle.ref_available() else port
        for k in ("num_patterns", "num_states", "initial_state", "max_pattern_len"):
            setattr(self, k, getattr(self._o, k))

    @staticmethod
    def set_threads(n):
        import oracle
        (oracle.RefOracle if oracle.ref_available() else oracle.Oracle).set_threads(n)

    def match(self, text, omp=True):
        return self._o.match(text, omp=omp)

    def match_shard(self, text, n_owned):
        """Results for positions [0, n_owned); walks may read all of `text` (owned + tail halo)."""
        if self.kind == "port":
            return self._o.match_shard(text, n_owned)
        return self._o.match(text, omp=True)[:n_owned]

    def reduce(self, dense):
        import oracle
        return oracle.reduce_dense(dense)

    def dump(self, path):
        return self._o.dump(path)
