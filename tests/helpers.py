"""Shared test helpers (CPU side)."""
import numpy as np


def read_patterns(path):
    """Patterns of a reference-grammar file, in ID order (ID = index + 1)."""
    data = open(path, "rb").read()
    parts = data.split(b"\n")
    return [p for p in parts[:-1] if p]  # the unterminated tail is dropped by the parser


def emulate_layout_walk(L, num_final, text, start, n_total=None):
    """Python restatement of the lookup sequence the CUDA kernels run on the compiled device
    layout (pfac_b200/csrc/pfac_kernels.cu): prefilter bit -> root row -> bucketed hash rows
    (hot for depth < hot_depth else cold) -> chain records with byte-wise tail compare.
    Test-only; checks the table compiler without a GPU.  L = TableCompiler.layout()."""
    n_total = len(text) if n_total is None else n_total
    avail = n_total - start
    if avail <= 0:
        return 0
    c0 = int(text[start])
    c1 = int(text[start + 1]) if avail >= 2 else 0
    idx = c0 | (c1 << 8)
    word = int(L["pre2"][idx >> 5])
    b = idx & 31
    if not ((word << b) >> 31) & 1:           # bit 31-(idx&31) of the word
        return 0
    s = int(L["root"][c0])
    assert s >= 0, "prefilter bit set but root row traps"
    best = s if s <= num_final else 0
    if avail < 2:
        return best                            # c1 was padding: only the 1-byte result counts
    rank = int(L["rank2"][idx >> 5]) + (bin(word >> (32 - b)).count("1") if b else 0)
    v = int(L["next2"][rank])                  # the walk after consuming c0, c1
    d = 1
    while True:
        if v == 0xFFFFFFFF:
            break
        if v & 0x80000000:
            off, ln, end, inline4 = (int(x) for x in L["chains"][v & 0x7FFFFFFF])
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
            s = v
            if s <= num_final:
                best = s
            d += 1
        if d >= avail:
            break
        key = ((s << 8) | int(text[start + d])) & 0xFFFFFFFF
        tab = L["hot"] if d < L["hot_depth"] else L["cold"]
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
