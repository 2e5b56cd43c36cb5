"""Synthetic workloads for the BASELINE.json configs.

Deterministic and counter-based: text byte i depends only on (seed, i), so any byte range can
be regenerated (shards on different ranks, bounded CPU samples) without holding the whole
input.  Patterns are planted because uniform random text over 256 symbols matches nothing
(SURVEY.md section 0.5).  numpy only; used by tests/, bench.py and __graft_entry__.smoke().
"""
import numpy as np

MASK64 = (1 << 64) - 1
SEED_BASE = 0x5046414300000000  # ASCII "PFAC" << 32, + config number


def _splitmix64(x):
    """Vectorised splitmix64 finaliser on uint64 arrays."""
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _mix_scalar(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK64
    return x ^ (x >> 31)


def random_bytes(seed, start, n):
    """text[start:start+n] of the uniform 256-symbol stream: 8 bytes per splitmix64 word."""
    if n <= 0:
        return np.zeros(0, dtype=np.uint8)
    w0, w1 = start // 8, (start + n + 7) // 8
    with np.errstate(over="ignore"):
        words = _splitmix64(np.arange(w0, w1, dtype=np.uint64) + np.uint64(seed & MASK64))
    b = words.view(np.uint8)  # little-endian bytes of each word
    off = start - w0 * 8
    return b[off:off + n].copy()


def ascii_weighted_bytes(seed, start, n, printable_frac=0.5):
    """Half the bytes forced into printable ASCII (0x20..0x7E), the rest uniform (config 3)."""
    raw = random_bytes(seed, start, n)
    sel = random_bytes(seed ^ 0xA5A5A5A5, start, n)
    printable = (raw.astype(np.uint16) * 95 >> 8).astype(np.uint8) + np.uint8(0x20)
    return np.where(sel < int(256 * printable_frac), printable, raw).astype(np.uint8)


def dna_bytes(seed, start, n):
    """Uniform ACGT text (config 4)."""
    raw = random_bytes(seed, start, n)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[raw & 3]


# ---------------------------------------------------------------------------------- patterns

def _dedupe_keep(pats, n):
    seen, out = set(), []
    for p in pats:
        if p and p not in seen and b"\n" not in p:
            seen.add(p)
            out.append(p)
            if len(out) == n:
                break
    return out


def patterns_c2(n=1000, seed=SEED_BASE + 2, min_len=4, max_len=32, prefix_pairs=50):
    """n distinct patterns, length uniform [min_len,max_len], bytes uniform over the 255 values
    != 0x0A, including `prefix_pairs` (P, P+suffix) pairs to exercise longest-match."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pats = []
    while len(pats) < n * 2:
        L = int(rng.integers(min_len, max_len + 1))
        b = rng.integers(0, 255, size=L, dtype=np.int64)
        b = np.where(b >= 10, b + 1, b).astype(np.uint8)  # skip 0x0A
        pats.append(b.tobytes())
    pats = _dedupe_keep(pats, n - prefix_pairs)
    ext = []
    for i in range(prefix_pairs):
        base = pats[int(rng.integers(0, len(pats)))]
        room = max_len - len(base)
        if room <= 0:
            base = base[:max_len - 4]
            room = 4
        k = int(rng.integers(1, room + 1))
        b = rng.integers(0, 255, size=k, dtype=np.int64)
        b = np.where(b >= 10, b + 1, b).astype(np.uint8)
        ext.append(base + b.tobytes())
    out = _dedupe_keep(pats + ext, n)
    order = rng.permutation(len(out))
    return [out[i] for i in order]


def patterns_snort_like(n=20000, seed=SEED_BASE + 3):
    """Snort-like law: lengths 1..243 heavy at 4..20 (mean ~21), 70 % printable / 30 % binary
    bytes, 30 % of patterns extend an earlier one; includes a few 1-byte patterns."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pats = []
    stems = [b"GET /", b"POST /", b"HTTP/1.", b"User-Agent: ", b"Content-", b"/cgi-bin/", b".php?",
             b"cmd.exe", b"SELECT ", b"\x90\x90\x90\x90", b"|00 00|", b"admin", b"passwd", b"/etc/"]
    pats += [b"\x7f", b"~"]  # 1-byte patterns: a whole root symbol always matches
    target = n * 2
    while len(pats) < target:
        r = rng.random()
        if r < 0.80:
            L = int(rng.integers(4, 21))
        elif r < 0.97:
            L = int(rng.integers(21, 65))
        else:
            L = int(rng.integers(65, 244))
        if rng.random() < 0.7:
            b = rng.integers(0x20, 0x7F, size=L, dtype=np.int64).astype(np.uint8)
        else:
            b = rng.integers(0, 255, size=L, dtype=np.int64)
            b = np.where(b >= 10, b + 1, b).astype(np.uint8)
        body = b.tobytes()
        u = rng.random()
        if u < 0.15 and pats:
            base = pats[int(rng.integers(0, len(pats)))]
            body = (base + body)[:243]
        elif u < 0.30:
            body = (stems[int(rng.integers(0, len(stems)))] + body)[:243]
        pats.append(body)
    out = _dedupe_keep(pats, n)
    order = rng.permutation(len(out))
    return [out[i] for i in order]


def patterns_dna(n=5000, seed=SEED_BASE + 4, min_len=8, max_len=24, short=0):
    """n distinct ACGT patterns, length uniform [min_len,max_len]; `short` extra patterns of
    length 4..6 give the high-density compaction variant."""
    rng = np.random.Generator(np.random.PCG64(seed))
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    pats = []
    while len(pats) < n * 2:
        L = int(rng.integers(min_len, max_len + 1))
        pats.append(alpha[rng.integers(0, 4, size=L)].tobytes())
    out = _dedupe_keep(pats, n)
    extra = []
    while len(extra) < short * 2:
        L = int(rng.integers(4, 7))
        extra.append(alpha[rng.integers(0, 4, size=L)].tobytes())
    out = _dedupe_keep(out + extra, n + short)
    return out


def pattern_file_image(patterns):
    """The file the reference parser expects: one pattern per line, file ends with '\\n'."""
    return b"".join(p + b"\n" for p in patterns)


def write_pattern_file(path, patterns):
    with open(path, "wb") as f:
        f.write(pattern_file_image(patterns))
    return path


# ---------------------------------------------------------------------------------- planting

def plant(text, start, total_len, patterns, seed, every=4096, boundary=1 << 20):
    """Overwrite `text` (= stream bytes [start, start+len(text))) with planted patterns:
      * one pattern per `every`-byte block at a hash-derived offset,
      * one pattern straddling every `boundary` multiple (tile / chunk / shard edges),
      * patterns cut off by the end of the stream (must NOT be reported),
    all as pure functions of the absolute position, so shards agree with the whole stream."""
    n = len(text)
    P = len(patterns)
    if P == 0 or n == 0:
        return text
    end = start + n

    def put(abs_pos, pat):
        a = max(abs_pos, start)
        b = min(abs_pos + len(pat), end, total_len)
        if b > a:
            text[a - start:b - start] = np.frombuffer(pat, dtype=np.uint8)[a - abs_pos:b - abs_pos]

    maxlen = max(len(p) for p in patterns)
    first_blk = max(0, (start - maxlen) // every)
    last_blk = (min(end, total_len) + every - 1) // every
    for blk in range(first_blk, last_blk):
        h = _mix_scalar((seed ^ 0x1234567) + blk)
        pat = patterns[h % P]
        span = every - len(pat)
        if span <= 0:
            continue
        put(blk * every + (h >> 20) % span, pat)
    first_b = max(1, start // boundary)
    for k in range(first_b, (min(end, total_len) + maxlen) // boundary + 1):
        h = _mix_scalar((seed ^ 0x7654321) + k)
        pat = patterns[h % P]
        if len(pat) < 2:
            continue
        cut = 1 + (h >> 24) % (len(pat) - 1)
        put(k * boundary - cut, pat)
    # truncated at the end of the stream: the last bytes are a proper prefix of a long pattern
    h = _mix_scalar(seed ^ 0xE0D)
    longest = max(patterns, key=len)
    if len(longest) >= 2 and total_len >= len(longest):
        put(total_len - (len(longest) - 1), longest[:-1])
    return text


def make_text(kind, seed, start, n, total_len, patterns=None, every=4096):
    """kind: 'random' (C2), 'ascii' (C3/C5), 'dna' (C4)."""
    if kind == "random":
        t = random_bytes(seed, start, n)
    elif kind == "ascii":
        t = ascii_weighted_bytes(seed, start, n)
    elif kind == "dna":
        t = dna_bytes(seed, start, n)
    else:
        raise ValueError(kind)
    if patterns and every:
        plant(t, start, total_len, patterns, seed, every=every)
    return t
