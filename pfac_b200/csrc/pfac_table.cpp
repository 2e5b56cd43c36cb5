// pfac_table.cpp -- table compiler (host only).  See pfac_table.h.
#include "pfac_table.h"

#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "PFAC.h"

namespace pfac {

namespace {

struct Pat {
    size_t off;
    int id;
};

// Order of reference pattern_cmp_functor (PFAC_reorder_Table.cpp:37-72): '\n'-terminated,
// plain char (signed on x86-64: 0x80..0xFF sort before 0x00..0x7F), proper prefix first.
// The reference returns true for equal strings (not a strict weak order); here equal strings
// compare "not less" and std::stable_sort keeps file order, so of two identical patterns the
// later ID wins in the matching table.
struct PatLess {
    const char* base;
    const int* len;  // by id
    bool operator()(const Pat& a, const Pat& b) const {
        const signed char* s = reinterpret_cast<const signed char*>(base + a.off);
        const signed char* t = reinterpret_cast<const signed char*>(base + b.off);
        const int ls = len[a.id], lt = len[b.id];
        const int n = ls < lt ? ls : lt;
        for (int i = 0; i < n; i++) {
            if (s[i] < t[i]) return true;
            if (s[i] > t[i]) return false;
        }
        return ls < lt;  // proper prefix first; equal strings: not less (stable sort keeps file order)
    }
};

inline uint32_t edgeKey(int state, int ch) { return (uint32_t(state) << 8) | uint32_t(ch); }

inline uint32_t homeBucket(uint32_t key, uint32_t mul, uint32_t nb) {
    return uint32_t((uint64_t(key * mul) * uint64_t(nb)) >> 32);
}

struct FlatEdge {
    uint32_t key;
    int next;
    int depth;  // depth of the source state
};

// Fill a bucketed table; returns the longest probe sequence (in buckets) over all keys.
int fillBuckets(const std::vector<FlatEdge>& edges, size_t begin, size_t end, uint32_t nb,
                uint32_t mul, std::vector<uint32_t>& out) {
    out.assign(size_t(nb) * 4, kEmptyKey);
    int maxProbe = 0;
    for (size_t i = begin; i < end; i++) {
        uint32_t b = homeBucket(edges[i].key, mul, nb);
        int probe = 1;
        for (;;) {
            uint32_t* e = &out[size_t(b) * 4];
            if (e[0] == kEmptyKey) { e[0] = edges[i].key; e[1] = uint32_t(edges[i].next); break; }
            if (e[2] == kEmptyKey) { e[2] = edges[i].key; e[3] = uint32_t(edges[i].next); break; }
            b = (b + 1 == nb) ? 0 : b + 1;
            probe++;
        }
        maxProbe = std::max(maxProbe, probe);
    }
    // a miss stops at the first bucket whose slot 1 is empty: its worst case is the longest
    // run of full buckets + 1
    int run = 0, longest = 0;
    for (uint32_t pass = 0; pass < 2 * nb; pass++) {
        uint32_t b = pass % nb;
        if (out[size_t(b) * 4 + 2] != kEmptyKey) { run++; longest = std::max(longest, run); if (run >= int(nb)) break; }
        else run = 0;
    }
    return std::max(maxProbe, longest + 1);
}

int finishMachine(Machine& m, std::vector<Pat>& pats, const std::vector<int>& lens);

}  // namespace

int buildMachine(const char* image, size_t size, Machine& m) {
    m = Machine();
    if (image == nullptr && size != 0) return PFAC_STATUS_INVALID_PARAMETER;
    if (size >= size_t(kMaxStates)) return PFAC_STATUS_INTERNAL_ERROR;  // states <= size+2
    m.image.assign(image ? image : "", size);
    m.image.push_back('\n');
    const char* buf = m.image.data();

    // Line split, reference PFAC_reorder_Table.cpp:176-193: only 0x0A separates; the '\n' at
    // i closes a pattern iff i>0 && buf[i-1] != '\n'; an unterminated last line is dropped.
    std::vector<Pat> pats;
    std::vector<int> lens;
    size_t lineStart = 0;
    bool pendingBlank = false;
    for (size_t i = 0; i < size; i++) {
        if (buf[i] != '\n') continue;
        if (i > 0 && buf[i - 1] != '\n') {
            // the reference would hand the trie builder a pointer to the blank line here and
            // abort on assert('\n' != ch) (:291); report it instead
            if (pendingBlank) { m = Machine(); return PFAC_STATUS_INVALID_PARAMETER; }
            pats.push_back(Pat{lineStart, int(pats.size()) + 1});
            lens.push_back(int(i - lineStart));
        } else {
            pendingBlank = true;
        }
        lineStart = i + 1;
    }
    return finishMachine(m, pats, lens);
}

namespace {

// patterns given as (offset into m.image, id) + lengths: order, trie, numbering
int finishMachine(Machine& m, std::vector<Pat>& pats, const std::vector<int>& lens) {
    const char* buf = m.image.data();
    const size_t size = m.image.size();
    const int k = int(pats.size());
    m.numPatterns = m.numFinal = k;
    m.initialState = k + 1;
    m.lenById.assign(size_t(k) + 1, 0);
    m.offById.assign(size_t(k) + 1, 0);
    for (int i = 0; i < k; i++) {
        m.lenById[size_t(i) + 1] = lens[size_t(i)];
        m.offById[size_t(i) + 1] = pats[size_t(i)].off;
        m.maxPatternLen = std::max(m.maxPatternLen, lens[size_t(i)]);
    }
    std::stable_sort(pats.begin(), pats.end(), PatLess{buf, m.lenById.data()});
    m.sortedOff.resize(size_t(k));
    m.sortedId.resize(size_t(k));
    for (int i = 0; i < k; i++) {
        m.sortedOff[size_t(i)] = pats[size_t(i)].off;
        m.sortedId[size_t(i)] = pats[size_t(i)].id;
    }

    // Trie with the reference numbering (PFAC_reorder_Table.cpp:279-321): state 0 unused,
    // finals 1..k = pattern ids, initial k+1, internal states from k+2 in first-visit order.
    // The last byte of a pattern ALWAYS appends (state,ch)->id; traversal follows the FIRST
    // edge for a byte (lookup(), :234-244).
    m.rows.assign(size_t(k) + 2, std::vector<Edge>());
    std::unordered_map<uint32_t, int> firstEdge;
    firstEdge.reserve(size * 2 + 16);
    int stateNum = m.initialState + 1;
    for (int p = 0; p < k; p++) {
        const unsigned char* pos = reinterpret_cast<const unsigned char*>(buf + m.sortedOff[size_t(p)]);
        const int id = m.sortedId[size_t(p)];
        const int len = m.lenById[size_t(id)];
        int state = m.initialState;
        for (int off = 0; off < len; off++) {
            const int ch = pos[off];
            const uint32_t key = edgeKey(state, ch);
            if (off == len - 1) {
                m.rows[size_t(state)].push_back(Edge{ch, id});
                firstEdge.emplace(key, id);  // keeps an earlier edge if one exists
            } else {
                auto it = firstEdge.find(key);
                if (it == firstEdge.end()) {
                    m.rows[size_t(state)].push_back(Edge{ch, stateNum});
                    firstEdge.emplace(key, stateNum);
                    m.rows.emplace_back();
                    state = stateNum++;
                } else {
                    state = it->second;
                }
            }
        }
    }
    m.numStates = stateNum;
    if (m.numStates > kMaxStates) { m = Machine(); return PFAC_STATUS_INTERNAL_ERROR; }
    m.rows.resize(size_t(m.numStates));
    for (int i = 1; i <= k; i++)
        if (m.rows[size_t(i)].empty()) m.numLeaves++;  // reference PFAC.cpp:716-722
    return PFAC_STATUS_SUCCESS;
}

}  // namespace

// Patterns from arrays: any byte may occur (the file grammar cannot express 0x0A), lengths are
// explicit, IDs are 1-based array order.  Same order / numbering rules as the file form.
int buildMachineFromArrays(const char* const* ptrs, const size_t* lens, size_t n, Machine& m) {
    m = Machine();
    if ((!ptrs || !lens) && n != 0) return PFAC_STATUS_INVALID_PARAMETER;
    size_t total = 0;
    for (size_t i = 0; i < n; i++) {
        if (lens[i] == 0 || !ptrs[i]) return PFAC_STATUS_INVALID_PARAMETER;  // empty patterns match nothing
        total += lens[i] + 1;
    }
    if (total >= size_t(kMaxStates)) return PFAC_STATUS_INTERNAL_ERROR;
    std::vector<Pat> pats;
    std::vector<int> plens;
    m.image.reserve(total + 1);
    for (size_t i = 0; i < n; i++) {
        pats.push_back(Pat{m.image.size(), int(i) + 1});
        plens.push_back(int(lens[i]));
        m.image.append(ptrs[i], lens[i]);
        m.image.push_back('\n');  // storage separator only; lengths are authoritative
    }
    return finishMachine(m, pats, plens);
}

void dumpMachine(const Machine& m, FILE* fp) {
    fprintf(fp, "# Transition table: number of states = %d, initial state = %d\n", m.numStates,
            m.initialState);
    fprintf(fp, "# (current state, input character) -> next state \n");
    for (int s = 0; s < m.numStates; s++) {
        for (const Edge& e : m.rows[size_t(s)]) {
            if (e.ch >= 32 && e.ch <= 126) fprintf(fp, "(%4d,%4c) -> %d \n", s, e.ch, e.next);
            else fprintf(fp, "(%4d,%4.2x) -> %d \n", s, e.ch, e.next);
        }
    }
    fprintf(fp, "# Output table: number of final states = %d\n", m.numFinal);
    fprintf(fp, "# [final state] [matched pattern ID] [pattern length] [pattern(string literal)] \n");
    for (int s = 1; s <= m.numFinal; s++) {
        const int len = m.lenById[size_t(s)];
        fprintf(fp, "%5d %5d %5d    ", s, s, len);
        const unsigned char* p = reinterpret_cast<const unsigned char*>(m.image.data() + m.offById[size_t(s)]);
        fputc('"', fp);
        for (int i = 0; i < len; i++) {
            if (p[i] >= 32 && p[i] <= 126) fputc(p[i], fp);
            else fprintf(fp, "%2.2x", p[i]);
        }
        fputc('"', fp);
        fputc('\n', fp);
    }
}

void compileLayout(const Machine& m, size_t hotBudgetBytes, DeviceLayout& L, int filterPolicy) {
    L = DeviceLayout();
    for (int c = 0; c < kCharSet; c++) L.root[c] = -1;
    L.pre2.assign(65536 / 32, 0u);

    // ---- matching automaton: per (state,ch) the LAST edge in insertion order wins, as in the
    // reference's dense table fill (PFAC.cpp:376-381).  BFS from the initial state gives every
    // reachable state's depth and its (sorted) out-edges.
    const size_t S = size_t(std::max(m.numStates, 1));
    std::vector<int> depth(S, -1);
    std::vector<std::vector<Edge>> out(S);
    std::vector<int> frontier;
    if (m.numStates > m.initialState) {
        depth[size_t(m.initialState)] = 0;
        frontier.push_back(m.initialState);
    }
    int last[kCharSet];
    std::fill(last, last + kCharSet, -1);
    std::vector<int> touched;
    for (size_t qi = 0; qi < frontier.size(); qi++) {
        const int s = frontier[qi];
        const std::vector<Edge>& row = m.rows[size_t(s)];
        touched.clear();
        for (size_t j = 0; j < row.size(); j++) {
            if (last[row[j].ch] < 0) touched.push_back(row[j].ch);
            last[row[j].ch] = int(j);
        }
        std::sort(touched.begin(), touched.end());
        for (int ch : touched) {
            const int nx = row[size_t(last[ch])].next;
            last[ch] = -1;
            if (depth[size_t(nx)] < 0) {
                depth[size_t(nx)] = depth[size_t(s)] + 1;
                frontier.push_back(nx);
            }
            out[size_t(s)].push_back(Edge{ch, nx});
            L.numEdges++;
            L.maxDepth = std::max(L.maxDepth, depth[size_t(s)] + 1);
        }
    }
    auto isFinal = [&](int s) { return s >= 1 && s <= m.numFinal; };

    // ---- root row --------------------------------------------------------------------------------
    if (!frontier.empty()) {
        for (const Edge& e : out[size_t(m.initialState)]) {
            L.root[e.ch] = e.next;
            L.rootFanout++;
        }
    }

    // ---- symbol codes: b bits per symbol, K = 16/b symbols per prefilter index ---------------------
    bool used[kCharSet] = {false};
    int alphabet = 0;
    for (size_t st = 0; st < S; st++)
        for (const Edge& e : out[st])
            if (!used[e.ch]) { used[e.ch] = true; alphabet++; }
    L.codeBits = (alphabet <= 4) ? 2 : (alphabet <= 16) ? 4 : 8;
    L.gramLen = 16 / L.codeBits;
    const int K = L.gramLen, B = L.codeBits;
    int byteOfCode[kCharSet];   // -1: code not in use
    std::fill(byteOfCode, byteOfCode + kCharSet, -1);
    int ncodes = 0;
    // 2-bit alphabets: when two bits of the byte itself tell the symbols apart (ACGT and acgt: bits 1-2),
    // those bits ARE the code, and the kernels code four text bytes at a time with a shift, a mask and a
    // multiply instead of one table lookup per byte (the hashed 10-mer first stage below needs that)
    L.codeShift = -1;
    if (B == 2 && alphabet > 0) {
        for (int sh = 0; sh <= 6 && L.codeShift < 0; sh++) {
            bool seen[4] = {false, false, false, false}, ok = true;
            for (int c = 0; c < kCharSet && ok; c++)
                if (used[c]) { ok = !seen[(c >> sh) & 3]; seen[(c >> sh) & 3] = true; }
            if (ok) L.codeShift = sh;
        }
    }
    for (int c = 0; c < kCharSet; c++) {
        if (B == 8) { L.lut[c] = uint8_t(c); byteOfCode[c] = c; continue; }  // identity
        if (!used[c]) { L.lut[c] = 0x80; continue; }
        if (L.codeShift >= 0) {
            L.lut[c] = uint8_t((c >> L.codeShift) & 3);
            byteOfCode[L.lut[c]] = c;
        } else {
            byteOfCode[ncodes] = c;
            L.lut[c] = uint8_t(ncodes++);
        }
    }
    if (L.codeShift >= 0) ncodes = 4;   // codes may have gaps: byteOfCode = -1 there
    if (B == 8) ncodes = 256;

    // ---- hash edges with chain compression.  A run of >= kMinChain non-final single-child
    // states behind an edge becomes one chain record whose tail bytes are compared directly
    // against the text (independent loads) instead of one dependent lookup per byte.
    std::vector<FlatEdge> edges;
    std::vector<char> queued(S, 0);
    std::vector<int> sources;
    std::vector<uint8_t> tail;
    // follow the edge e: returns the value to store (state or chain reference) and queues the
    // state the walk stands in afterwards when it has out-edges of its own
    auto compress = [&](const Edge& e) -> uint32_t {
        int cur = e.next;
        tail.clear();
        while (!isFinal(cur) && out[size_t(cur)].size() == 1) {
            tail.push_back(uint8_t(out[size_t(cur)][0].ch));
            cur = out[size_t(cur)][0].next;
        }
        uint32_t val;
        int target;
        if (int(tail.size()) >= kMinChain) {
            const uint32_t idx = uint32_t(L.numChains++);
            const uint32_t off = uint32_t(L.tails.size());
            uint32_t inline4 = 0;
            for (size_t b = 0; b < tail.size(); b++) {
                L.tails.push_back(tail[b]);
                if (b < 4) inline4 |= uint32_t(tail[b]) << (8 * b);
            }
            while (L.tails.size() & 3) L.tails.push_back(0);
            const bool leaf = out[size_t(cur)].empty();
            L.chains.push_back(off);
            L.chains.push_back(uint32_t(tail.size()));
            L.chains.push_back(uint32_t(cur) | (leaf ? kLeafFlag : 0u));
            L.chains.push_back(inline4);
            val = kChainFlag | (uint32_t(tail[0]) << kChainByteShift) | idx;
            target = cur;
        } else {
            val = uint32_t(e.next) | (out[size_t(e.next)].empty() ? kLeafPlain : 0u);
            target = e.next;
        }
        if (!out[size_t(target)].empty() && !queued[size_t(target)]) {
            queued[size_t(target)] = 1;
            sources.push_back(target);
        }
        return val;
    };
    auto findEdge = [&](int st, int ch) -> const Edge* {
        for (const Edge& e : out[size_t(st)]) if (e.ch == ch) return &e;
        return nullptr;
    };

    // ---- K-gram prefilter + direct table: walk every K-gram over the pattern alphabet ------------
    // outcome per gram: dies inside the K symbols (result = longest pattern seen, bit set iff
    // non-zero), or alive after K symbols (bit set; next = the edge into the K-th state)
    struct Gram { uint32_t idx; uint32_t next; uint32_t best; };
    std::vector<Gram> grams;
    if (!frontier.empty()) {
        // depth-first over codes so that shared prefixes are walked once
        struct Frame { int state; int best; };
        std::vector<Frame> stack(size_t(K) + 1);
        stack[0] = Frame{m.initialState, 0};
        // iterative enumeration of all code strings of length K (codes < ncodes)
        std::vector<int> pos(size_t(K), -1);
        int depthNow = 0;
        uint32_t idxNow = 0;
        while (depthNow >= 0) {
            if (depthNow == K) { depthNow--; continue; }
            int& dgt = pos[size_t(depthNow)];
            dgt++;
            if (dgt >= ncodes) { dgt = -1; depthNow--; continue; }
            idxNow = (idxNow & ((1u << (B * depthNow)) - 1u)) | (uint32_t(dgt) << (B * depthNow));
            const Frame& f = stack[size_t(depthNow)];
            const Edge* e = (f.state >= 0 && byteOfCode[dgt] >= 0) ? findEdge(f.state, byteOfCode[dgt]) : nullptr;
            if (!e) {
                // the walk dies here: every completion of this prefix reports f.best
                if (f.best != 0) {
                    const int rest = K - depthNow - 1;
                    uint64_t count = 1;
                    for (int r = 0; r < rest; r++) count *= uint64_t(ncodes);
                    for (uint64_t t = 0; t < count; t++) {
                        uint32_t idx = idxNow & ((1u << (B * (depthNow + 1))) - 1u);
                        uint64_t tt = t;
                        for (int r = 0; r < rest; r++) {
                            idx |= uint32_t(tt % uint64_t(ncodes)) << (B * (depthNow + 1 + r));
                            tt /= uint64_t(ncodes);
                        }
                        grams.push_back(Gram{idx, kTrap, uint32_t(f.best)});
                    }
                }
                continue;
            }
            if (depthNow == K - 1) {
                // alive after K symbols: the value of the edge into the K-th state
                grams.push_back(Gram{idxNow, compress(*e), uint32_t(f.best)});
                continue;
            }
            stack[size_t(depthNow) + 1] = Frame{e->next, isFinal(e->next) ? e->next : f.best};
            depthNow++;
        }
    }
    std::sort(grams.begin(), grams.end(), [](const Gram& a, const Gram& b) { return a.idx < b.idx; });
    auto setBit = [&](uint32_t idx) { L.pre2[idx >> 5] |= 0x80000000u >> (idx & 31); };
    bool anyBest = false;
    for (const Gram& g : grams) { setBit(g.idx); anyBest = anyBest || g.best != 0; }
    L.rank2.assign(2048, 0);
    for (size_t w = 0; w < 2048; w++) {
        L.rank2[w] = uint16_t(L.pre2BitsSet);
        L.pre2BitsSet += __builtin_popcount(L.pre2[w]);
    }
    L.next2.assign(std::max<size_t>(grams.size(), 1), kTrap);
    // K == 2: the only pattern inside K-1 symbols is a 1-byte pattern, which the root row already
    // tells (kernels read root[c0] instead): no best2 array, more shared memory for next2
    if (K == 2) anyBest = false;
    if (anyBest) L.best2.assign(L.next2.size(), 0u);
    for (size_t i = 0; i < grams.size(); i++) {  // sorted by idx == rank order
        L.next2[i] = grams[i].next;
        if (anyBest) L.best2[i] = grams[i].best;
    }

    // second prefilter stage: per valid K-gram a 16-bit set of (next byte & 15) values that keep the
    // walk alive; all ones when the K symbols alone already produce a result (then the position
    // must reach the walker whatever follows).  Only worth its shared memory and instructions when
    // the first stage is unselective.
    if (B == 8 && L.pre2BitsSet > 65536 * 3 / 100) {  // small alphabets: the K-gram stage is selective enough
        L.chk2.assign(L.next2.size(), 0);
        for (size_t i = 0; i < grams.size(); i++) {
            const Gram& g = grams[i];
            uint16_t mask = 0;
            if (g.best != 0 || g.next == kTrap) {
                mask = 0xFFFF;  // (trap entries exist only with best != 0)
            } else if (g.next & kChainFlag) {
                mask = uint16_t(1u << ((g.next >> kChainByteShift) & 15u));
            } else {
                const int st = int(g.next & ~kLeafPlain);
                if (isFinal(st)) mask = 0xFFFF;
                else for (const Edge& e : out[size_t(st)]) mask |= uint16_t(1u << (e.ch & 15));
            }
            L.chk2[i] = mask;
        }
    }

    // hashed 4-gram first stage (see pfac_table.h) for every byte-alphabet dictionary whose
    // shared-memory budget holds its 32 KB: measured faster than the exact 2-gram stage from 1,000
    // random patterns (+11 %) to 20,000 Snort-like ones (+29 %)
    size_t hfiltBytes = size_t(kHashFilterWords) * 4;
    // Pair filter (hfiltK = 3) for sparse dictionaries without 1- and 2-byte patterns: ONE lookup decides two
    // start positions.  The positions q and q+1 share the text bytes c1 c2 c3 = text[q+1..q+3]: a pattern that
    // starts at q+1 shows its bytes 0..2 there, one that starts at q its bytes 1..3 (a 3-byte pattern its
    // bytes 1..2 and anything).  Both kinds of 3-grams of every pattern go into the same 256-Kbit table, keyed by
    // y = c1 | c2<<8 | c3<<16 through h = (y * kHashFilterMul) mod 2^24 -- the low 24 bits of the product of
    // the 4-byte text window with the multiplier, which the fourth byte cannot reach: word (h >> 2) & 8191
    // (c1 and seven bits of c2), bit 31 - (h >> 19).  A pair that passes sends both its positions to the
    // walker.  3.5 instead of 7 instructions per start position; 1,000 patterns set 2,000 bits, 0.8 % of the
    // pairs pass.
    bool minLen3 = !frontier.empty();
    for (const Edge& e : out[size_t(m.initialState)]) {
        if (isFinal(e.next)) minLen3 = false;
        for (const Edge& e2 : out[size_t(e.next)]) if (isFinal(e2.next)) minLen3 = false;
    }
    if (B == 8 && minLen3 && hotBudgetBytes >= hfiltBytes && (filterPolicy == kFilterAuto || filterPolicy == kFilterHashed)) {
        L.hfilt.assign(size_t(kHashFilterWords), 0u);
        auto setKey = [&](uint32_t y) {
            const uint32_t h = (y * kHashFilterMul) & 0xFFFFFFu;
            L.hfilt[(h >> 2) & uint32_t(kHashFilterWords - 1)] |= 0x80000000u >> (h >> 19);
        };
        struct Node { int state; uint32_t x; int d; };
        std::vector<Node> todo;
        todo.push_back(Node{m.initialState, 0u, 0});
        while (!todo.empty()) {
            const Node f = todo.back();
            todo.pop_back();
            for (const Edge& e : out[size_t(f.state)]) {
                const uint32_t x = f.x | (uint32_t(e.ch) << (8 * f.d));
                const int d = f.d + 1;
                if (d == 3) {
                    setKey(x);                                      // starts at q+1: bytes 0..2
                    if (isFinal(e.next))                            // a 3-byte pattern starting at q: bytes 1..2, any c3
                        for (uint32_t c = 0; c < 256; c++) setKey((x >> 8) | (c << 16));
                }
                if (d == 4) { setKey(x >> 8); continue; }           // starts at q: bytes 1..3
                if (!out[size_t(e.next)].empty()) todo.push_back(Node{e.next, x, d});
            }
        }
        L.hfiltBitsSet = 0;
        for (uint32_t w : L.hfilt) L.hfiltBitsSet += __builtin_popcount(w);
        if (L.hfiltBitsSet <= kHashFilterWords * 32 / 64) {   // <= 1.6 % of the pairs pass by chance
            L.hfiltK = 3;
            hotBudgetBytes -= hfiltBytes;
        } else {
            L.hfilt.clear();
            L.hfiltBitsSet = 0;
        }
    }
    if (L.hfiltK == 0 && B == 8 && !frontier.empty() && hotBudgetBytes >= hfiltBytes && filterPolicy != kFilterExact) {
        // one bit per 4-gram first, word picked by a hash of (c0, c1 & 127): sparse dictionaries.  A table
        // that comes out dense (> 2 % of its bits, whole words of short patterns aside) is rebuilt with two
        // bits per gram, twice the words when the budget holds them, and the word picked by the text bits
        // themselves, x & (words - 1) = c0 | (c1 & 63 or 31) << 8: the words one first byte can reach are
        // then its own (64 or 32 of them), so a 1-byte pattern fills those and nothing else.  Hashed, its
        // 128 all-ones words are shared with three other (c0, c1) pairs each: on the 20,000-pattern
        // Snort-like dictionary 11 of the 25 survivors per 512 positions were such collisions, and the
        // 32 KB table passed 9 false 4-grams more (64 KB, row-indexed: 11 survivors, 4 of them true).
        for (L.hfiltK = 1; L.hfiltK <= 2; L.hfiltK++) {
            size_t words = size_t(kHashFilterWords);
            if (L.hfiltK == 2 && hotBudgetBytes >= size_t(kHashFilterWordsMax) * 4) words = size_t(kHashFilterWordsMax);
            L.hfilt.assign(words, 0u);
            L.hfiltBitsSet = 0;
            const bool rows = L.hfiltK == 2;
            auto word = [&](uint32_t x) -> uint32_t& {
                return rows ? L.hfilt[x & uint32_t(words - 1)]
                            : L.hfilt[((x * kHashFilterMul) >> 2) & uint32_t(kHashFilterWords - 1)];
            };
            auto setGram = [&](uint32_t x) {
                const uint32_t b1 = uint32_t((uint64_t(x) * kHashFilterMul2) >> 32) & 31u;
                const uint32_t b2 = uint32_t((uint64_t(x) * kHashFilterMul3) >> 32) & 31u;
                word(x) |= (0x80000000u >> b1) | (L.hfiltK == 2 ? 0x80000000u >> b2 : 0u);
            };
            struct Node { int state; uint32_t x; int d; };
            std::vector<Node> todo;
            todo.push_back(Node{m.initialState, 0u, 0});
            while (!todo.empty()) {
                const Node f = todo.back();
                todo.pop_back();
                for (const Edge& e : out[size_t(f.state)]) {
                    const uint32_t x = f.x | (uint32_t(e.ch) << (8 * f.d));
                    const int d = f.d + 1;
                    if (d == 4) { setGram(x); continue; }
                    if (isFinal(e.next)) {  // a pattern shorter than the gram: whatever follows must pass
                        if (d == 1) for (uint32_t c = 0; c < 256; c++) word(x | (c << 8)) = 0xFFFFFFFFu;
                        else if (d == 2) word(x) = 0xFFFFFFFFu;
                        else for (uint32_t c = 0; c < 256; c++) setGram(x | (c << 24));
                    }
                    if (!out[size_t(e.next)].empty()) todo.push_back(Node{e.next, x, d});
                }
            }
            int gramBits = 0;  // all-ones words (1- and 2-byte patterns) pass whatever the number of bits
            for (uint32_t w : L.hfilt) {
                L.hfiltBitsSet += __builtin_popcount(w);
                if (w != 0xFFFFFFFFu) gramBits += __builtin_popcount(w);
            }
            if (L.hfiltK == 2 || gramBits <= kHashFilterWords * 32 / 50) break;
        }
        hfiltBytes = L.hfilt.size() * 4;
        if (size_t(L.hfiltBitsSet) > L.hfilt.size() * 32 / 2 && filterPolicy != kFilterHashed) {
            // saturated (e.g. dozens of 1-byte patterns): it would pass most positions to the walker; the
            // exact 2-gram stage with its inline second stage does better
            L.hfilt.clear();
            L.hfiltK = 0;
            L.hfiltBitsSet = 0;
        } else {
            hotBudgetBytes -= hfiltBytes;
        }
    }
    hfiltBytes = size_t(kHashFilterWords) * 4;

    // hashed 10-mer first stage for 2-bit alphabets with an arithmetic code (DNA): the exact 8-mer set of
    // 5,000 patterns passes 7.3 % of all positions (4,794 of 65,536 8-mers), four walker batches per
    // 1,536-position tile; ten symbols hashed into the same 256-Kbit table with two bits per gram (a
    // blocked Bloom filter) pass the true matches plus well under 1 %.  x = the ten 2-bit codes, symbol i
    // at bits 2i; word (x * kHashFilterMul) >> 19, bits 31 - ((x * kHashFilterMul2) >> 27) and
    // 31 - ((x * kHashFilterMul3) >> 27), products taken mod 2^32.  A pattern shorter than ten symbols
    // sets the bits of all its 4^(10-len) continuations, so neither bytes past the end of the input nor
    // bytes outside the alphabet (which alias to some code) can hide a match; the walker re-checks
    // survivors exactly (lut, pre2, next2) and takes the generic path when a window holds such a byte.
    if (B == 2 && L.codeShift >= 0 && !frontier.empty() && hotBudgetBytes >= hfiltBytes && filterPolicy != kFilterExact) {
        L.hfilt.assign(size_t(kHashFilterWords), 0u);
        L.hfiltK = 2;
        auto setGram = [&](uint32_t x) {
            const uint32_t w = (uint32_t(x * kHashFilterMul) >> 19) & uint32_t(kHashFilterWords - 1);
            L.hfilt[w] |= (0x80000000u >> (uint32_t(x * kHashFilterMul2) >> 27)) |
                          (0x80000000u >> (uint32_t(x * kHashFilterMul3) >> 27));
        };
        struct Node { int state; uint32_t x; int d; };
        std::vector<Node> todo;
        todo.push_back(Node{m.initialState, 0u, 0});
        while (!todo.empty()) {
            const Node f = todo.back();
            todo.pop_back();
            for (const Edge& e : out[size_t(f.state)]) {
                const uint32_t x = f.x | (uint32_t(L.lut[e.ch] & 3u) << (2 * f.d));
                const int d = f.d + 1;
                if (d == kDnaGram) { setGram(x); continue; }
                if (isFinal(e.next)) {  // a pattern shorter than the gram: whatever follows must pass
                    const uint32_t rest = 1u << (2 * (kDnaGram - d));
                    for (uint32_t t = 0; t < rest; t++) setGram(x | (t << (2 * d)));
                }
                if (!out[size_t(e.next)].empty()) todo.push_back(Node{e.next, x, d});
            }
        }
        L.hfiltBitsSet = 0;
        for (uint32_t w : L.hfilt) L.hfiltBitsSet += __builtin_popcount(w);
        if (L.hfiltBitsSet > kHashFilterWords * 32 / 2 && filterPolicy != kFilterHashed) {
            L.hfilt.clear();   // saturated (many very short patterns): the exact 8-mer stage does better
            L.hfiltK = 0;
            L.hfiltBitsSet = 0;
        } else {
            hotBudgetBytes -= hfiltBytes;
        }
    }

    // deeper transitions (source depth >= K): hash rows, hot by depth
    for (size_t qi = 0; qi < sources.size(); qi++) {
        const int s = sources[qi];
        for (const Edge& e : out[size_t(s)])
            edges.push_back(FlatEdge{edgeKey(s, e.ch), int(compress(e)), depth[size_t(s)]});
    }
    // generic path (K > 2 only): walks that start with a byte outside the alphabet or too close
    // to the end of the input go root row -> hash rows from depth 1.  Those edges are rarely
    // used; they are given depth 0x7fff so that they always land in the cold table.
    if (K > 2 && !frontier.empty()) {
        std::vector<int> low;  // states of depth 1..K-1 with out-edges
        for (const Edge& e : out[size_t(m.initialState)])
            if (!out[size_t(e.next)].empty()) low.push_back(e.next);
        for (size_t qi = 0; qi < low.size(); qi++) {
            const int s = low[qi];
            for (const Edge& e : out[size_t(s)]) {
                const uint32_t val = uint32_t(e.next) | (out[size_t(e.next)].empty() ? kLeafPlain : 0u);
                edges.push_back(FlatEdge{edgeKey(s, e.ch), int(val), 0x7fff});
                if (depth[size_t(e.next)] < K && !out[size_t(e.next)].empty()) low.push_back(e.next);
                else if (depth[size_t(e.next)] >= K && !out[size_t(e.next)].empty() && !queued[size_t(e.next)]) {
                    // first reached through the generic path only: its out-edges must exist too
                    queued[size_t(e.next)] = 1;
                    const size_t before = sources.size();
                    sources.push_back(e.next);
                    for (size_t q2 = before; q2 < sources.size(); q2++) {
                        const int s2 = sources[q2];
                        for (const Edge& e2 : out[size_t(s2)])
                            edges.push_back(FlatEdge{edgeKey(s2, e2.ch), int(compress(e2)), depth[size_t(s2)]});
                    }
                }
            }
        }
    }
    L.hashEdges = int(edges.size());
    while (L.tails.size() & 15) L.tails.push_back(0);
    if (L.chains.empty()) L.chains.assign(4, 0u);  // keep device pointers valid

    // ---- hot/cold split by source depth.  The walker knows its depth (= bytes consumed), so
    // which table to probe costs no lookup.  Load factor 0.5: buckets(2 slots) = #edges.
    std::stable_sort(edges.begin(), edges.end(),
                     [](const FlatEdge& a, const FlatEdge& b) { return a.depth < b.depth; });
    std::vector<size_t> upto(size_t(L.maxDepth) + 2, 0);  // upto[d] = #edges with depth < d
    size_t genericOnly = 0;                               // depth 0x7fff: generic-path edges, never hot
    for (const FlatEdge& e : edges) {
        if (e.depth == 0x7fff) genericOnly++;
        else upto[size_t(e.depth) + 1]++;
    }
    for (size_t d = 1; d < upto.size(); d++) upto[d] += upto[d - 1];
    // shared-memory budget: next2 first (touched by every survivor), then hash rows by depth,
    // and chains + tails too when everything fits
    // the check stage only works from shared memory: it comes first, or it is dropped
    const size_t chkBytes = ((L.chk2.size() * 2 + 15) / 16) * 16;
    if (chkBytes > hotBudgetBytes) {
        L.chk2.clear();
    } else {
        hotBudgetBytes -= chkBytes;
    }
    const size_t next2Bytes = ((L.next2.size() * 4 + 15) / 16) * 16 * (L.best2.empty() ? 1 : 2);
    if (next2Bytes <= hotBudgetBytes) {
        L.next2Hot = true;
        hotBudgetBytes -= next2Bytes;
    } else {
        hotBudgetBytes = 0;  // deeper rows are colder than next2: keep them all in L2 as well
    }
    const size_t chainBytes = L.chains.size() * 4 + L.tails.size();
    int H = 1;
    if (L.next2Hot && (edges.size() - genericOnly) * 16 + chainBytes <= hotBudgetBytes) {
        H = int(upto.size()) - 1;      // everything in shared memory, chains and tails too
        L.chainsHot = true;
    } else {
        const size_t hotSlots = hotBudgetBytes / 16;
        for (int d = 2; d < int(upto.size()); d++) {  // upto[] is non-decreasing
            if (upto[size_t(d)] <= hotSlots) H = d;
            else break;
        }
        if (upto[size_t(H)] == 0) H = 1;
    }
    L.hotDepth = H;
    const size_t nHot = upto[size_t(H)];
    const size_t nCold = edges.size() - nHot;
    L.hotBuckets = uint32_t(nHot);         // 0 => kernels never probe the hot table
    // cold rows live in L2 where space is cheap and every extra bucket probe is a dependent L2
    // access: load factor 0.25 (a miss leaves its home bucket 9 % of the time instead of 26 %)
    L.coldBuckets = uint32_t(std::max<size_t>(nCold * 2, 1));

    static const uint32_t muls[] = {0x9E3779B1u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu,
                                    0x165667B1u, 0xD3A2646Du, 0xFD7046C5u, 0xB55A4F09u};
    int bestHot = 1 << 30, bestCold = 1 << 30;
    std::vector<uint32_t> hot, cold;
    for (uint32_t mul : muls) {
        int ph = 0, pc = 0;
        if (L.hotBuckets) ph = fillBuckets(edges, 0, nHot, L.hotBuckets, mul, hot);
        else hot.clear();
        pc = fillBuckets(edges, nHot, edges.size(), L.coldBuckets, mul, cold);
        if (ph < bestHot || (ph == bestHot && pc < bestCold)) {
            bestHot = ph; bestCold = pc; L.mul = mul;
            L.hot.swap(hot); L.cold.swap(cold);
        }
    }
    L.hotMaxProbe = L.hotBuckets ? bestHot : 0;
    L.coldMaxProbe = bestCold;
}

// ---- compiled-table files -----------------------------------------------------------------------
namespace {

constexpr char kFileMagic[8] = {'P', 'F', 'A', 'C', 'B', '2', '0', '0'};
constexpr uint32_t kFileVersion = 6;  // bump whenever Machine / DeviceLayout or their meaning change

struct Writer {
    std::string buf;
    void raw(const void* p, size_t n) { buf.append(static_cast<const char*>(p), n); }
    template <typename T> void pod(const T& v) { raw(&v, sizeof(T)); }
    template <typename T> void vec(const std::vector<T>& v) {
        pod<uint64_t>(v.size());
        if (!v.empty()) raw(v.data(), v.size() * sizeof(T));
    }
};

struct Reader {
    const char* p;
    const char* end;
    bool ok = true;
    void raw(void* dst, size_t n) {
        if (!ok || size_t(end - p) < n) { ok = false; return; }
        memcpy(dst, p, n);
        p += n;
    }
    template <typename T> void pod(T& v) { raw(&v, sizeof(T)); }
    template <typename T> void vec(std::vector<T>& v) {
        uint64_t n = 0;
        pod(n);
        if (!ok || n > uint64_t(end - p) / sizeof(T)) { ok = false; return; }
        v.resize(size_t(n));
        if (n) raw(v.data(), size_t(n) * sizeof(T));
    }
};

uint64_t checksum(const char* p, size_t n) {  // FNV-1a, 64 bit
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) h = (h ^ uint8_t(p[i])) * 1099511628211ull;
    return h;
}

void putMachine(Writer& w, const Machine& m) {
    w.pod<uint64_t>(m.image.size());
    w.raw(m.image.data(), m.image.size());
    w.pod(m.numPatterns); w.pod(m.numFinal); w.pod(m.initialState); w.pod(m.numStates);
    w.pod(m.maxPatternLen); w.pod(m.numLeaves);
    std::vector<uint64_t> off64(m.sortedOff.begin(), m.sortedOff.end());
    w.vec(off64);
    w.vec(m.sortedId);
    w.vec(m.lenById);
    std::vector<uint64_t> offById64(m.offById.begin(), m.offById.end());
    w.vec(offById64);
    // rows, flattened: per-state edge counts, then all edges in state order / insertion order
    std::vector<uint32_t> counts(m.rows.size());
    std::vector<Edge> flat;
    for (size_t i = 0; i < m.rows.size(); i++) {
        counts[i] = uint32_t(m.rows[i].size());
        flat.insert(flat.end(), m.rows[i].begin(), m.rows[i].end());
    }
    w.vec(counts);
    w.vec(flat);
}

void getMachine(Reader& r, Machine& m) {
    uint64_t n = 0;
    r.pod(n);
    if (!r.ok || n > uint64_t(r.end - r.p)) { r.ok = false; return; }
    m.image.assign(r.p, size_t(n));
    r.p += n;
    r.pod(m.numPatterns); r.pod(m.numFinal); r.pod(m.initialState); r.pod(m.numStates);
    r.pod(m.maxPatternLen); r.pod(m.numLeaves);
    std::vector<uint64_t> off64, offById64;
    r.vec(off64);
    r.vec(m.sortedId);
    r.vec(m.lenById);
    r.vec(offById64);
    m.sortedOff.assign(off64.begin(), off64.end());
    m.offById.assign(offById64.begin(), offById64.end());
    std::vector<uint32_t> counts;
    std::vector<Edge> flat;
    r.vec(counts);
    r.vec(flat);
    if (!r.ok) return;
    uint64_t total = 0;
    for (uint32_t c : counts) total += c;
    if (total != flat.size() || int64_t(counts.size()) < int64_t(m.numStates)) { r.ok = false; return; }
    m.rows.assign(counts.size(), std::vector<Edge>());
    size_t at = 0;
    for (size_t i = 0; i < counts.size(); i++) {
        m.rows[i].assign(flat.begin() + long(at), flat.begin() + long(at + counts[i]));
        at += counts[i];
    }
}

void putLayout(Writer& w, const DeviceLayout& L) {
    w.raw(L.root, sizeof(L.root));
    w.raw(L.lut, sizeof(L.lut));
    w.pod(L.codeBits); w.pod(L.gramLen); w.pod(L.codeShift);
    w.vec(L.pre2); w.vec(L.rank2); w.vec(L.next2); w.vec(L.best2); w.vec(L.chk2); w.vec(L.hfilt);
    w.pod(L.hfiltK); w.pod(L.hfiltBitsSet);
    w.pod<uint8_t>(L.next2Hot);
    w.vec(L.hot); w.vec(L.cold); w.vec(L.chains); w.vec(L.tails);
    w.pod(L.hotBuckets); w.pod(L.coldBuckets); w.pod(L.mul); w.pod(L.hotDepth);
    w.pod<uint8_t>(L.chainsHot);
    w.pod(L.maxDepth); w.pod(L.numEdges); w.pod(L.hashEdges); w.pod(L.numChains);
    w.pod(L.hotMaxProbe); w.pod(L.coldMaxProbe); w.pod(L.pre2BitsSet); w.pod(L.rootFanout);
}

void getLayout(Reader& r, DeviceLayout& L) {
    r.raw(L.root, sizeof(L.root));
    r.raw(L.lut, sizeof(L.lut));
    r.pod(L.codeBits); r.pod(L.gramLen); r.pod(L.codeShift);
    r.vec(L.pre2); r.vec(L.rank2); r.vec(L.next2); r.vec(L.best2); r.vec(L.chk2); r.vec(L.hfilt);
    r.pod(L.hfiltK); r.pod(L.hfiltBitsSet);
    uint8_t b = 0;
    r.pod(b); L.next2Hot = b != 0;
    r.vec(L.hot); r.vec(L.cold); r.vec(L.chains); r.vec(L.tails);
    r.pod(L.hotBuckets); r.pod(L.coldBuckets); r.pod(L.mul); r.pod(L.hotDepth);
    r.pod(b); L.chainsHot = b != 0;
    r.pod(L.maxDepth); r.pod(L.numEdges); r.pod(L.hashEdges); r.pod(L.numChains);
    r.pod(L.hotMaxProbe); r.pod(L.coldMaxProbe); r.pod(L.pre2BitsSet); r.pod(L.rootFanout);
    // sizes the kernels rely on
    if (r.ok && (L.pre2.size() != 2048 || L.rank2.size() != 2048 || L.next2.empty() ||
                 (!L.hfilt.empty() && L.hfilt.size() != size_t(kHashFilterWords) && L.hfilt.size() != size_t(kHashFilterWordsMax)) ||
                 L.hot.size() != size_t(L.hotBuckets) * 4 || L.cold.size() != size_t(L.coldBuckets) * 4 ||
                 (L.chains.size() & 3) || (L.tails.size() & 15)))
        r.ok = false;
}

// A compiled-table file is input like any other: a truncated, mismatched or hostile image must be
// rejected here (the caller then compiles from the pattern file) instead of indexing out of bounds
// in the dump, in a recompile for another budget, or on the device.
#ifdef PFAC_VALID_DEBUG
#define PFAC_INVALID do { fprintf(stderr, "compiled table rejected at %s:%d\n", __FILE__, __LINE__); return false; } while (0)
#else
#define PFAC_INVALID return false
#endif
bool validMachine(const Machine& m) {
    if (m.numPatterns < 0 || m.numFinal != m.numPatterns || m.initialState != m.numFinal + 1) PFAC_INVALID;
    if (m.numStates <= m.initialState || m.numStates > kMaxStates || m.maxPatternLen < 0 || m.numLeaves < 0) PFAC_INVALID;
    if (m.lenById.size() != size_t(m.numFinal) + 1 || m.offById.size() != size_t(m.numFinal) + 1) PFAC_INVALID;
    if (m.sortedId.size() != size_t(m.numPatterns) || m.sortedOff.size() != size_t(m.numPatterns)) PFAC_INVALID;
    if (m.rows.size() < size_t(m.numStates)) PFAC_INVALID;
    for (int id = 1; id <= m.numFinal; id++) {
        const int len = m.lenById[size_t(id)];
        if (len <= 0 || len > m.maxPatternLen) PFAC_INVALID;
        if (m.offById[size_t(id)] > m.image.size() || size_t(len) > m.image.size() - m.offById[size_t(id)]) PFAC_INVALID;
    }
    for (int i = 0; i < m.numPatterns; i++) {
        if (m.sortedId[size_t(i)] < 1 || m.sortedId[size_t(i)] > m.numFinal) PFAC_INVALID;
        if (m.sortedOff[size_t(i)] >= m.image.size()) PFAC_INVALID;
    }
    for (size_t st = 0; st < m.rows.size(); st++)
        for (const Edge& e : m.rows[st])
            if (e.ch < 0 || e.ch >= kCharSet || e.next <= 0 || e.next >= m.numStates) PFAC_INVALID;
    return true;
}

bool validLayout(const DeviceLayout& L, const Machine& m) {
    if (!((L.codeBits == 8 || L.codeBits == 4 || L.codeBits == 2) && L.gramLen == 16 / L.codeBits)) PFAC_INVALID;
    if (L.pre2.size() != 2048 || L.rank2.size() != 2048) PFAC_INVALID;
    size_t bits = 0;
    for (size_t w = 0; w < 2048; w++) {
        if (L.rank2[w] != bits) PFAC_INVALID;
        bits += size_t(__builtin_popcount(L.pre2[w]));
    }
    if (L.next2.size() != std::max<size_t>(bits, 1)) PFAC_INVALID;
    if (!L.best2.empty() && L.best2.size() != L.next2.size()) PFAC_INVALID;
    if (!L.chk2.empty() && L.chk2.size() != L.next2.size()) PFAC_INVALID;
    if (L.codeShift < -1 || L.codeShift > 6 || (L.codeShift >= 0 && L.codeBits != 2)) PFAC_INVALID;
    if (!L.hfilt.empty() && (L.hfiltK < 1 || L.hfiltK > 3 ||
                             !(L.codeBits == 8 || (L.codeBits == 2 && L.codeShift >= 0 && L.hfiltK == 2)))) PFAC_INVALID;
    // 64 KB only for the row-indexed two-bit filter of byte alphabets
    if (!L.hfilt.empty() && L.hfilt.size() != size_t(kHashFilterWords) &&
        !(L.hfilt.size() == size_t(kHashFilterWordsMax) && L.codeBits == 8 && L.hfiltK == 2)) PFAC_INVALID;
    if (L.codeShift >= 0)   // the kernels code text bytes arithmetically: lut must agree for every alphabet byte
        for (int c = 0; c < kCharSet; c++)
            if (!(L.lut[c] & 0x80) && L.lut[c] != uint8_t((c >> L.codeShift) & 3)) PFAC_INVALID;
    if (L.hot.size() != size_t(L.hotBuckets) * 4 || L.cold.size() != size_t(L.coldBuckets) * 4 || L.coldBuckets == 0) PFAC_INVALID;
    if ((L.chains.size() & 3) || (L.tails.size() & 15) || L.hotDepth < 1) PFAC_INVALID;
    if (L.numChains < 0 || size_t(L.numChains) > L.chains.size() / 4) PFAC_INVALID;
    const size_t nchains = size_t(L.numChains);   // an empty table still holds one all-zero record
    const uint32_t states = uint32_t(m.numStates);
    auto okValue = [&](uint32_t v) {   // what next2 / a hash slot may hold
        if (v == kTrap) return true;
        if (v & kChainFlag) return size_t(v & kChainIndexMask) < nchains;
        return (v & ~kLeafPlain) < states && (v & ~kLeafPlain) != 0;
    };
    for (uint32_t v : L.next2) if (!okValue(v)) PFAC_INVALID;
    for (uint32_t v : L.best2) if (v > uint32_t(m.numFinal)) PFAC_INVALID;
    for (int c = 0; c < kCharSet; c++) if (L.root[c] < -1 || L.root[c] >= m.numStates || L.root[c] == 0) PFAC_INVALID;
    for (const std::vector<uint32_t>* tab : {&L.hot, &L.cold}) {
        bool terminator = tab->empty();   // probing ends at a bucket whose second slot is empty
        for (size_t b = 0; b + 3 < tab->size(); b += 4) {
            const uint32_t* e = tab->data() + b;
            for (int sl = 0; sl < 2; sl++) {
                if (e[2 * sl] == kEmptyKey) continue;
                if ((e[2 * sl] >> 8) >= states || !okValue(e[2 * sl + 1]) || e[2 * sl + 1] == kTrap) PFAC_INVALID;
            }
            if (e[0] == kEmptyKey && e[2] != kEmptyKey) PFAC_INVALID;   // slot 0 fills first
            terminator = terminator || e[2] == kEmptyKey;
        }
        if (!terminator) PFAC_INVALID;
    }
    for (size_t i = 0; i < nchains; i++) {
        const uint32_t off = L.chains[4 * i], len = L.chains[4 * i + 1], end = L.chains[4 * i + 2];
        if ((off & 3) || len == 0 || len > uint32_t(m.maxPatternLen)) PFAC_INVALID;
        if (size_t(off) + ((size_t(len) + 3) & ~size_t(3)) > L.tails.size()) PFAC_INVALID;
        if ((end & ~kLeafFlag) >= states || (end & ~kLeafFlag) == 0) PFAC_INVALID;
    }
    return true;
}

#undef PFAC_INVALID

}  // namespace

bool saveCompiled(const char* filename, const Machine& m, const std::vector<const CompiledLayout*>& layouts) {
    Writer w;
    putMachine(w, m);
    w.pod<uint32_t>(uint32_t(layouts.size()));
    for (const CompiledLayout* c : layouts) {
        w.pod(c->budget);
        w.pod(c->policy);
        putLayout(w, c->layout);
    }
    FILE* fp = fopen(filename, "wb");
    if (!fp) return false;
    const uint64_t size = w.buf.size(), sum = checksum(w.buf.data(), w.buf.size());
    const uint32_t consts[3] = {kHashFilterMul, kHashFilterMul2, kHashFilterMul3};
    bool ok = fwrite(kFileMagic, 1, 8, fp) == 8 && fwrite(&kFileVersion, 4, 1, fp) == 1 &&
              fwrite(consts, 4, 3, fp) == 3 && fwrite(&size, 8, 1, fp) == 1 && fwrite(&sum, 8, 1, fp) == 1 &&
              fwrite(w.buf.data(), 1, w.buf.size(), fp) == w.buf.size();
    ok = (fclose(fp) == 0) && ok;
    return ok;
}

bool loadCompiled(const char* filename, Machine& m, std::vector<CompiledLayout>& layouts) {
    FILE* fp = fopen(filename, "rb");
    if (!fp) return false;
    char magic[8];
    uint32_t version = 0, consts[3] = {0, 0, 0};
    uint64_t size = 0, sum = 0;
    bool ok = fread(magic, 1, 8, fp) == 8 && fread(&version, 4, 1, fp) == 1 && fread(consts, 4, 3, fp) == 3 &&
              fread(&size, 8, 1, fp) == 1 && fread(&sum, 8, 1, fp) == 1;
    ok = ok && !memcmp(magic, kFileMagic, 8) && version == kFileVersion && consts[0] == kHashFilterMul &&
         consts[1] == kHashFilterMul2 && consts[2] == kHashFilterMul3 && size < (uint64_t(1) << 34);
    std::string buf;
    if (ok) {
        buf.resize(size_t(size));
        ok = fread(&buf[0], 1, buf.size(), fp) == buf.size() && fgetc(fp) == EOF;
    }
    fclose(fp);
    if (!ok || checksum(buf.data(), buf.size()) != sum) return false;
    Reader r{buf.data(), buf.data() + buf.size()};
    Machine mm;
    getMachine(r, mm);
    uint32_t n = 0;
    r.pod(n);
    if (!r.ok || n > 16) return false;
    std::vector<CompiledLayout> ls(n);
    for (uint32_t i = 0; i < n && r.ok; i++) {
        r.pod(ls[i].budget);
        r.pod(ls[i].policy);
        getLayout(r, ls[i].layout);
    }
    if (!r.ok || r.p != r.end) return false;
    if (!validMachine(mm)) return false;
    for (const CompiledLayout& c : ls)
        if (!validLayout(c.layout, mm)) return false;
    m = std::move(mm);
    layouts = std::move(ls);
    return true;
}

}  // namespace pfac
