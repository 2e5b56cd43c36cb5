// pfac_table.h -- host-side table compiler: pattern image -> reference-numbered automaton
// -> B200 device layout.  No CUDA dependency.
//
// Reference behaviour restated (not copied) from PFAC/src/PFAC_reorder_Table.cpp:121-329
// (parse, order, trie numbering) and PFAC/src/PFAC.cpp:345-402 (which edge wins in the
// matching table).  The device layout is new: see DESIGN.md "Device layout".
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace pfac {

constexpr int kCharSet = 256;
constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;   // hash slot never holds (state<<8|ch) == this
constexpr int kMaxStates = (1 << 24) - 2;     // state ids must fit 24 bits of the hash key

struct Edge {
    int ch;
    int next;
};

// The automaton with the reference's state numbering (needed for a byte-identical dump and
// because final-state ids ARE the pattern ids the kernels emit).
struct Machine {
    std::string image;               // pattern file bytes + '\n' sentinel
    int numPatterns = 0;             // k
    int numFinal = 0;                // k
    int initialState = 0;            // k+1   (reference PFAC.cpp:693)
    int numStates = 0;               // next unused id, counts the unused state 0
    int maxPatternLen = 0;
    int numLeaves = 0;
    std::vector<size_t> sortedOff;   // pattern start offsets into image, sorted order
    std::vector<int> sortedId;       // original 1-based id, sorted order
    std::vector<int> lenById;        // [0..k], [0] = 0
    std::vector<size_t> offById;     // [0..k]
    std::vector<std::vector<Edge>> rows;  // per state, insertion order, duplicates kept
};

// Returns a PFAC_status_t value (0 = success).
int buildMachine(const char* image, size_t size, Machine& m);
int buildMachineFromArrays(const char* const* ptrs, const size_t* lens, size_t n, Machine& m);

// Text dump, byte-identical to reference PFAC_dumpTransitionTable (PFAC.cpp:1188-1246).
void dumpMachine(const Machine& m, FILE* fp);

// What gets uploaded.
//   hash rows: buckets of 4 x uint32 {key0, val0, key1, val1}; key = state<<8|ch; slot 0 is
//     filled before slot 1; bucket = umulhi(key * mul, nbuckets); linear probing.
//     val = next state, or kChainFlag | first tail byte << kChainByteShift | chain index when the
//     edge starts a compressed chain: the walker checks that one byte against the text before it
//     fetches the chain record (which usually sits in L2), so most false candidates never do.
//   chains: 4 x uint32 {tail offset (bytes, multiple of 4), len, end state | kLeafFlag,
//     first 4 tail bytes}: after the edge's own byte, `len` more bytes must equal the tail; the
//     walk then stands in the end state.  Chain interiors are non-final single-child states, so
//     skipping them cannot change any result; kLeafFlag = the end state has no out-edges.
constexpr uint32_t kChainFlag = 0x80000000u;
constexpr int kChainByteShift = 23;            // bits 23..30 of a chain reference: first tail byte
constexpr uint32_t kChainIndexMask = (1u << kChainByteShift) - 1u;
constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr int kMinChain = 2;          // compress runs of at least this many single-child states

//   lut / pre2 / rank2 / next2 / best2: the first K symbols need no hashing.  Every byte has a
//     b-bit code (lut[c] = code, bit 7 set = byte occurs in no pattern); b = 8 (identity, K = 2),
//     4 (K = 4) or 2 (K = 8) depending on the pattern alphabet.  pre2 is a 65536-bit set over
//     idx = sum code(c_i) << (b*i), stored bit-reversed inside each 32-bit word (idx -> word
//     idx>>5, bit 31-(idx&31)) so that `word << (idx&31)` moves the wanted bit to the sign
//     position.  A bit is set iff a walk over those K symbols can produce a result.
//     rank2[w] = number of set bits in words [0,w); next2[rank] = what the walk holds after
//     its K-th symbol: a state (kLeafPlain = it has no out-edges), kChainFlag|chain index, or
//     kTrap when the walk dies inside the K symbols; best2[rank] = longest pattern matched
//     within the first K-1 symbols (0 if none; array omitted when it would be all zero).
//     K-grams containing a byte outside the alphabet, or cut off by the end of the input, take
//     the generic path from the root row (hash rows then also hold the edges of depth < K).
constexpr uint32_t kTrap = 0xFFFFFFFFu;
constexpr uint32_t kLeafPlain = 0x40000000u;  // plain state value: no out-edges, stop after it

//   hfilt (byte alphabets, when the shared-memory budget holds it): with tens of thousands of
//     patterns the exact 2-gram set passes a third of all text positions, and even with 1,000 it
//     passes 1.5 %.  A hashed 4-gram filter takes its place as the per-position test:
//     x = c0|c1<<8|c2<<16|c3<<24 picks one word and one or two bits of it,
//     31 - (umulhi(x, kHashFilterMul2) & 31) and 31 - (umulhi(x, kHashFilterMul3) & 31).
//       hfiltK = 1 (sparse tables, C2: 0.4 % of the positions pass): 8192 words, word
//     ((x * kHashFilterMul) >> 2) & 8191 -- bits 2..14 of the product depend on c0 and seven bits of c1
//     only -- and the first bit.
//       hfiltK = 2 (dense tables): both bits (two hash functions into one word, a blocked Bloom filter)
//     and the word is x & (words - 1) = c0 | (c1 & 63) << 8 with 16384 words (64 KB) when the budget
//     holds them, c0 | (c1 & 31) << 8 with 8192.  Row-indexed: the words a first byte can reach are its
//     own, so the all-ones words of a 1-byte pattern catch that byte and nothing else (hashed, each was
//     shared with three other 2-grams: 11 of C3's 25 survivors per 512 positions).
//       hfiltK = 3 (sparse tables whose patterns all have three bytes or more): the pair filter, one lookup
//     per two start positions, keyed by the three text bytes they share (compileLayout).
//     With hfiltK = 1, 2 the word is chosen by (c0, c1), the bits by all four bytes.  A 4-byte prefix of a
//     pattern sets its bits, a 3-byte pattern the bits of its 256 continuations, a 2-byte pattern its
//     whole word, a 1-byte pattern every word of (c0,*).  Bytes past the end of the input therefore
//     never hide a short match.
//     The walker re-checks survivors against pre2 (+ chk2) exactly.  (One IMAD + one LOP3 give the
//     byte offset of the word, one IMAD.HI each bit: the multiplies run on the FMA pipe, which the
//     kernels otherwise leave idle, instead of the ALU pipe that bounds them.)
constexpr uint32_t kHashFilterMul = 0x9E3779B1u;
constexpr uint32_t kHashFilterMul2 = 0x85EBCA6Bu;
constexpr uint32_t kHashFilterMul3 = 0xC2B2AE35u;
constexpr int kHashFilterWords = 8192;           // 32 KB of shared memory
constexpr int kHashFilterWordsMax = 16384;       // row-indexed two-bit filter when the budget holds 64 KB
constexpr int kDnaGram = 10;                     // symbols hashed by the first stage of 2-bit alphabets
enum FilterPolicy { kFilterAuto = 0, kFilterExact = 1, kFilterHashed = 2, kFilterNoPair = 3 };

struct DeviceLayout {
    int32_t root[kCharSet];          // next state from the initial state, -1 = trap
    uint8_t lut[kCharSet];           // symbol code | 0x80 if the byte occurs in no pattern
    int codeBits = 8;                // b
    int gramLen = 2;                 // K = 16 / b
    int codeShift = -1;              // b = 2 only: code = (byte >> codeShift) & 3 for every alphabet byte, -1 = lut only
    std::vector<uint32_t> pre2;      // 2048 words, bit-reversed within each word
    std::vector<uint16_t> rank2;     // 2048 prefix popcounts
    std::vector<uint32_t> next2;     // one entry per set bit, in idx order
    std::vector<uint32_t> best2;     // parallel to next2; empty when all zero
    // chk2[rank]: which (byte & 15) values can follow the K-gram and still lead somewhere; 0xFFFF
    // when the K-gram alone already yields a result.  Used as a second prefilter stage (in shared
    // memory) when the first one lets many positions through; empty = stage off.
    std::vector<uint16_t> chk2;
    std::vector<uint32_t> hfilt;     // kHashFilterWords (or, hfiltK == 2, kHashFilterWordsMax) words, or empty (exact 2-gram first stage)
    int hfiltK = 0;                  // 1: one bit per lookup (sparse table, hashed word index); 2: two bits (dense table; byte
                                     // alphabets: word index x & (words - 1)); 3: pair filter, one bit per TWO start
                                     // positions keyed by the three text bytes they share (see compileLayout)
    int hfiltBitsSet = 0;
    bool next2Hot = false;           // next2 (+ best2) fit the shared-memory budget
    std::vector<uint32_t> hot;       // edges with source depth in [K,hotDepth)  -> smem
    std::vector<uint32_t> cold;      // edges with source depth >= hotDepth      -> global/L2
    std::vector<uint32_t> chains;    // 4 words per chain record
    std::vector<uint8_t> tails;      // chain tail bytes, each tail padded to 4
    uint32_t hotBuckets = 0, coldBuckets = 0;
    uint32_t mul = 0x9E3779B1u;
    int hotDepth = 1;
    bool chainsHot = false;          // chain records + tails fit the shared-memory budget too
    int maxDepth = 0;
    int numEdges = 0;                // transitions of the (uncompressed) automaton, root included
    int hashEdges = 0;               // entries in hot + cold after chain compression
    int numChains = 0;
    int hotMaxProbe = 0, coldMaxProbe = 0;
    int pre2BitsSet = 0;
    int rootFanout = 0;
    size_t deviceBytes() const {
        return sizeof(root) + sizeof(lut) + hfilt.size() * 4 + pre2.size() * 4 + rank2.size() * 2 + next2.size() * 4 + chk2.size() * 2 +
               best2.size() * 4 + hot.size() * 4 +
               cold.size() * 4 + chains.size() * 4 + tails.size();
    }
};

// hotBudgetBytes: shared-memory bytes the kernels may spend on hash rows (+ chains and tails
// when everything fits).
// filterPolicy: kFilterAuto / kFilterHashed use the hashed 4-gram filter for byte alphabets when the
// budget holds it; kFilterExact keeps the exact K-gram stage (A/B measurements, tests).
void compileLayout(const Machine& m, size_t hotBudgetBytes, DeviceLayout& L, int filterPolicy = kFilterAuto);

// ---- compiled-table files -----------------------------------------------------------------------
// Binary image of a Machine plus any number of DeviceLayouts (each tagged with the shared-memory
// budget and filter policy it was compiled for), so that a large dictionary is parsed, sorted,
// numbered and laid out once and later processes only read it back.  Little-endian, versioned,
// with a checksum; loadCompiled rejects anything it does not recognise (returns false), and the
// caller then compiles from the pattern file as usual.
struct CompiledLayout {
    uint64_t budget = 0;
    int32_t policy = kFilterAuto;
    DeviceLayout layout;
};
bool saveCompiled(const char* filename, const Machine& m, const std::vector<const CompiledLayout*>& layouts);
bool loadCompiled(const char* filename, Machine& m, std::vector<CompiledLayout>& layouts);

}  // namespace pfac
