// pfac_kernels.cu -- sm_100a kernels of the PFAC matching path.
//
// What is computed is the reference's per-position failureless walk
// (reference PFAC/src/PFAC_CPU.cpp:60-100 is the spec; PFAC_kernel.cu:377-458 and
// PFAC_reduce_kernel.cu:639-867 are the 2012 GPU forms being replaced).  How it is computed
// is new -- see DESIGN.md:
//   * a prefilter in shared memory rejects most start positions with one LDS -- a 256-Kbit hashed
//     4-gram filter for byte alphabets, a hashed 10-mer filter over arithmetic 2-bit codes for
//     DNA-like alphabets, an exact 64-Kbit K-gram set for the other symbol-coded small alphabets;
//     survivors are compacted into a per-warp queue;
//   * survivors are walked 32 at a time (one per lane, until the batch's longest walk ends): root row and
//     shallow ("hot") hash rows in shared memory, deep ("cold") rows through L1/L2, and
//     path-compressed chains whose tail bytes are compared directly against the text;
//   * dense kernel: one persistent 32-warp CTA per SM, every warp runs its own pipeline --
//     1,536-byte input tiles (+halo) arrive by 1-D TMA bulk copies on per-warp mbarriers, the tile's
//     6 KB of zeros leave at once (one TMA bulk store from a shared zero buffer, or plain 16-byte stores
//     in the HBM-bound sparse-table kernel) and the few matches are stored over them; no CTA-wide
//     barrier in the loop, so one long walk delays one warp, not a block;
//   * reduce kernel: the same pipeline on 1,536-position warp tiles; matches are parked per walker
//     batch (shared-memory ring + a spill ring in global memory) and written in position order by a
//     round-structured look-back across CTAs, one pass; on several GPUs the same kernel exchanges
//     the per-GPU match counts over peer memory (NVLink) and takes their exclusive prefix.
#include "pfac_kernels.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace pfac {

namespace {

constexpr int kPosPerThread = 16;
constexpr int kWarpTile = 32 * kPosPerThread;  // 512 start positions per warp per iteration
constexpr uint32_t kEmpty = 0xFFFFFFFFu;       // empty hash slot / trap
constexpr uint32_t kChainBit = 0x80000000u;
constexpr int kChainByteShift = 23;               // chain reference: bits 23..30 = first tail byte
constexpr uint32_t kChainIndexMask = (1u << kChainByteShift) - 1u;
constexpr uint32_t kLeafPlainBit = 0x40000000u;  // plain state without out-edges
constexpr uint32_t kHashFilterMul = kKernelHashFilterMul;
constexpr uint32_t kHashFilterMul2 = kKernelHashFilterMul2;
constexpr uint32_t kHashFilterMul3 = kKernelHashFilterMul3;
constexpr int kHashFilterWords = kKernelHashFilterWords;
constexpr unsigned kSlowFlag = 0x8000u;           // queue entry: walk from the root row (generic path)
constexpr int kMaxSmem = 232448;               // 227 KB opt-in dynamic shared memory per CTA

// ---- dense kernel geometry -------------------------------------------------------------------
constexpr int kDenseWarps = 32;
constexpr int kDenseThreads = kDenseWarps * 32;
constexpr int kDenseMaxHalo = 64;              // staged halo cap; walks that run further read global memory

// CTA-tile aggregate word: [63:62] != 0 once published, [61:0] match count
constexpr unsigned long long kStatusAgg = 1ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

constexpr int kCommMaxRanks = kKernelCommMaxRanks;
constexpr int kCommShift = 40;                    // mailbox word: [63:40] epoch, [39:0] count
constexpr unsigned long long kCommMask = (1ull << kCommShift) - 1;

struct KParams {
    const unsigned char* in;
    long long n_owned;
    long long n_total;
    long long num_tiles;        // 512-position warp tiles
    int* out;                   // dense
    int* out_id;                // reduce
    void* out_pos;              // reduce
    unsigned long long out_cap; // reduce: entries out_id / out_pos hold (matches beyond it are counted, not stored)
    long long pos_base;
    unsigned long long* desc;
    unsigned long long* park;   // reduce: per-warp spill rings
    unsigned long long* dbg;    // reduce: 8 mapped host words for the wait watchdog (may be null)
    unsigned long long* total;
    const int32_t* root;
    const uint32_t* pre2;
    const unsigned short* rank2;
    const unsigned char* lut;
    const uint32_t* next2;
    const uint32_t* best2;
    const unsigned short* chk2;
    const uint32_t* hfilt;
    const uint4* hot;
    const uint4* cold;
    const uint4* chains;
    const unsigned char* tails;
    uint32_t hfilt_bytes;       // 0, kHashFilterWords * 4 or (row-indexed filter) twice that (hashed 4-gram first stage, always staged in smem)
    uint32_t chk2_bytes;        // multiple of 16, 0 = no second prefilter stage (always staged in smem)
    uint32_t next2_bytes;       // multiple of 16 (copied to smem when next2_hot; best2 has the same size)
    int has_best2;
    uint32_t hot_buckets;
    uint32_t cold_buckets;
    uint32_t chain_bytes;       // bytes of chain records (copied to smem when chains_hot)
    uint32_t tail_bytes;
    uint32_t mul;
    uint32_t bulk_tiles;        // dense: tiles [0,bulk_tiles) are staged by TMA
    int hot_depth;
    int next2_hot;
    int chains_hot;
    int num_final;
    int code_shift;
    int halo;                   // staged halo, multiple of 16, >= 16
    int in_aligned;             // in is 16-byte aligned
    int out_aligned;            // out is 16-byte aligned
    // cross-GPU count exchange fused into the reduce kernel (comm_world == 0: off)
    unsigned long long* comm_peer[kCommMaxRanks];  // every rank's mailbox (peer-mapped), [comm_rank] = own
    unsigned long long* comm_scan;                 // out: {exclusive offset, total, own count}
    int comm_world, comm_rank;
    unsigned comm_epoch;
};

// tables as the walker sees them (shared-memory copies where available)
struct Tables {
    const int* root;            // smem
    const uint32_t* pre2;       // smem
    const unsigned short* rank2;  // smem
    const unsigned char* lut;   // smem: symbol code | 0x80 (byte in no pattern)
    const uint32_t* next2;      // smem or global
    const uint32_t* best2;      // smem or global; nullptr when no pattern is shorter than K
    const unsigned short* chk2; // smem; nullptr when the second prefilter stage is off
    const uint32_t* hfilt;      // smem; nullptr unless the first stage is the hashed 4-gram filter
    uint32_t hfilt_mask;        // byte-offset mask of the row-indexed filter (FILT 3): words * 4 - 4
    const uint4* hot;           // smem
    const uint4* cold;          // global
    const uint4* chains;        // smem or global
    const unsigned char* tails; // smem or global
    uint32_t hot_buckets, cold_buckets, mul;
    int hot_depth, num_final;
    int code_shift;             // 2-bit alphabets: code = (byte >> code_shift) & 3 (hashed 10-mer first stage)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 1-D TMA bulk copy shared -> global, bulk async-group completion
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_but_newest() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// The one cross-GPU step of the path, done by one warp over NVLink peer memory: every
// rank stores {epoch, its match count} into slot [epoch & 1][rank] of every rank's mailbox, waits
// until its own mailbox holds the current epoch from all ranks, and takes the exclusive prefix.
// One 8-byte word carries flag and value, so relaxed system-scope accesses suffice.  Two parities:
// a rank can be at most one call ahead of the slowest reader of its previous word (it cannot finish
// call e+1 before every rank has published e+1, which each does after its own wait of call e).
__device__ __forceinline__ void comm_exchange_scan(const KParams& p, unsigned long long total, int lane) {
    const unsigned long long ep = static_cast<unsigned long long>(p.comm_epoch & 0xFFFFFFu);
    const int par = static_cast<int>(p.comm_epoch & 1u) * kCommMaxRanks;
    unsigned long long c = 0;
    if (lane < p.comm_world) {
        st_relaxed_sys_u64(p.comm_peer[lane] + par + p.comm_rank, (ep << kCommShift) | total);
        const unsigned long long* mine = p.comm_peer[p.comm_rank] + par + lane;
        unsigned long long v = ld_relaxed_sys_u64(mine);
        while ((v >> kCommShift) != ep) {
            __nanosleep(200);
            v = ld_relaxed_sys_u64(mine);
        }
        c = v & kCommMask;
    }
    __syncwarp();
    unsigned long long incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    const unsigned long long all = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == p.comm_rank) {
        p.comm_scan[0] = incl - c;
        p.comm_scan[1] = all;
        p.comm_scan[2] = c;
    }
}

__device__ __forceinline__ uint32_t home_bucket(uint32_t key, uint32_t mul, uint32_t nb) {
    return __umulhi(key * mul, nb);
}

// hot rows: shared memory.  Returns next state, kChainBit|index, or kEmpty (trap).
__device__ __forceinline__ uint32_t probe_hot(const uint4* tab, uint32_t nb, uint32_t mul, uint32_t key) {
    uint32_t b = home_bucket(key, mul, nb);
    for (;;) {
        const uint4 e = tab[b];
        if (e.x == key) return e.y;
        if (e.z == key) return e.w;
        if (e.z == kEmpty) return kEmpty;
        b = (b + 1 == nb) ? 0 : b + 1;
    }
}
// cold rows: global memory through the read-only path (L1/L2 resident)
__device__ __forceinline__ uint32_t probe_cold(const uint4* __restrict__ tab, uint32_t nb, uint32_t mul,
                                               uint32_t key) {
    uint32_t b = home_bucket(key, mul, nb);
    for (;;) {
        const uint4 e = __ldg(tab + b);
        if (e.x == key) return e.y;
        if (e.z == key) return e.w;
        if (e.z == kEmpty) return kEmpty;
        b = (b + 1 == nb) ? 0 : b + 1;
    }
}

// copy the compiled tables into shared memory (whole CTA), returns the walker's view.
// s_fixed: root 1 KB | pre2 8 KB | rank2 4 KB | lut 256 B;  s_var: [hfilt][chk2][next2][best2][hot buckets][chains][tails]
constexpr int kFixedTableBytes = 1024 + 8192 + 4096 + 256;
__device__ __forceinline__ Tables stage_tables(const KParams& p, unsigned char* s_fixed, unsigned char* s_var,
                                               int tid, int nthreads) {
    int* root = reinterpret_cast<int*>(s_fixed);
    uint4* pre2 = reinterpret_cast<uint4*>(s_fixed + 1024);
    uint4* rank2 = reinterpret_cast<uint4*>(s_fixed + 1024 + 8192);
#pragma unroll 1
    for (int i = tid; i < 256; i += nthreads) root[i] = p.root[i];
#pragma unroll 1
    for (int i = tid; i < 8192 / 16; i += nthreads) pre2[i] = reinterpret_cast<const uint4*>(p.pre2)[i];
#pragma unroll 1
    for (int i = tid; i < 4096 / 16; i += nthreads) rank2[i] = reinterpret_cast<const uint4*>(p.rank2)[i];
    uint4* lut = reinterpret_cast<uint4*>(s_fixed + 1024 + 8192 + 4096);
#pragma unroll 1
    for (int i = tid; i < 256 / 16; i += nthreads) lut[i] = reinterpret_cast<const uint4*>(p.lut)[i];
    Tables t;
    t.root = root;
    t.pre2 = reinterpret_cast<const uint32_t*>(pre2);
    t.rank2 = reinterpret_cast<const unsigned short*>(rank2);
    t.lut = reinterpret_cast<const unsigned char*>(lut);
    t.next2 = p.next2;
    t.best2 = p.has_best2 ? p.best2 : nullptr;
    uint4* var = reinterpret_cast<uint4*>(s_var);
    t.hfilt = nullptr;
    if (p.hfilt_bytes) {
#pragma unroll 1
    #pragma unroll 1
    for (uint32_t i = tid; i < p.hfilt_bytes / 16; i += nthreads)
            var[i] = reinterpret_cast<const uint4*>(p.hfilt)[i];
        t.hfilt = reinterpret_cast<const uint32_t*>(var);
        var += p.hfilt_bytes / 16;
    }
    t.hfilt_mask = p.hfilt_bytes ? p.hfilt_bytes - 4u : 0u;
    t.chk2 = nullptr;
    if (p.chk2_bytes) {
#pragma unroll 1
    #pragma unroll 1
    for (uint32_t i = tid; i < p.chk2_bytes / 16; i += nthreads)
            var[i] = reinterpret_cast<const uint4*>(p.chk2)[i];
        t.chk2 = reinterpret_cast<const unsigned short*>(var);
        var += p.chk2_bytes / 16;
    }
    if (p.next2_hot) {
#pragma unroll 1
    #pragma unroll 1
    for (uint32_t i = tid; i < p.next2_bytes / 16; i += nthreads)
            var[i] = reinterpret_cast<const uint4*>(p.next2)[i];
        t.next2 = reinterpret_cast<const uint32_t*>(var);
        var += p.next2_bytes / 16;
        if (p.has_best2) {
    #pragma unroll 1
    #pragma unroll 1
    for (uint32_t i = tid; i < p.next2_bytes / 16; i += nthreads)
                var[i] = reinterpret_cast<const uint4*>(p.best2)[i];
            t.best2 = reinterpret_cast<const uint32_t*>(var);
            var += p.next2_bytes / 16;
        }
    }
#pragma unroll 1
    for (uint32_t i = tid; i < p.hot_buckets; i += nthreads) var[i] = p.hot[i];
    t.hot = var;
    var += p.hot_buckets;
    t.cold = p.cold;
    t.chains = p.chains;
    t.tails = p.tails;
    if (p.chains_hot) {
        uint4* st = var + p.chain_bytes / 16;
#pragma unroll 1
    #pragma unroll 1
    for (uint32_t i = tid; i < p.chain_bytes / 16; i += nthreads) var[i] = p.chains[i];
#pragma unroll 1
    #pragma unroll 1
    for (uint32_t i = tid; i < p.tail_bytes / 16; i += nthreads)
            st[i] = reinterpret_cast<const uint4*>(p.tails)[i];
        t.chains = var;
        t.tails = reinterpret_cast<const unsigned char*>(st);
    }
    t.hot_buckets = p.hot_buckets;
    t.cold_buckets = p.cold_buckets;
    t.mul = p.mul;
    t.hot_depth = p.hot_depth;
    t.num_final = p.num_final;
    t.code_shift = p.code_shift;
    return t;
}

// 16 consecutive start positions at inb[lb..].  The prefilter index of a position is K = 16/CODE
// symbols of CODE bits each.  CODE == 8: the two text bytes themselves (one LDS.128 + one LDS.32
// of text, one prefilter LDS per position).  CODE == 4 / 2: every byte goes through the symbol
// lut once (code | 0x80 = byte in no pattern), the codes are packed into a bit stream and each
// position's index is a 16-bit window of it; a window holding a byte outside the alphabet cannot
// be indexed and is flagged for the generic path instead.
// Bit q of `cand` = position lb+q survives the prefilter; bit q of `slow` = it needs the generic path.
// second prefilter stage for a position whose first-stage bit is set: the valid K-gram's rank
// selects a 16-bit set of (next byte & 15) values that can still lead somewhere
// what the prefilter reads of the tables (all in shared memory); passed by value to out-of-line code
struct FilterView {
    const uint32_t* pre2;
    const uint32_t* hfilt;
    const unsigned short* rank2;
    const unsigned short* chk2;
    const unsigned char* lut;
    int code_shift;
    uint32_t hfilt_mask;
};
template <typename TT>
__device__ __forceinline__ bool second_stage(const TT& T, uint32_t idx, uint32_t word, uint32_t next_byte) {
    const uint32_t rank = T.rank2[(idx >> 5) & 0x7FFu] + __popc(word & ~(0xFFFFFFFFu >> (idx & 31u)));
    return (T.chk2[rank] >> (next_byte & 15u)) & 1u;
}

// FILT: 0 = exact K-gram set only, 1 = + inline second stage (chk2), 2 / 3 = hashed 4-gram filter
// testing one / two bits (byte alphabets; pfac_table.h) whose survivors the walker re-checks exactly,
// 4 = hashed 10-mer filter (2-bit alphabets), 5 = pair filter: one lookup per two start positions
template <int CODE, int FILT, typename TT>
__device__ __forceinline__ void prefilter16(const unsigned char* inb, int lb, const TT& T, uint32_t& cand,
                                            uint32_t& slow) {
    constexpr int K = 16 / CODE;
    const uint32_t* s_pre2 = T.pre2;
    constexpr bool two_stage = FILT == 1;  // second stage compiled in only for the tables that use it
    cand = 0;
    slow = 0;
    if (CODE == 2 && FILT == 4) {
        // hashed 10-mer first stage (pfac_table.h): 25 text bytes -> 25 two-bit codes, four bytes at a time
        // (shift, mask, one multiply gathers the four 2-bit fields into a byte), then per position a
        // 20-bit window of the code stream, one filter word, two bits of it.  Bytes outside the
        // alphabet alias to some code and bytes past the input are zeros: the filter stays free of false
        // negatives either way (short patterns set all their continuations) and the walker re-checks.
        const uint4 v = *reinterpret_cast<const uint4*>(inb + lb);
        const uint2 u = *reinterpret_cast<const uint2*>(inb + lb + 16);
        const uint32_t w6 = *reinterpret_cast<const uint32_t*>(inb + lb + 24);
        const int sh = T.code_shift;
        auto code4 = [&](uint32_t w) -> uint32_t { return (((w >> sh) & 0x03030303u) * 0x01041040u) >> 24; };
        const uint32_t lo = code4(v.x) | (code4(v.y) << 8) | (code4(v.z) << 16) | (code4(v.w) << 24);  // symbols 0..15
        const uint32_t hi = code4(u.x) | (code4(u.y) << 8) | (code4(w6) << 16);                       // symbols 16..27
#pragma unroll
        for (int q = 15; q >= 0; q--) {
            const uint32_t x = ((q == 0) ? lo : __funnelshift_r(lo, hi, 2 * q)) & 0xFFFFFu;
            const uint32_t hw = *reinterpret_cast<const uint32_t*>(
                reinterpret_cast<const unsigned char*>(T.hfilt) + (((x * kHashFilterMul) >> 17) & static_cast<uint32_t>(kHashFilterWords * 4 - 4)));
            const uint32_t rot = __funnelshift_l(hw, hw, (x * kHashFilterMul2) >> 27) &
                                 __funnelshift_l(hw, hw, (x * kHashFilterMul3) >> 27);
            cand = __funnelshift_l(rot, cand, 1);
        }
        return;
    }
    if (CODE == 8 && FILT == 5) {
        // pair filter (pfac_table.cpp): the start positions q and q+1 share text[q+1..q+3]; the low 24 bits of
        // (those three bytes and the next) * multiplier do not depend on the fourth byte: bits 2..14 are the
        // word's byte offset, bits 19..23 the bit (the rotate takes its amount mod 32).  A pair that passes
        // sets both its positions.  7 instructions per pair.
        uint32_t w[5];
        const uint4 v = *reinterpret_cast<const uint4*>(inb + lb);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        w[4] = *reinterpret_cast<const uint32_t*>(inb + lb + 16);
#pragma unroll
        for (int k = 3; k >= 0; k--) {
#pragma unroll
            for (int j = 3; j >= 1; j -= 2) {   // positions 4k+j-1 and 4k+j: shared bytes start at byte j of word k
                const uint32_t h = __funnelshift_r(w[k], w[k + 1], 8 * j) * kHashFilterMul;
                const uint32_t hw = *reinterpret_cast<const uint32_t*>(
                    reinterpret_cast<const unsigned char*>(T.hfilt) + (h & static_cast<uint32_t>(kHashFilterWords * 4 - 4)));
                const uint32_t rot = __funnelshift_l(hw, hw, h >> 19);
                cand = __funnelshift_l(rot, cand, 1);
                cand = __funnelshift_l(rot, cand, 1);
            }
        }
        return;
    }
    if (CODE == 8) {
        uint32_t w[5];
        const uint4 v = *reinterpret_cast<const uint4*>(inb + lb);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        w[4] = *reinterpret_cast<const uint32_t*>(inb + lb + 16);
#pragma unroll
        for (int k = 3; k >= 0; k--) {
#pragma unroll
            for (int j = 3; j >= 0; j--) {
                const uint32_t x = (j == 0) ? w[k] : __funnelshift_r(w[k], w[k + 1], 8 * j);  // c0 | c1<<8 | ..
                if (FILT >= 2) {
                    // word picked by (c0,c1), bits by all four bytes.  Sparse tables (FILT 2): bits 2..14 of
                    // the product are the word's byte offset; dense tables (FILT 3): row-indexed, word
                    // x & (words - 1).  The multiplies run on the FMA pipe, the ALU pipe is the busy one
                    const uint32_t off = (FILT == 3) ? ((x << 2) & T.hfilt_mask)
                                                     : ((x * kHashFilterMul) & static_cast<uint32_t>(kHashFilterWords * 4 - 4));
                    const uint32_t hw = *reinterpret_cast<const uint32_t*>(
                        reinterpret_cast<const unsigned char*>(T.hfilt) + off);
                    // rotate the bit to the top; dense tables (FILT 3) test a second one
                    uint32_t rot = __funnelshift_l(hw, hw, __umulhi(x, kHashFilterMul2));
                    if (FILT == 3) rot &= __funnelshift_l(hw, hw, __umulhi(x, kHashFilterMul3));
                    cand = __funnelshift_l(rot, cand, 1);
                    continue;
                }
                const uint32_t word = s_pre2[(x >> 5) & 0x7FFu];
                // the table stores bit idx at position 31-(idx&31): one shift brings it to bit 31,
                // one funnel shift moves it into cand from the right
                uint32_t t = word << (x & 31u);
                if (two_stage && static_cast<int>(t) < 0) {
                    // x holds text bytes q..q+3: byte q+2 is the symbol after the 2-gram
                    if (!second_stage(T, x & 0xFFFFu, word, x >> 16)) t = 0;
                }
                cand = __funnelshift_l(t, cand, 1);
            }
        }
    } else {
        constexpr int NBYTES = 16 + K - 1;            // 23 (CODE 2) or 19 (CODE 4)
        constexpr int NWORDS = (NBYTES + 3) / 4;      // 6 or 5
        constexpr int PER = 32 / CODE;                // symbols per 32-bit stream word
        uint32_t w[6];
        const uint4 v = *reinterpret_cast<const uint4*>(inb + lb);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        w[4] = *reinterpret_cast<const uint32_t*>(inb + lb + 16);
        w[5] = (NWORDS > 5) ? *reinterpret_cast<const uint32_t*>(inb + lb + 20) : 0u;
        uint32_t st[3] = {0u, 0u, 0u};                // packed symbol codes
        uint32_t bad = 0;                             // bit i: byte i is outside the alphabet
#pragma unroll
        for (int i = 0; i < NBYTES; i++) {
            const uint32_t byte = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
            const uint32_t e = T.lut[byte];
            st[i / PER] |= (e & ((1u << CODE) - 1u)) << (CODE * (i % PER));
            bad |= (e >> 7) << i;
        }
#pragma unroll
        for (int q = 15; q >= 0; q--) {
            const int wi = (q * CODE) / 32, sh = (q * CODE) % 32;
            const uint32_t x = (sh == 0) ? st[wi] : __funnelshift_r(st[wi], st[wi + 1], sh);
            const uint32_t word = s_pre2[(x >> 5) & 0x7FFu];
            uint32_t t = word << (x & 31u);
            if (two_stage && static_cast<int>(t) < 0) {
                const uint32_t nb = (w[(q + K) >> 2] >> (8 * ((q + K) & 3))) & 0xFFu;  // text byte after the K-gram
                if (!second_stage(T, x & 0xFFFFu, word, nb)) t = 0;
            }
            cand = __funnelshift_l(t, cand, 1);
        }
        if (bad) {  // rare: some window holds a byte that occurs in no pattern
#pragma unroll
            for (int q = 0; q < 16; q++)
                if ((bad >> q) & ((1u << K) - 1u)) slow |= 1u << q;
            cand &= ~slow;
        }
    }
}

// Walk one batch of queued survivors, one per lane, until the batch's longest walk ends (most
// survivors die at their first step, so a batch is usually one trip through the loop).  inb: staged
// text of the tile (stage_bytes bytes, local position 0 = first byte), gin: the same bytes in global
// memory (for walks that run past the staged halo), tile_rem: real input bytes from local position 0
// to the end of the input.  qe: this lane's queue entry (local position | kSlowFlag); returns the id
// of the longest pattern starting there (0 = none, also for inactive lanes).
//
// Per survivor: root row (is c0 alone a match?) -> next2[rank of (c0,c1)] (the walk after two
// bytes, no hashing) -> then per step either a chain (tail compared 4 bytes at a time) or a
// hash probe (hot rows in shared memory below hot_depth, cold rows through L1/L2).
// n (1..4) text bytes from local position `at` when the word is not wholly inside the staged bytes (a walk
// running past the staged halo): byte-wise, staged or from global memory.  Out of line: rare, and
// inlined it sat five times in the walker.
__device__ __noinline__ uint32_t text_word_slow(const unsigned char* inb, int stage_bytes,
                                                const unsigned char* __restrict__ gin, int at, int n) {
    uint32_t x = 0;
    const int nn = n < 4 ? n : 4;
    for (int k = 0; k < nn; k++) {
        const int q = at + k;
        x |= static_cast<uint32_t>((q < stage_bytes) ? inb[q] : gin[q]) << (8 * k);
    }
    return x;
}

template <int CODE, bool HASHED, bool INLINE_SLOW = false>
__device__ __forceinline__ int walk_batch(const Tables& T, const unsigned char* inb, int stage_bytes,
                                          const unsigned char* __restrict__ gin, int tile_rem, bool active,
                                          unsigned qe, int& pl_out) {
    auto text_byte = [&](int at) -> uint32_t { return (at < stage_bytes) ? inb[at] : gin[at]; };
    // n (1..4) text bytes from `at`, little-endian, possibly with junk above byte n-1
    auto text_word = [&](int at, int n) -> uint32_t {
        if (at + 4 <= stage_bytes) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(inb + (at & ~3));
            return __funnelshift_r(w[0], w[1], (at & 3) * 8);
        }
        // past the staged halo: byte-wise, never beyond the bytes that exist
        if (!INLINE_SLOW) return text_word_slow(inb, stage_bytes, gin, at, n);
        uint32_t x = 0;
        const int nn = n < 4 ? n : 4;
        for (int k = 0; k < nn; k++) x |= text_byte(at + k) << (8 * k);
        return x;
    };
    constexpr int K = 16 / CODE;
    int pl = 0, d = 1, limit = 0, best = 0;
    uint32_t v = kEmpty;
    if (active) {
        pl = qe & (kSlowFlag - 1u);
        limit = tile_rem - pl;                           // real input bytes from this position
        // prefilter index again (survivors are few), its rank, and the direct tables
        uint32_t idx;
        bool generic = (qe & kSlowFlag) != 0;
        if (CODE == 8) {
            idx = inb[pl] | (static_cast<uint32_t>(inb[pl + 1]) << 8);  // pl+1 is always staged
        } else {
            idx = 0;
            uint32_t bad = 0;
#pragma unroll
            for (int i = 0; i < K; i++) {
                const uint32_t e = T.lut[inb[pl + i]];
                idx |= (e & ((1u << CODE) - 1u)) << (CODE * i);
                bad |= e;
            }
            // survivors of the hashed first stage were not screened for bytes outside the alphabet or
            // windows cut off by the end of the input: those walk from the root row
            if (HASHED && ((bad & 0x80u) || limit < K)) generic = true;
        }
        if (!generic) {
            const uint32_t word = T.pre2[idx >> 5];
            const uint32_t rank = T.rank2[idx >> 5] + __popc(word & ~(0xFFFFFFFFu >> (idx & 31u)));
            // survivors of the hashed filter may hold any bytes: the exact K-gram bit comes first
            // (unset: c0 starts no pattern at all, or the K-gram leads nowhere and holds no pattern)
            bool viable = !HASHED || static_cast<int>(word << (idx & 31u)) < 0;
            if (CODE == 8) {  // K-1 = 1 symbol: the root row tells whether c0 alone is a pattern
                const int r = viable ? T.root[idx & 0xFFu] : 0;
                best = (r <= T.num_final) ? r : 0;
            } else {
                best = (viable && T.best2) ? static_cast<int>(T.best2[rank]) : 0;  // longest pattern inside K-1 symbols
            }
            // ... then the chk2 set of byte 2 (pl+2 is staged; a byte past the input can only fail
            // a walk that needs it)
            if (HASHED && CODE == 8 && viable && T.chk2) viable = (T.chk2[rank] >> (inb[pl + 2] & 15u)) & 1u;
            v = (viable && limit >= K) ? T.next2[rank] : kEmpty;   // CODE 8: last byte of the input
            d = K - 1;
        } else {
            // generic path: from the root row, hash rows for every further step
            const int r = T.root[inb[pl]];
            v = (r < 0) ? kEmpty : static_cast<uint32_t>(r);
            d = 0;
        }
    }
    // v = the transition out of depth d (d bytes consumed): trap, chain or plain state
    while (__any_sync(0xffffffffu, active)) {
        if (active) {
            bool done = false;
            uint32_t s = 0;
            if (v == kEmpty) {
                done = true;
            } else if (v & kChainBit) {
                // the reference carries the chain's first tail byte: most false candidates end
                // here, before the record (usually an L2 access) is touched
                const bool first_ok =
                    (d + 1 < limit) && text_byte(pl + d + 1) == ((v >> kChainByteShift) & 0xFFu);
                const uint4 rec = first_ok ? T.chains[v & kChainIndexMask]  // {tail offset, len, end|leaf, 4 bytes}
                                           : make_uint4(0u, 0u, 0u, 0u);
                const int len = static_cast<int>(rec.y);
                if (!first_ok || d + 1 + len > limit) {
                    done = true;  // no match, or cut off by the end of the input
                } else {
                    const int at0 = pl + d + 1;
                    {   // first 4 tail bytes travel inside the record: most candidates die here
                        const uint32_t mask = (len >= 4) ? 0xFFFFFFFFu : ((1u << (8 * len)) - 1u);
                        if ((text_word(at0, len) ^ rec.w) & mask) done = true;
                    }
                    // the rest 16 bytes per trip: four independent tail loads in flight at once, and the 16
                    // text bytes from five aligned words shifted by the walk's (constant) misalignment -- one
                    // bounds check and five loads per trip instead of four checks and eight loads
                    const uint32_t* tw = reinterpret_cast<const uint32_t*>(T.tails + rec.x);
                    const int sh = (at0 & 3) * 8;
                    for (int i = 4; i < len && !done; i += 16) {
                        uint32_t t[4], x[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) t[k] = (i + 4 * k < len) ? tw[(i >> 2) + k] : 0u;
                        const int at = at0 + i;
                        if (at + 20 <= stage_bytes) {
                            const uint32_t* w = reinterpret_cast<const uint32_t*>(inb + (at & ~3));
                            const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4];
                            x[0] = __funnelshift_r(a0, a1, sh);
                            x[1] = __funnelshift_r(a1, a2, sh);
                            x[2] = __funnelshift_r(a2, a3, sh);
                            x[3] = __funnelshift_r(a3, a4, sh);
                        } else {  // near or past the end of the staged bytes: word by word, never beyond the bytes that exist
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const int n = len - (i + 4 * k);
                                x[k] = (n > 0) ? text_word(at + 4 * k, n) : 0u;
                            }
                        }
                        uint32_t diff = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int n = len - (i + 4 * k);
                            const uint32_t mask = (n >= 4) ? 0xFFFFFFFFu : (n > 0 ? ((1u << (8 * n)) - 1u) : 0u);
                            diff |= (x[k] ^ t[k]) & mask;
                        }
                        if (diff) done = true;
                    }
                    if (!done) {
                        s = rec.z & ~kChainBit;
                        d += 1 + len;
                        if (static_cast<int>(s) <= T.num_final) best = static_cast<int>(s);
                        if (rec.z & kChainBit) done = true;  // leaf: no out-edges
                    }
                }
            } else {
                s = v & ~kLeafPlainBit;
                if (static_cast<int>(s) <= T.num_final) best = static_cast<int>(s);
                d++;
                if (v & kLeafPlainBit) done = true;  // no out-edges
            }
            if (!done) {
                if (d >= limit) {
                    done = true;
                } else {
                    const uint32_t key = (s << 8) | text_byte(pl + d);
                    // hot rows hold the edges of depth [K, hot_depth); the generic path's
                    // shallow edges and everything deeper are cold
                    v = (d >= K && d < T.hot_depth) ? probe_hot(T.hot, T.hot_buckets, T.mul, key)
                                                    : probe_cold(T.cold, T.cold_buckets, T.mul, key);
                    if (v == kEmpty) done = true;
                }
            }
            if (done) active = false;
        }
    }
    pl_out = pl;
    return best;
}

// ---- helpers shared by the dense and the reduce kernel -------------------------------------------
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// positions of this thread whose K-symbol window would run past the end of the input cannot use
// the prefilter (their padding is not text): route them to the generic path
template <int CODE>
__device__ __forceinline__ void clip_windows(int tile_rem, int lb, uint32_t& cand, uint32_t& slow) {
    constexpr int K = 16 / CODE;
    if (CODE != 8 && tile_rem < kWarpTile + K - 1) {
        int nfast = tile_rem - (K - 1) - lb;  // leading positions of this thread with a full window
        nfast = nfast < 0 ? 0 : (nfast > kPosPerThread ? kPosPerThread : nfast);
        const uint32_t tail = 0xFFFFu & ~((1u << nfast) - 1u);
        slow |= tail;
        cand &= ~tail;
    }
}

// =================================================================================================
// Reduce kernel: fused match + ordered stream compaction, one pass.
//
// Same warp-autonomous pipeline as the dense kernel (persistent CTA per SM, per-warp TMA input
// stages, prefilter, queue, walk), with 31 matcher warps and one scanner warp per CTA.  A warp
// tile is kRedSub = 3 consecutive 512-position blocks (1536 positions, one bulk copy): the walker
// entry, the survivor scan, the tile bookkeeping and the ordering protocol below are paid once per
// 1.5 KB of input instead of once per 512 B (the dense kernel took the tile over from here).
// Ordering:
//   * "CTA tile" c = 31 consecutive warp tiles; CTA b handles c = r*G + b in round r (G = grid);
//     matcher warp w takes warp tile w of it.  The launch is cooperative, so all G CTAs are
//     co-resident and may wait on each other.
//   * walker batches append their (id, position) pairs to the warp's parking as they finish: a small
//     ring in shared memory, and, for what does not fit there (dense matches), the warp's spill ring
//     in global memory (L2-resident; one whole tile always fits an empty one, so no tile is ever
//     walked twice).  At the end of its tile the matcher deposits the tile's match count in a
//     shared-memory ring slot, arrives, and goes on matching; it writes a parked tile's pairs at
//     base + (counts of lower warps) once the scanner has posted the round's base, and never runs
//     more than kLag rounds ahead of the scanner.
//   * the scanner publishes each round's CTA-tile aggregate as soon as all 31 counts are in
//     (one store for the same round's later CTAs, one 64-bit atomic {1 CTA, matches} into the
//     round word) and, independently, resolves bases in order: base(r, b) = matches of rounds < r
//     (round words with every CTA accounted for) + aggregates of CTAs < b in round r, all read
//     with the loads in flight at once.  No CTA publishes a prefix for the others, so rounds are
//     not chained through a single writer.
//   * shared-memory flags (arrived / ready) are block-scope release/acquire operations; the words
//     they guard (counts, base) are plain accesses ordered by them.
//   * every wait loop carries a watchdog (PFAC_SPIN_GUARD): a wait that never ends reports where it
//     stood and traps instead of hanging the GPU.
// (A variant in which the warps of a CTA claim tiles from a shared counter instead of owning a slot
// was measured 8 % slower on C2 and no faster on C4: profiles/r2_history.md.)
// shared memory: [mbarriers][root | pre2 | rank2 | lut][ring][per matcher: queue 512 B | parking ids 512 B |
// parking positions 256 B | records 192 B | 2 input stages][hfilt][chk2][next2][hot][chains][tails]
// =================================================================================================
constexpr int kRedWarps = 32;                  // 31 matcher warps + 1 scanner warp
constexpr int kRedMatchers = kRedWarps - 1;
constexpr int kRedSlots = kRedMatchers;        // warp tiles per CTA tile
constexpr int kRedThreads = kRedWarps * 32;
constexpr int kRedMaxHalo = kDenseMaxHalo;
constexpr int kRedStages = 2;                  // input stages per matcher warp
constexpr int kRedSub = 3;                     // 512-position blocks per warp tile
constexpr int kRedTile = kRedSub * kWarpTile;  // 1536 start positions per warp per round
#ifndef PFAC_QUEUE_CAP
#define PFAC_QUEUE_CAP 256
#endif
#ifndef PFAC_PEND_CAP
#define PFAC_PEND_CAP 128
#endif
constexpr int kQueueCap = PFAC_QUEUE_CAP;      // survivors walked per pass; denser tiles go block by block
#ifndef PFAC_KLAG
#define PFAC_KLAG 6
#endif
constexpr int kLag = PFAC_KLAG;                 // a matcher may run this many rounds ahead of the scanner
// Ring slot reuse: a matcher that passed the lag wait of iteration j has written out every parked
// record of rounds <= j-kLag (bases are posted in round order, and the flush that follows the wait
// writes out everything whose base is there), so after iteration j it holds records > j-kLag only.
// Slot r is rewritten by the first arrival at r+kRing, which needs ready(r+kRing-kLag), i.e. every
// matcher finished iteration r+kRing-kLag-1 and holds records > r+kRing-2*kLag-1 only: kRing >= 2*kLag+1.
constexpr int kRing = (PFAC_KLAG > 7) ? 32 : 16; // arrival ring slots
// Polling intervals.  A poll is ~8 issued instructions and the kernel is bound by instruction issue: at
// 40 / 100 ns a third of all executed instructions were polls; 500 / 1000 ns measured +1 % (C4, C5;
// profiles/r2_history.md).  A matcher in the lag wait is 6 rounds (~30 us) ahead and a base is needed
// ~5 us after its round: waking up a fraction of a microsecond late costs nothing.
#ifndef PFAC_SCANNER_NAP
#define PFAC_SCANNER_NAP 500
#endif
#ifndef PFAC_MATCHER_NAP
#define PFAC_MATCHER_NAP 1000
#endif
constexpr unsigned kScannerNap = PFAC_SCANNER_NAP;   // ns between scanner polls without progress
constexpr unsigned kMatcherNap = PFAC_MATCHER_NAP;   // ns between polls of a matcher waiting for a base
constexpr int kPendCap = PFAC_PEND_CAP;        // matches a warp can park in shared memory while bases are computed
constexpr int kPendRecs = (PFAC_KLAG > 7) ? 16 : 8; // ... spread over at most this many rounds (power of two, > kLag)
constexpr int kPendRecBytes = kPendRecs * 24;  // {round, n_smem, n_spill, tile start (u64)} per record
constexpr int kSpillCap = 2048;                // per-warp spill ring in global memory (entries of 8 bytes)
constexpr int kSlotBytes = 192;                // counts[32] | arrived | ready | base
constexpr int kRingBytes = kRing * kSlotBytes;
constexpr int kRedWarpFixed = kQueueCap * 2 + kPendCap * 6 + kPendRecBytes;
static_assert(kRing >= 2 * kLag + 1 && (kRing & (kRing - 1)) == 0, "ring size");
static_assert((kPendCap & (kPendCap - 1)) == 0 && (kPendRecs & (kPendRecs - 1)) == 0 && kPendRecs > kLag, "parking rings");
static_assert(kRedSub <= 3 && kRedTile <= 2048, "survivor counts are scanned as 10-bit fields; queue entries hold 11-bit positions");
static_assert(kQueueCap >= 8 * kPosPerThread, "a group of 8 lanes of one block must fit the queue");
static_assert(kSpillCap >= kRedTile && (kSpillCap & (kSpillCap - 1)) == 0, "one tile always fits an empty spill ring");
static_assert((kRedWarpFixed & 15) == 0, "input stages are 16-byte aligned (bulk copies)");

// per-round word: [63:40] CTAs that published, [39:0] matches of the round so far
constexpr int kRoundShift = 40;
constexpr unsigned long long kRoundMask = (1ull << kRoundShift) - 1;

struct RingSlot {
    int counts[kRedWarps];
    int arrived;
    int ready;                 // round+1 once base is valid
    unsigned long long base;
    int pad[kSlotBytes / 4 - kRedWarps - 4];
};
static_assert(sizeof(RingSlot) == kSlotBytes, "ring slot layout");

// block-scope release/acquire on shared-memory flags (PTX memory model; the plain accesses to the
// words a flag guards are ordered by it)
__device__ __forceinline__ int ld_acquire_cta_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_s32(int* p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_release_cta_s32(int* p, int v) {
    asm volatile("red.release.cta.shared.add.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// Watchdog of every wait loop of the reduce kernel: a wait that outlasts ~2^24 polls (seconds; a healthy
// wait is microseconds) records where it stood in the handle's mapped host words and traps, so a
// protocol fault surfaces as an error return instead of a GPU that never comes back.
constexpr unsigned long long kSpinLimit = 1ull << 24;
__device__ __noinline__ void pfac_stuck(unsigned long long* dbg, int site, long long a, long long b_, long long c) {
    if (dbg && atomicCAS(dbg, 0ull, 1ull) == 0ull) {   // first reporter only
        dbg[1] = static_cast<unsigned long long>(site);
        dbg[2] = (static_cast<unsigned long long>(blockIdx.x) << 32) | threadIdx.x;
        dbg[3] = static_cast<unsigned long long>(a);
        dbg[4] = static_cast<unsigned long long>(b_);
        dbg[5] = static_cast<unsigned long long>(c);
        __threadfence_system();
    }
    __nanosleep(1000000);   // let the report land before the context dies
    __trap();
}
#define PFAC_SPIN_GUARD(n, dbg, site, a, b_, c) \
    if (++(n) > kSpinLimit) pfac_stuck(dbg, site, (long long)(a), (long long)(b_), (long long)(c));

// inclusive warp scan of a packed word (fields must not overflow into each other)
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// survivors of one 512-position block of a tile: prefilter + the clips for the end of the input
// (windows reaching past it take the generic path) and for positions this shard does not own
template <int CODE, int FILT, typename TT>
__device__ __forceinline__ uint32_t block_survivors(const TT& T, const unsigned char* inb, int blk, int lane,
                                                    int tile_rem, int tile_valid, uint32_t& slow) {
    const int lb = lane * kPosPerThread;
    uint32_t cand;
    prefilter16<CODE, FILT>(inb + blk * kWarpTile, lb, T, cand, slow);
    if (FILT != 4) clip_windows<CODE>(tile_rem - blk * kWarpTile, lb, cand, slow);  // FILT 4: the walker sorts those out
    const int valid = tile_valid - blk * kWarpTile;
    if (valid < kWarpTile) {
        int nv = valid - lb;
        nv = nv < 0 ? 0 : (nv > kPosPerThread ? kPosPerThread : nv);
        cand &= (1u << nv) - 1u;
        slow &= (1u << nv) - 1u;
    }
    return cand;
}
// the same, out of line, for the rare tiles whose survivors are walked block by block: keeps a second
// copy of the prefilter out of the instruction stream of the common path.  Returns cand | slow << 16.
template <int CODE, int FILT>
__device__ __noinline__ uint32_t block_survivors_cold(FilterView fv, const unsigned char* inb, int blk, int lane,
                                                      int tile_rem, int tile_valid) {
    uint32_t slow;
    const uint32_t cand = block_survivors<CODE, FILT>(fv, inb, blk, lane, tile_rem, tile_valid, slow);
    return cand | (slow << 16);
}

template <bool POS64, int CODE, int FILT>
__global__ void __launch_bounds__(kRedThreads, 1) pfac_reduce_kernel(const KParams p) {
    constexpr int NSTAGE = kRedStages;
    constexpr bool HASHED = FILT >= 2;
    extern __shared__ __align__(128) unsigned char smem[];
    const int stage = kRedTile + p.halo;
    const int per_warp = kRedWarpFixed + NSTAGE * stage;
    constexpr int kBarBytes = ((kRedWarps * NSTAGE * 8 + 127) / 128) * 128;
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem);
    unsigned char* s_fixed = smem + kBarBytes;
    RingSlot* ring = reinterpret_cast<RingSlot*>(s_fixed + kFixedTableBytes);
    unsigned char* s_warp = s_fixed + kFixedTableBytes + kRingBytes;
    unsigned char* s_var = s_warp + kRedMatchers * per_warp;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, static_cast<uint32_t>(tid) >> 5, 0);
    const unsigned lt_mask = (1u << lane) - 1u;

    const Tables T = stage_tables(p, s_fixed, s_var, tid, kRedThreads);
    unsigned long long* bar = s_bar + warp * NSTAGE;
    if (lane == 0) {
        for (int i = 0; i < NSTAGE; i++) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#pragma unroll 1
    for (int i = tid; i < kRingBytes / 4; i += kRedThreads) reinterpret_cast<int*>(ring)[i] = 0;
    __syncthreads();  // the only CTA-wide barrier

    const uint32_t G = gridDim.x;
    const uint32_t b = blockIdx.x;
    const uint32_t num_tiles = static_cast<uint32_t>(p.num_tiles);
    const uint32_t num_ctiles = (num_tiles + kRedSlots - 1) / kRedSlots;
    const uint32_t num_rounds = (num_ctiles + G - 1) / G;
    // rounds this CTA takes part in: r with r*G + b < num_ctiles
    const uint32_t my_rounds = (num_ctiles > b) ? (num_ctiles - b + G - 1) / G : 0;
    unsigned long long* g_desc = p.desc;                      // [num_ctiles] aggregate per CTA tile
    unsigned long long* g_rs = p.desc + num_ctiles;           // [num_rounds] {CTAs published, matches} of round r

    if (warp == kRedMatchers) {
        // ======================= scanner warp: publishes aggregates, resolves bases ==================
        // Never blocks on one round: publishing round r (needs only this CTA's counts) is not held up
        // by resolving an earlier round (needs other CTAs' aggregates and the previous round total).
        uint32_t pub_r = 0, res_r = 0;
        unsigned long long idle = 0;       // consecutive polls without progress (watchdog)
        unsigned long long run_total = 0;  // matches of all rounds < res_r - 1 (lane 0)
        auto round_size = [&](uint32_t r) -> uint32_t { return (r + 1 < num_rounds) ? G : (num_ctiles - r * G); };
        while (res_r < my_rounds) {
            bool progress = false;
            if (pub_r < my_rounds) {
                RingSlot* slot = &ring[pub_r & (kRing - 1)];
                if (ld_acquire_cta_s32(&slot->arrived) == kRedMatchers) {
                    int c = (lane < kRedMatchers) ? slot->counts[lane] : 0;
#pragma unroll
                    for (int dd = 16; dd > 0; dd >>= 1) c += __shfl_xor_sync(0xffffffffu, c, dd);
                    if (lane == 0) {
                        const unsigned long long ttotal = static_cast<unsigned long long>(c);
                        st_relaxed_u64(g_desc + pub_r * G + b, kStatusAgg | ttotal);
                        atomicAdd(g_rs + pub_r, (1ull << kRoundShift) | ttotal);
                        // the next arrivals at this slot (round pub_r + kRing) follow a `ready`
                        // released after this store (ring slot reuse, above)
                        slot->arrived = 0;
                    }
                    __syncwarp();
                    pub_r++;
                    progress = true;
                }
            }
            if (res_r < pub_r) {
                const uint32_t r = res_r;
                // needed: aggregates of the same round's earlier CTAs, and the previous round's word
                // with every CTA accounted for (its total then completes the running prefix: no CTA
                // has to publish a prefix for the others, so there is no serial chain between rounds)
                unsigned long long part = 0;
                bool have = true;
                for (uint32_t k = lane; k < b; k += 32) {
                    const unsigned long long v = ld_relaxed_u64(g_desc + r * G + k);
                    have = have && ((v >> 62) != 0);
                    part += v & kValueMask;
                }
                unsigned long long prev = 0;
                if (lane == 0) {
                    prev = run_total;
                    if (r > 0) {
                        const unsigned long long v = ld_relaxed_u64(g_rs + (r - 1));
                        have = have && ((v >> kRoundShift) == round_size(r - 1));
                        prev += v & kRoundMask;
                    }
                }
                if (__all_sync(0xffffffffu, have)) {
#pragma unroll
                    for (int dd = 16; dd > 0; dd >>= 1) part += __shfl_xor_sync(0xffffffffu, part, dd);
                    if (lane == 0) {
                        RingSlot* slot = &ring[r & (kRing - 1)];
                        slot->base = prev + part;
                        st_release_cta_s32(&slot->ready, static_cast<int>(r + 1));
                        run_total = prev;
                    }
                    __syncwarp();
                    res_r++;
                    progress = true;
                }
            }
            if (!progress) {
                __nanosleep(kScannerNap);
                if (lane == 0) { PFAC_SPIN_GUARD(idle, p.dbg, 4, (static_cast<long long>(pub_r) << 32) | res_r, my_rounds, ring[pub_r & (kRing - 1)].arrived) }
            } else {
                idle = 0;
            }
        }
        if (b == 0) {  // CTA 0 takes part in every round: it owns the grand total
            unsigned long long total = 0;
            if (lane == 0) {
                unsigned long long v = ld_relaxed_u64(g_rs + (num_rounds - 1));
                unsigned long long spins = 0;
                while ((v >> kRoundShift) != round_size(num_rounds - 1)) {
                    __nanosleep(kScannerNap);
                    v = ld_relaxed_u64(g_rs + (num_rounds - 1));
                    PFAC_SPIN_GUARD(spins, p.dbg, 5, num_rounds, v >> kRoundShift, v & kRoundMask)
                }
                total = run_total + (v & kRoundMask);
                *p.total = total;
            }
            // multi-GPU: publish the count to every rank over peer memory and scan, in this kernel
            if (p.comm_world > 0) comm_exchange_scan(p, __shfl_sync(0xffffffffu, total, 0), lane);
        }
        return;
    }

    // ============================ matcher warps ==========================================================
    unsigned char* mine = s_warp + warp * per_warp;
    unsigned short* q16 = reinterpret_cast<unsigned short*>(mine);
    int* pend_id = reinterpret_cast<int*>(mine + kQueueCap * 2);                                   // ring of kPendCap
    unsigned short* pend_pos = reinterpret_cast<unsigned short*>(mine + kQueueCap * 2 + kPendCap * 4);  // ring of kPendCap
    struct PendRec { uint32_t round; uint32_t n_smem, n_spill, pad; unsigned long long start; };
    static_assert(sizeof(PendRec) == kPendRecBytes / kPendRecs, "record layout");
    PendRec* recs = reinterpret_cast<PendRec*>(mine + kQueueCap * 2 + kPendCap * 6);  // ring of kPendRecs
    unsigned char* s_in = mine + kRedWarpFixed;
    // this warp's spill ring in global memory (L2-resident): entry = id | position in tile << 32
    unsigned long long* spill = p.park + (static_cast<size_t>(b) * kRedMatchers + warp) * kSpillCap;
    const FilterView fv{T.pre2, T.hfilt, T.rank2, T.chk2, T.lut, T.code_shift, T.hfilt_mask};

    auto issue_load = [&](uint32_t t, int st) {
        if (t < p.bulk_tiles) {
            mbar_arrive_expect_tx(&bar[st], static_cast<uint32_t>(stage));
            tma_load_1d(s_in + st * stage, p.in + static_cast<size_t>(t) * kRedTile, static_cast<uint32_t>(stage),
                        &bar[st]);
        }
    };
    auto tile_of = [&](uint32_t r) -> uint32_t { return (r * G + b) * kRedMatchers + warp; };
    if (elect_one()) {
#pragma unroll
        for (int i = 0; i < NSTAGE; i++) issue_load(tile_of(i), i);
    }

    auto round_ready = [&](uint32_t r) -> bool {
        return ld_acquire_cta_s32(&ring[r & (kRing - 1)].ready) >= static_cast<int>(r + 1);
    };
    auto wait_ready = [&](uint32_t r, int site) {
        unsigned long long spins = 0;
        while (!round_ready(r)) {
            __nanosleep(kMatcherNap);
            PFAC_SPIN_GUARD(spins, p.dbg, site, r, ring[r & (kRing - 1)].ready, warp)
        }
    };
    auto store_pair = [&](unsigned long long at, int id, long long gpos) {
        if (at >= p.out_cap) return;   // caller-stated capacity: counted, not stored
        p.out_id[at] = id;
        if (POS64) reinterpret_cast<long long*>(p.out_pos)[at] = gpos;
        else reinterpret_cast<int*>(p.out_pos)[at] = static_cast<int>(gpos);
    };

    // parking: a tile's matches wait for their round's base in the shared-memory ring; what does not
    // fit there goes on to the warp's spill ring in global memory
    int nrec = 0;         // records in use (closed tiles)
    int rec_head = 0;     // ring index of the oldest record
    int pend_head = 0;    // shared-memory ring: index of the oldest entry, entries in use
    int pend_used = 0;    // (those of the tile being matched included)
    uint32_t spill_head = 0, spill_used = 0;
    auto oldest_round = [&]() -> uint32_t { return recs[rec_head].round; };
    auto pop_oldest = [&]() {  // caller made sure its round is ready
        const PendRec rec = recs[rec_head];
        RingSlot* slot = &ring[rec.round & (kRing - 1)];
        int lower = (static_cast<uint32_t>(lane) < warp) ? slot->counts[lane] : 0;
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) lower += __shfl_xor_sync(0xffffffffu, lower, dd);
        const unsigned long long off = slot->base + static_cast<unsigned long long>(lower);
        const long long gbase = p.pos_base + static_cast<long long>(rec.start);
#pragma unroll 1
        for (uint32_t i = lane; i < rec.n_smem; i += 32) {
            const int e = (pend_head + static_cast<int>(i)) & (kPendCap - 1);
            store_pair(off + i, pend_id[e], gbase + pend_pos[e]);
        }
#pragma unroll 1
        for (uint32_t i = lane; i < rec.n_spill; i += 32) {
            const unsigned long long e = spill[(spill_head + i) & (kSpillCap - 1)];
            store_pair(off + rec.n_smem + i, static_cast<int>(e & 0xFFFFFFFFu), gbase + static_cast<long long>(e >> 32));
        }
        __syncwarp();
        pend_head = (pend_head + static_cast<int>(rec.n_smem)) & (kPendCap - 1);
        pend_used -= static_cast<int>(rec.n_smem);
        spill_head = (spill_head + rec.n_spill) & (kSpillCap - 1);
        spill_used -= rec.n_spill;
        rec_head = (rec_head + 1) & (kPendRecs - 1);
        nrec--;
    };
    // writes out parked tiles whose base has arrived, and keeps going (waiting for bases) while more
    // than keep_recs records stay parked (0: everything out)
    auto flush = [&](int keep_recs) {
        while (nrec > 0) {
            if (!round_ready(oldest_round())) {
                if (nrec <= keep_recs) break;
                wait_ready(oldest_round(), 1);
            }
            pop_oldest();
        }
    };

    uint32_t parity = 0u;
    int st = 0;
    for (uint32_t r = 0; r < my_rounds; ++r) {
        const uint32_t tile = tile_of(r);
        const size_t start = static_cast<size_t>(tile) * kRedTile;
        const unsigned char* inb = s_in + st * stage;
        uint32_t n_smem = 0;     // this tile's matches parked in shared memory ...
        uint32_t n_spill = 0;    // ... and, after those, in the spill ring

        if (tile < num_tiles) {
            if (tile < p.bulk_tiles) {
                unsigned long long spins = 0;
                while (!mbar_try_wait(&bar[st], (parity >> st) & 1u)) { PFAC_SPIN_GUARD(spins, p.dbg, 6, tile, st, r) }
                parity ^= 1u << st;
            } else {
                // odd pointers and tail tiles: guarded copy, zero fill past the end of the input
                unsigned char* w = s_in + st * stage;
#pragma unroll 1
                for (int i = lane; i < stage; i += 32) {
                    const long long g = static_cast<long long>(start) + i;
                    w[i] = (g < p.n_total) ? p.in[g] : static_cast<unsigned char>(0);
                }
                __syncwarp();
            }
            const long long total_left = p.n_total - static_cast<long long>(start);
            const int tile_rem = total_left > 0x7fffffffLL ? 0x7fffffff : static_cast<int>(total_left);
            const long long owned_left = p.n_owned - static_cast<long long>(start);
            const int tile_valid = owned_left > kRedTile ? kRedTile : static_cast<int>(owned_left);

            // ---- survivors of the three blocks, one scan for all three counts (10-bit fields)
            uint32_t cand[kRedSub], slow[kRedSub];
            uint32_t packed = 0;
#ifndef PFAC_UNROLLED_PREFILTER
            // one copy of the prefilter in the instruction stream, run three times: the kernel's instruction
            // footprint decides how well 31 divergent warps share the instruction caches (unrolled three
            // times it measured 3 % slower on C2, 7 % on C4 and C5: profiles/r2_history.md)
            cand[0] = cand[1] = cand[2] = slow[0] = slow[1] = slow[2] = 0u;
#pragma unroll 1
            for (int j = 0; j < kRedSub; j++) {
                uint32_t sl;
                const uint32_t cd = block_survivors<CODE, FILT>(T, inb, j, lane, tile_rem, tile_valid, sl);
                if (j == 0) { cand[0] = cd; slow[0] = sl; }
                else if (j == 1) { cand[1] = cd; slow[1] = sl; }
                else { cand[2] = cd; slow[2] = sl; }
                packed |= static_cast<uint32_t>(__popc(cd | sl)) << (10 * j);
            }
#else
#pragma unroll
            for (int j = 0; j < kRedSub; j++) {
                cand[j] = block_survivors<CODE, FILT>(T, inb, j, lane, tile_rem, tile_valid, slow[j]);
                packed |= static_cast<uint32_t>(__popc(cand[j] | slow[j])) << (10 * j);
            }
#endif
            const uint32_t incl = warp_incl_scan(packed, lane);
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            const int wtotal = static_cast<int>((tot & 1023u) + ((tot >> 10) & 1023u) + (tot >> 20));
            // normally one segment: the whole tile's survivors.  Dense survivors are walked one block at a
            // time, or 8 lanes of a block at a time (<= 128 survivors) when a block alone overflows the
            // queue, the block's prefilter recomputed out of line; one walker call site either way.
            int nseg = 0, seg_total = wtotal;
            if (wtotal > 0 && wtotal <= kQueueCap) {
                nseg = 1;
                const uint32_t excl = incl - packed;
                int qbase = 0;
#pragma unroll
                for (int j = 0; j < kRedSub; j++) {
                    uint32_t all = cand[j] | slow[j];
                    int o = qbase + static_cast<int>((excl >> (10 * j)) & 1023u);
                    const int lb = j * kWarpTile + lane * kPosPerThread;
                    while (all) {
                        const int bit = __ffs(all) - 1;
                        all &= all - 1;
                        q16[o++] = static_cast<unsigned short>((lb + bit) | (((slow[j] >> bit) & 1u) ? kSlowFlag : 0u));
                    }
                    qbase += static_cast<int>((tot >> (10 * j)) & 1023u);
                }
            } else if (wtotal > 0) {
                const bool by_block = (tot & 1023u) <= kQueueCap && ((tot >> 10) & 1023u) <= kQueueCap && (tot >> 20) <= kQueueCap;
                nseg = by_block ? kRedSub : kRedSub * 4;
            }
            for (int seg = 0; seg < nseg; seg++) {
                if (nseg > 1) {
                    const int j = (nseg == kRedSub) ? seg : (seg >> 2);
                    const uint32_t both = block_survivors_cold<CODE, FILT>(fv, inb, j, lane, tile_rem, tile_valid);
                    const uint32_t sl = both >> 16;
                    uint32_t all = (both & 0xFFFFu) | sl;
                    const uint32_t cnt = static_cast<uint32_t>(__popc(all));
                    const uint32_t inc = warp_incl_scan(cnt, lane);
                    uint32_t lo = 0, hi = __shfl_sync(0xffffffffu, inc, 31);
                    if (nseg != kRedSub) {
                        const int g = seg & 3;
                        lo = g ? __shfl_sync(0xffffffffu, inc, 8 * g - 1) : 0u;
                        hi = __shfl_sync(0xffffffffu, inc, 8 * g + 7);
                        if ((lane >> 3) != g) all = 0;
                    }
                    int o = static_cast<int>(inc - cnt - lo);
                    const int lb = j * kWarpTile + lane * kPosPerThread;
                    while (all) {
                        const int bit = __ffs(all) - 1;
                        all &= all - 1;
                        q16[o++] = static_cast<unsigned short>((lb + bit) | (((sl >> bit) & 1u) ? kSlowFlag : 0u));
                    }
                    seg_total = static_cast<int>(hi - lo);
                }
                __syncwarp();
                // ---- walk the queue 32 survivors at a time; park the matches of every batch
                for (int base = 0; base < seg_total; base += 32) {
                    const int qslot = base + lane;
                    const bool active = qslot < seg_total;
                    const unsigned qe = active ? q16[qslot] : 0u;
                    int pl;
                    const int best = walk_batch<CODE, HASHED>(T, inb, stage, p.in + start, tile_rem, active, qe, pl);
                    const unsigned m = __ballot_sync(0xffffffffu, best != 0);
                    if (m == 0) continue;
                    const uint32_t c = static_cast<uint32_t>(__popc(m));
                    const uint32_t mine_at = static_cast<uint32_t>(__popc(m & lt_mask));
                    if (n_spill == 0 && pend_used + static_cast<int>(c) <= kPendCap) {
                        if (best) {
                            const int e = (pend_head + pend_used + static_cast<int>(mine_at)) & (kPendCap - 1);
                            pend_id[e] = best;
                            pend_pos[e] = static_cast<unsigned short>(pl);
                        }
                        pend_used += static_cast<int>(c);
                        n_smem += c;
                    } else {
                        // the rest of this tile goes to the spill ring (order: shared-memory part first)
                        if (spill_used + c > kSpillCap) flush(0);   // older tiles out: this one then fits
                        if (best)
                            spill[(spill_head + spill_used + mine_at) & (kSpillCap - 1)] =
                                static_cast<unsigned long long>(static_cast<uint32_t>(best)) |
                                (static_cast<unsigned long long>(pl) << 32);
                        spill_used += c;
                        n_spill += c;
                    }
                }
                __syncwarp();
            }
            if (elect_one()) issue_load(tile_of(r + NSTAGE), st);  // every lane is done with this stage
            st ^= 1;
        }

        // ---- flow control: nobody runs more than kLag rounds ahead of the scanner, so a ring slot
        // (round r-kRing) is never rewritten while a warp may still read it
        if (r >= kLag) wait_ready(r - kLag, 3);

        // ---- arrive: deposit the count; the scanner warp takes it from here -------------------------
        RingSlot* slot = &ring[r & (kRing - 1)];
        const uint32_t nmatch = n_smem + n_spill;
        if (lane == 0) {
            slot->counts[warp] = static_cast<int>(nmatch);
            red_add_release_cta_s32(&slot->arrived, 1);
        }

        // ---- close this tile's record; write parked tiles whose base has arrived (after the lag wait
        // that is everything of rounds <= r - kLag); the last round waits for everything
        if (nmatch > 0) {
            if (lane == 0) {
                PendRec& rec = recs[(rec_head + nrec) & (kPendRecs - 1)];
                rec.round = r;
                rec.n_smem = n_smem;
                rec.n_spill = n_spill;
                rec.start = static_cast<unsigned long long>(start);
            }
            __syncwarp();
            nrec++;
        }
        flush((r + 1 == my_rounds) ? 0 : kPendRecs - 1);
    }
}

// =================================================================================================
// Dense kernel: one persistent CTA of 32 autonomous warps per SM.
//
// The dense result is zeros but for one int per match, and matches are few, so the result is not
// built in shared memory (round 1 and most of round 2 did: clear a 2 KB slice per 512 positions, patch
// it, bulk-store it).  A warp tile is kDenseSub = 3 consecutive 512-position blocks (1536 positions,
// one bulk load, the survivor scan and the walker entry paid once, as in the reduce kernel).  The
// tile's 6 KB of zeros leave first, then the walker's matches are stored over them, 4 bytes each:
//   * kZeroByTma (every table but the sparse hashed one): ONE bulk store from a zero buffer the whole
//     CTA shares, issued before the tile's input is even waited for.  A patch may only follow once that
//     store has completed (cp.async.bulk.wait_group by the issuing lane: its writes are then visible to
//     that thread; __syncwarp orders the other lanes' stores after it), and nobody waits for a store
//     just issued: a tile's matches are parked in a 64-entry list and written at the start of the next
//     tile, one tile time after their zeros left (wait_group 1: everything but the newest store); only
//     a tile with more matches than the list holds waits for its own zeros.
//   * otherwise (FILT 2 and 5, sparse dictionaries: the 1,000-pattern headline config): twelve st.global.v4 of zeros per lane,
//     then the patches, ordered by __syncwarp.  That kernel is bound by HBM, and with its zeros on the
//     bulk-copy queue it measured 12 % slower (1.03 ms instead of 0.914 ms per GiB; a third of the
//     stall samples sat on the bulk-copy issue slots, loads included; splitting the store, issuing it a
//     tile early, throttling it and separate zero buffers all measured the same), while the kernels
//     bound by issue and walker latency measured 3 % faster with the bulk store than with the plain ones.
// The patches land on lines the zeros have just put into L2, so DRAM still sees 4 bytes per position.
// What the form buys: the walker is entered once per 1536 positions (with the row-indexed filter a batch
// of 32 lanes is two thirds full on the 20,000-pattern dictionary), no per-tile clearing, and 64 KB of
// shared memory for the tables (the 64 KB filter, or chains and tails of a 1,000-pattern dictionary):
// C3 774 -> 946 GB/s (974 with the walker's cheaper chain compare); C2 at the rate of a plain kernel with its
// traffic mix (profiles/r2_history.md).
// shared memory: [mbarriers][root | pre2 | rank2 | lut][zeros 6 KB][per warp: queue 512 B | parked ids 256 B |
// parked positions 128 B | 2 input stages][hfilt][chk2][next2][hot][chains][tails]
// =================================================================================================
constexpr int kDenseSub = kRedSub;
constexpr int kDenseTile = kDenseSub * kWarpTile;    // 1536 start positions per warp per iteration
constexpr int kDenseStages = 2;
constexpr int kDenseZeroBytes = kDenseTile * 4;
constexpr int kDensePend = 64;                       // matches a warp parks until their tile's zeros have landed
constexpr int kDenseWarpFixed = kQueueCap * 2 + kDensePend * 6;
static_assert((kDenseWarpFixed & 15) == 0, "input stages are 16-byte aligned (bulk copies)");
static_assert(kDenseSub == 3, "the survivor scan below packs three 10-bit counts");

template <int CODE, int FILT>
__global__ void __launch_bounds__(kDenseThreads, 1) pfac_dense_kernel(const KParams p) {
    constexpr int NSTAGE = kDenseStages;
    constexpr bool HASHED = FILT >= 2;
    constexpr bool kZeroByTma = FILT != 2 && FILT != 5;
    extern __shared__ __align__(128) unsigned char smem[];
    const int stage = kDenseTile + p.halo;
    const int per_warp = kDenseWarpFixed + NSTAGE * stage;
    constexpr int kBarBytes = ((kDenseWarps * NSTAGE * 8 + 127) / 128) * 128;
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem);
    unsigned char* s_fixed = smem + kBarBytes;
    unsigned char* s_zero = s_fixed + kFixedTableBytes;
    unsigned char* s_warp = s_zero + kDenseZeroBytes;
    unsigned char* s_var = s_warp + kDenseWarps * per_warp;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    // warp-uniform by construction, so addresses derived from it can live in uniform registers
    const uint32_t warp = __shfl_sync(0xffffffffu, static_cast<uint32_t>(tid) >> 5, 0);

    const Tables T = stage_tables(p, s_fixed, s_var, tid, kDenseThreads);
    if (kZeroByTma) {
#pragma unroll 1
        for (int i = tid; i < kDenseZeroBytes / 16; i += kDenseThreads) reinterpret_cast<uint4*>(s_zero)[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();  // the zeros are read by bulk stores (async proxy) only
    }
    unsigned long long* bar = s_bar + warp * NSTAGE;
    if (lane == 0) {
        for (int i = 0; i < NSTAGE; i++) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier

    unsigned char* mine = s_warp + warp * per_warp;
    unsigned short* q16 = reinterpret_cast<unsigned short*>(mine);
    int* pend_id = reinterpret_cast<int*>(mine + kQueueCap * 2);
    unsigned short* pend_pos = reinterpret_cast<unsigned short*>(mine + kQueueCap * 2 + kDensePend * 4);
    unsigned char* s_in = mine + kDenseWarpFixed;
    const FilterView fv{T.pre2, T.hfilt, T.rank2, T.chk2, T.lut, T.code_shift, T.hfilt_mask};
    const unsigned lt_mask = (1u << lane) - 1u;

    // consecutive warps of a CTA take consecutive tiles; tiles advance by the grid
    const uint32_t num_tiles = static_cast<uint32_t>(p.num_tiles);
    const uint32_t full_tiles = static_cast<uint32_t>(p.n_owned / kDenseTile);  // tiles with 1536 owned positions
    const uint32_t tstride = gridDim.x * kDenseWarps;
    uint32_t tile = blockIdx.x * kDenseWarps + warp;

    auto issue_load = [&](uint32_t t, int st) {  // one elected lane; tiles beyond bulk_tiles are copied at use
        if (t < p.bulk_tiles) {
            mbar_arrive_expect_tx(&bar[st], static_cast<uint32_t>(stage));
            tma_load_1d(s_in + st * stage, p.in + static_cast<size_t>(t) * kDenseTile, static_cast<uint32_t>(stage),
                        &bar[st]);
        }
    };
    if (elect_one()) {
#pragma unroll
        for (int i = 0; i < NSTAGE; i++) issue_load(tile + i * tstride, i);
    }

    uint32_t parity = 0u;  // bit s = phase of bar[s]
    int st = 0;
    int npend = 0;         // parked matches, all of the previous tile (of the current one while it is walked)
    auto write_parked = [&](uint32_t t) {  // t: the tile they belong to
        int* o = p.out + static_cast<size_t>(t) * kDenseTile;
#pragma unroll 1
        for (int i = lane; i < npend; i += 32) o[pend_pos[i]] = pend_id[i];
        npend = 0;
    };
    for (; tile < num_tiles; tile += tstride) {
        const unsigned char* inb = s_in + st * stage;
        const size_t start = static_cast<size_t>(tile) * kDenseTile;
        int* gout = p.out + start;
        const long long owned_left = p.n_owned - static_cast<long long>(start);
        const int tile_valid = owned_left > kDenseTile ? kDenseTile : static_cast<int>(owned_left);
        // ---- the tile's zeros leave first
        const bool whole = p.out_aligned && tile < full_tiles;
        const bool bulk_out = kZeroByTma && whole;
        if (bulk_out) {
            if (elect_one()) {
                tma_store_1d(gout, s_zero, kDenseTile * 4);
                tma_store_commit();
            }
        } else if (whole) {
#pragma unroll
            for (int k = 0; k < kDenseTile * 4 / 512; k++) reinterpret_cast<uint4*>(gout)[k * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
        } else {  // odd pointers and the tail tile
#pragma unroll 1
            for (int i = lane; i < tile_valid; i += 32) gout[i] = 0;
        }
        __syncwarp();
        // ---- the previous tile's matches: its zeros left a whole tile ago
        if (kZeroByTma && npend > 0) {  // (parked matches imply that tile's zeros left by a bulk store)
            if (elect_one()) {  // everything but the store just issued (if one was) has been written
                if (bulk_out) tma_store_wait_but_newest();
                else tma_store_wait_all();
            }
            __syncwarp();
            write_parked(tile - tstride);
        }
        if (tile < p.bulk_tiles) {
            mbar_wait(&bar[st], (parity >> st) & 1u);
            parity ^= 1u << st;
        } else {
            // odd pointers and tail tiles: guarded copy, zero fill past the end of the input
            unsigned char* w = s_in + st * stage;
#pragma unroll 1
            for (int i = lane; i < stage; i += 32) {
                const long long g = static_cast<long long>(start) + i;
                w[i] = (g < p.n_total) ? p.in[g] : static_cast<unsigned char>(0);
            }
            __syncwarp();
        }
        const long long total_left = p.n_total - static_cast<long long>(start);
        const int tile_rem = total_left > 0x7fffffffLL ? 0x7fffffff : static_cast<int>(total_left);

        // ---- survivors of the three blocks (one copy of the prefilter, run three times), one scan for
        // all three counts (10-bit fields)
        uint32_t cand[kDenseSub], slow[kDenseSub];
        uint32_t packed = 0;
        cand[0] = cand[1] = cand[2] = slow[0] = slow[1] = slow[2] = 0u;
#pragma unroll 1
        for (int j = 0; j < kDenseSub; j++) {
            uint32_t sl;
            const uint32_t cd = block_survivors<CODE, FILT>(T, inb, j, lane, tile_rem, tile_valid, sl);
            if (j == 0) { cand[0] = cd; slow[0] = sl; }
            else if (j == 1) { cand[1] = cd; slow[1] = sl; }
            else { cand[2] = cd; slow[2] = sl; }
            packed |= static_cast<uint32_t>(__popc(cd | sl)) << (10 * j);
        }
        const uint32_t incl = warp_incl_scan(packed, lane);
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        const int wtotal = static_cast<int>((tot & 1023u) + ((tot >> 10) & 1023u) + (tot >> 20));
        // normally one segment: the whole tile's survivors.  Dense survivors are walked one block at a
        // time, or 8 lanes of a block at a time (<= 128 survivors) when a block alone overflows the
        // queue, the block's prefilter recomputed out of line; one walker call site either way.
        int nseg = 0, seg_total = wtotal;
        if (wtotal > 0 && wtotal <= kQueueCap) {
            nseg = 1;
            const uint32_t excl = incl - packed;
            int qbase = 0;
#pragma unroll
            for (int j = 0; j < kDenseSub; j++) {
                uint32_t all = cand[j] | slow[j];
                int o = qbase + static_cast<int>((excl >> (10 * j)) & 1023u);
                const int lb = j * kWarpTile + lane * kPosPerThread;
                while (all) {
                    const int bit = __ffs(all) - 1;
                    all &= all - 1;
                    q16[o++] = static_cast<unsigned short>((lb + bit) | (((slow[j] >> bit) & 1u) ? kSlowFlag : 0u));
                }
                qbase += static_cast<int>((tot >> (10 * j)) & 1023u);
            }
        } else if (wtotal > 0) {
            const bool by_block = (tot & 1023u) <= kQueueCap && ((tot >> 10) & 1023u) <= kQueueCap && (tot >> 20) <= kQueueCap;
            nseg = by_block ? kDenseSub : kDenseSub * 4;
        }
        bool zeros_landed = !bulk_out;  // plain zero stores are ordered before the patches by the __syncwarp above
        for (int seg = 0; seg < nseg; seg++) {
            if (nseg > 1) {
                __syncwarp();  // the previous segment's queue has been read
                const int j = (nseg == kDenseSub) ? seg : (seg >> 2);
                const uint32_t both = block_survivors_cold<CODE, FILT>(fv, inb, j, lane, tile_rem, tile_valid);
                const uint32_t sl = both >> 16;
                uint32_t all = (both & 0xFFFFu) | sl;
                const uint32_t cnt = static_cast<uint32_t>(__popc(all));
                const uint32_t inc = warp_incl_scan(cnt, lane);
                uint32_t lo = 0, hi = __shfl_sync(0xffffffffu, inc, 31);
                if (nseg != kDenseSub) {
                    const int g = seg & 3;
                    lo = g ? __shfl_sync(0xffffffffu, inc, 8 * g - 1) : 0u;
                    hi = __shfl_sync(0xffffffffu, inc, 8 * g + 7);
                    if ((lane >> 3) != g) all = 0;
                }
                int o = static_cast<int>(inc - cnt - lo);
                const int lb = j * kWarpTile + lane * kPosPerThread;
                while (all) {
                    const int bit = __ffs(all) - 1;
                    all &= all - 1;
                    q16[o++] = static_cast<unsigned short>((lb + bit) | (((sl >> bit) & 1u) ? kSlowFlag : 0u));
                }
                seg_total = static_cast<int>(hi - lo);
            }
            __syncwarp();
            // ---- walk the queue 32 survivors at a time
            for (int base = 0; base < seg_total; base += 32) {
                const int qslot = base + lane;
                const bool active = qslot < seg_total;
                const unsigned qe = active ? q16[qslot] : 0u;
                int pl;
                const int best = walk_batch<CODE, HASHED>(T, inb, stage, p.in + start, tile_rem, active, qe, pl);
                if (!kZeroByTma) {  // zeros and patches are plain stores of this warp, in order
                    if (best) gout[pl] = best;
                    continue;
                }
                const unsigned m = __ballot_sync(0xffffffffu, best != 0);
                if (m == 0) continue;
                const int c = __popc(m);
                if (!zeros_landed && npend + c <= kDensePend) {
                    if (best) {
                        const int e = npend + __popc(m & lt_mask);
                        pend_id[e] = best;
                        pend_pos[e] = static_cast<unsigned short>(pl);
                    }
                    npend += c;
                    continue;
                }
                // more matches than the list holds (or plain zero stores): straight to the result from here on
                if (!zeros_landed) {
                    if (elect_one()) tma_store_wait_all();   // this tile's zeros (and all earlier ones) are written
                    zeros_landed = true;
                }
                __syncwarp();
                write_parked(tile);
                if (best) gout[pl] = best;
            }
        }
        __syncwarp();
        if (elect_one()) issue_load(tile + NSTAGE * tstride, st);  // every lane is done with this stage
        st = (st + 1 == NSTAGE) ? 0 : st + 1;
    }
    if (kZeroByTma) {
        if (elect_one()) tma_store_wait_all();  // shared memory must outlive the bulk stores
        __syncwarp();
        write_parked(tile - tstride);
    }
}

// a rank whose shard is empty still takes part in the exchange
__global__ void pfac_comm_scan_kernel(const KParams p) { comm_exchange_scan(p, 0ull, threadIdx.x & 31); }

// Optional second step: every rank copies its run of (id, position) pairs into ONE
// list that lives on rank dst, at its scanned offset, with plain stores to peer memory (NVLink);
// the last CTA to finish raises this rank's "placed" word in dst's mailbox.
struct PlaceParams {
    const int* ids;
    const long long* pos;
    const unsigned long long* scan;      // {offset, total, count} of this rank (device)
    int* g_ids;                          // list on dst (peer-mapped)
    long long* g_pos;
    unsigned long long capacity;
    unsigned long long* placed;          // dst mailbox word of this rank for this epoch's parity
    unsigned long long* ticket;          // own mailbox scratch word (zero between calls)
    unsigned long long epoch_tag;        // epoch << kCommShift
};
__global__ void __launch_bounds__(256) pfac_place_kernel(const PlaceParams q) {
    const unsigned long long off = q.scan[0], n = q.scan[2];
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
        if (off + i < q.capacity) {
            q.g_ids[off + i] = q.ids[i];
            q.g_pos[off + i] = q.pos[i];
        }
    }
    __threadfence_system();   // this thread's peer stores before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(q.ticket, 1ull);
        if (t + 1 == gridDim.x) {
            *q.ticket = 0;
            __threadfence_system();
            st_release_sys_u64(q.placed, q.epoch_tag | (n & kCommMask));
        }
    }
}
// dst: the list is complete when every rank's "placed" word carries this epoch
__global__ void pfac_wait_placed_kernel(const unsigned long long* placed, int world, unsigned long long epoch) {
    const int lane = threadIdx.x;
    if (lane < world) {
        unsigned long long v = ld_acquire_sys_u64(placed + lane);
        while ((v >> kCommShift) != epoch) {
            __nanosleep(200);
            v = ld_acquire_sys_u64(placed + lane);
        }
    }
}

std::atomic<unsigned long long> g_launches{0};

int roundHalo(int maxPatternLen, int cap) {
    int h = maxPatternLen - 1;
    if (h < 4) h = 4;  // the prefilter reads one word past the tile
    h = (h + 15) & ~15;
    if (h > cap) h = cap;
    return h;
}

size_t denseFixedBytes(int halo) {
    const size_t bar = size_t((kDenseWarps * kDenseStages * 8 + 127) / 128) * 128;
    return bar + kFixedTableBytes + size_t(kDenseZeroBytes) +
           size_t(kDenseWarps) * (kDenseWarpFixed + kDenseStages * (kDenseTile + halo));
}

size_t reduceFixedBytes(int halo) {
    const size_t bar = size_t((kRedWarps * kRedStages * 8 + 127) / 128) * 128;
    return bar + kFixedTableBytes + kRingBytes +
           size_t(kRedMatchers) * (kRedWarpFixed + kRedStages * (kRedTile + halo));
}

size_t tableSmemBytes(const DeviceTable& t) {
    return size_t(t.hfiltBytes) + size_t(t.chk2Bytes) + (t.next2Hot ? size_t(t.next2Bytes) * (t.hasBest2 ? 2 : 1) : 0) +
           size_t(t.hotBuckets) * 16 +
           (t.chainsHot ? size_t(t.chainBytes) + t.tailBytes : 0);
}

KParams baseParams(const DeviceTable& t, const unsigned char* in, size_t n_owned, size_t n_total, int halo,
                   int tileSize) {
    KParams p{};
    p.in = in;
    p.n_owned = (long long)n_owned;
    p.n_total = (long long)n_total;
    p.num_tiles = ((long long)n_owned + tileSize - 1) / tileSize;
    p.root = t.root;
    p.pre2 = t.pre2;
    p.rank2 = t.rank2;
    p.lut = t.lut;
    p.next2 = t.next2;
    p.best2 = t.best2;
    p.chk2 = t.chk2;
    p.chk2_bytes = t.chk2Bytes;
    p.hfilt = t.hfilt;
    p.hfilt_bytes = t.hfiltBytes;
    p.has_best2 = t.hasBest2 ? 1 : 0;
    p.next2_bytes = t.next2Bytes;
    p.next2_hot = t.next2Hot ? 1 : 0;
    p.hot = t.hot;
    p.cold = t.cold;
    p.chains = t.chains;
    p.tails = t.tails;
    p.hot_buckets = t.hotBuckets;
    p.cold_buckets = t.coldBuckets;
    p.chain_bytes = t.chainBytes;
    p.tail_bytes = t.tailBytes;
    p.mul = t.mul;
    p.hot_depth = t.hotDepth;
    p.chains_hot = t.chainsHot ? 1 : 0;
    p.num_final = t.numFinal;
    p.code_shift = t.codeShift;
    p.halo = halo;
    p.in_aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    return p;
}

}  // namespace

size_t tableSmemBudget(int maxPatternLen, bool reduceKernel) {
    const int halo = roundHalo(maxPatternLen, kDenseMaxHalo);
    const size_t fixed = reduceKernel ? reduceFixedBytes(halo) : denseFixedBytes(halo);
    return fixed < size_t(kMaxSmem) ? size_t(kMaxSmem) - fixed : 0;
}

// per CTA tile (46.5 KB of input): one aggregate word; per round: a running total and a counter
size_t reduceWorkspaceWords(size_t n_owned) {
    const size_t tiles = (n_owned + kRedTile - 1) / kRedTile;
    const size_t ctiles = (tiles + kRedSlots - 1) / kRedSlots;
    return 3 * ctiles + 8;
}

// one spill ring per matcher warp of the grid (never initialised: written before it is read)
size_t reduceParkWords(const LaunchConfig& cfg) { return size_t(cfg.numSMs) * kRedMatchers * kSpillCap; }

unsigned long long kernelLaunchCount() { return g_launches.load(); }

namespace {
const void* denseKernelFor(const DeviceTable& t);
const void* reduceKernelFor(const DeviceTable& t, bool pos64);
}  // namespace

// CUDA loads a kernel's code on first use (lazy module loading): asking for the attribute now moves
// that cost (about 2 ms per kernel) from the caller's first match call to the pattern load, where the
// reference pays its own table upload.  One call per table.
cudaError_t prepareKernels(const DeviceTable& dense, const DeviceTable& reduce) {
    const void* ks[3] = {denseKernelFor(dense), reduceKernelFor(reduce, false), reduceKernelFor(reduce, true)};
    for (const void* k : ks) {
        if (!k) return cudaErrorInvalidValue;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launchMatchDense(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                             size_t n_owned, size_t n_total, int* out, cudaStream_t stream) {
    if (n_owned == 0) return cudaSuccess;
    const int halo = roundHalo(t.maxPatternLen, kDenseMaxHalo);
    KParams p = baseParams(t, in, n_owned, n_total, halo, kDenseTile);
    if (p.num_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    p.out = out;
    p.out_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    {   // tile t is staged by TMA iff the pointer is 16-byte aligned and t*1536 + stage <= n_total
        const long long stage = kDenseTile + halo;
        long long bulk = 0;
        if (p.in_aligned && p.n_total >= stage) bulk = (p.n_total - stage) / kDenseTile + 1;
        if (bulk > p.num_tiles) bulk = p.num_tiles;
        p.bulk_tiles = static_cast<uint32_t>(bulk);
    }
    const size_t smem = denseFixedBytes(halo) + tableSmemBytes(t);
    if (smem > size_t(kMaxSmem)) return cudaErrorInvalidConfiguration;
    const void* kernel = denseKernelFor(t);
    if (!kernel) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return e;
    const long long ctaTiles = (p.num_tiles + kDenseWarps - 1) / kDenseWarps;
    long long grid = cfg.numSMs;
    if (grid > ctaTiles) grid = ctaTiles;
    void* args[] = {&p};
    e = cudaLaunchKernel(kernel, dim3(unsigned(grid)), dim3(kDenseThreads), args, smem, stream);
    if (e != cudaSuccess) return e;
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launchMatchReduce(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                              size_t n_owned, size_t n_total, long long pos_base, int* out_id,
                              void* out_pos, bool pos64, unsigned long long* desc, unsigned long long* park,
                              unsigned long long* d_total, cudaStream_t stream, const CommLaunch* comm,
                              unsigned long long* dbg, unsigned long long out_cap) {
    if (n_owned == 0) {
        if (!comm) return cudaSuccess;
        KParams p{};   // nothing to match: the count exchange alone
        for (int i = 0; i < comm->world; i++) p.comm_peer[i] = comm->peer[i];
        p.comm_scan = comm->scan;
        p.comm_world = comm->world;
        p.comm_rank = comm->rank;
        p.comm_epoch = comm->epoch;
        pfac_comm_scan_kernel<<<1, 32, 0, stream>>>(p);
        g_launches++;
        return cudaGetLastError();
    }
    const int halo = roundHalo(t.maxPatternLen, kRedMaxHalo);
    KParams p = baseParams(t, in, n_owned, n_total, halo, kRedTile);
    if (p.num_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    p.out_id = out_id;
    p.out_pos = out_pos;
    p.out_cap = out_cap;
    p.pos_base = pos_base;
    p.desc = desc;
    p.park = park;
    p.dbg = dbg;
    p.total = d_total;
    if (comm) {
        for (int i = 0; i < comm->world; i++) p.comm_peer[i] = comm->peer[i];
        p.comm_scan = comm->scan;
        p.comm_world = comm->world;
        p.comm_rank = comm->rank;
        p.comm_epoch = comm->epoch;
    }
    {
        const long long stage = kRedTile + halo;
        long long bulk = 0;
        if (p.in_aligned && p.n_total >= stage) bulk = (p.n_total - stage) / kRedTile + 1;
        if (bulk > p.num_tiles) bulk = p.num_tiles;
        p.bulk_tiles = static_cast<uint32_t>(bulk);
    }
    const size_t smem = reduceFixedBytes(halo) + tableSmemBytes(t);
    if (smem > size_t(kMaxSmem)) return cudaErrorInvalidConfiguration;
    const void* kernel = reduceKernelFor(t, pos64);
    if (!kernel) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return e;
    const long long ctaTiles = (p.num_tiles + kRedSlots - 1) / kRedSlots;
    long long grid = cfg.numSMs;
    if (grid > ctaTiles) grid = ctaTiles;
    // cooperative launch: every CTA is resident, so waiting on another CTA's aggregate is safe
    void* args[] = {&p};
    e = cudaLaunchCooperativeKernel(kernel, dim3(unsigned(grid)), dim3(kRedThreads), args, smem, stream);
    if (e != cudaSuccess) return e;
    g_launches++;
    return cudaGetLastError();
}

namespace {
// the kernel instantiation a table runs on (nullptr: a table the kernels were not built for)
const void* denseKernelFor(const DeviceTable& t) {
    const void* kernel = nullptr;
    const int filt = t.hfiltBytes ? (t.hfiltK == 2 ? 3 : t.hfiltK == 3 ? 5 : 2) : (t.chk2Bytes ? 1 : 0);
    switch (t.codeBits) {
        case 8:
            if (filt == 5) kernel = (const void*)pfac_dense_kernel<8, 5>;
            else if (filt == 3) kernel = (const void*)pfac_dense_kernel<8, 3>;
            else if (filt == 2) kernel = (const void*)pfac_dense_kernel<8, 2>;
            else if (filt == 1) kernel = (const void*)pfac_dense_kernel<8, 1>;
            else kernel = (const void*)pfac_dense_kernel<8, 0>;
            break;
        case 4: kernel = (const void*)pfac_dense_kernel<4, 0>; break;
        case 2:
            if (t.hfiltBytes) kernel = (const void*)pfac_dense_kernel<2, 4>;
            else kernel = (const void*)pfac_dense_kernel<2, 0>;
            break;
        default: return nullptr;
    }
    if (t.codeBits == 4 && filt) return nullptr;
    if (t.codeBits == 2 && (t.chk2Bytes || (t.hfiltBytes && (t.hfiltK != 2 || t.codeShift < 0)))) return nullptr;
    return kernel;
}
const void* reduceKernelFor(const DeviceTable& t, bool pos64) {
    const void* kernel = nullptr;
    const int filt = t.hfiltBytes ? (t.hfiltK == 2 ? 3 : t.hfiltK == 3 ? 5 : 2) : (t.chk2Bytes ? 1 : 0);
    switch (t.codeBits) {
        case 8:
            if (filt == 5) kernel = pos64 ? (const void*)pfac_reduce_kernel<true, 8, 5> : (const void*)pfac_reduce_kernel<false, 8, 5>;
            else if (filt == 3) kernel = pos64 ? (const void*)pfac_reduce_kernel<true, 8, 3> : (const void*)pfac_reduce_kernel<false, 8, 3>;
            else if (filt == 2) kernel = pos64 ? (const void*)pfac_reduce_kernel<true, 8, 2> : (const void*)pfac_reduce_kernel<false, 8, 2>;
            else if (filt == 1) kernel = pos64 ? (const void*)pfac_reduce_kernel<true, 8, 1> : (const void*)pfac_reduce_kernel<false, 8, 1>;
            else kernel = pos64 ? (const void*)pfac_reduce_kernel<true, 8, 0> : (const void*)pfac_reduce_kernel<false, 8, 0>;
            break;
        case 4: kernel = pos64 ? (const void*)pfac_reduce_kernel<true, 4, 0> : (const void*)pfac_reduce_kernel<false, 4, 0>; break;
        case 2:
            if (t.hfiltBytes) kernel = pos64 ? (const void*)pfac_reduce_kernel<true, 2, 4> : (const void*)pfac_reduce_kernel<false, 2, 4>;
            else kernel = pos64 ? (const void*)pfac_reduce_kernel<true, 2, 0> : (const void*)pfac_reduce_kernel<false, 2, 0>;
            break;
        default: return nullptr;
    }
    if (t.codeBits == 4 && filt) return nullptr;
    if (t.codeBits == 2 && (t.chk2Bytes || (t.hfiltBytes && (t.hfiltK != 2 || t.codeShift < 0)))) return nullptr;
    return kernel;
}
}  // namespace

cudaError_t launchPlaceRun(const int* ids, const long long* pos, const unsigned long long* scan, int* g_ids,
                           long long* g_pos, unsigned long long capacity, unsigned long long* placed,
                           unsigned long long* ticket, unsigned epoch, int numSMs, cudaStream_t stream) {
    PlaceParams q{ids, pos, scan, g_ids, g_pos, capacity, placed, ticket,
                  static_cast<unsigned long long>(epoch & 0xFFFFFFu) << kCommShift};
    pfac_place_kernel<<<numSMs * 4, 256, 0, stream>>>(q);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launchWaitPlaced(const unsigned long long* placed, int world, unsigned epoch, cudaStream_t stream) {
    pfac_wait_placed_kernel<<<1, 32, 0, stream>>>(placed, world, static_cast<unsigned long long>(epoch & 0xFFFFFFu));
    g_launches++;
    return cudaGetLastError();
}

}  // namespace pfac
