// pfac_kernels.cu -- sm_100a kernels of the PFAC matching path.
//
// What is computed is the reference's per-position failureless walk
// (reference PFAC/src/PFAC_CPU.cpp:60-100 is the spec; PFAC_kernel.cu:377-458 and
// PFAC_reduce_kernel.cu:639-867 are the 2012 GPU forms being replaced).  How it is computed
// is new -- see DESIGN.md:
//   * a 64-Kbit two-byte prefilter in shared memory rejects most start positions with one
//     LDS; survivors are compacted into a per-warp queue;
//   * lanes pull survivors from the queue and walk them (refill on early exit): root row and
//     shallow ("hot") hash rows in shared memory, deep ("cold") rows through L1/L2, and
//     path-compressed chains whose tail bytes are compared directly against the text;
//   * dense kernel: one persistent 32-warp CTA per SM, every warp runs its own pipeline --
//     512-byte input tiles (+halo) arrive by 1-D TMA bulk copies on per-warp mbarriers, the
//     warp's 2 KB of results leave by a TMA bulk store; no CTA-wide barrier in the loop, so
//     one long walk delays one warp, not a block;
//   * reduce kernel: ordered warp/CTA compaction + decoupled look-back across tiles writes
//     (id, position) pairs in one pass.
#include "pfac_kernels.h"

#include <atomic>

namespace pfac {

namespace {

constexpr int kPosPerThread = 16;
constexpr int kWarpTile = 32 * kPosPerThread;  // 512 start positions per warp per iteration
constexpr uint32_t kEmpty = 0xFFFFFFFFu;       // empty hash slot / trap
constexpr uint32_t kChainBit = 0x80000000u;
constexpr int kMaxSmem = 232448;               // 227 KB opt-in dynamic shared memory per CTA

// ---- dense kernel geometry -------------------------------------------------------------------
constexpr int kDenseWarps = 32;
constexpr int kDenseThreads = kDenseWarps * 32;
constexpr int kDenseMaxHalo = 512;             // staged halo cap; longer walks read global

// ---- reduce kernel geometry ------------------------------------------------------------------
constexpr int kRedThreads = 256;
constexpr int kRedWarps = kRedThreads / 32;
constexpr int kRedTile = kRedThreads * kPosPerThread;  // 4096
constexpr int kRedMaxHalo = 1024;

// look-back descriptor: [63:62] status, [61:0] value
constexpr unsigned long long kStatusAgg = 1ull << 62;
constexpr unsigned long long kStatusIncl = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

struct KParams {
    const unsigned char* in;
    long long n_owned;
    long long n_total;
    long long num_tiles;        // dense: 512-position warp tiles; reduce: 4096-position CTA tiles
    int* out;                   // dense
    int* out_id;                // reduce
    void* out_pos;              // reduce
    long long pos_base;
    unsigned long long* desc;
    unsigned long long* ticket;  // tile ticket counter (zeroed with desc)
    unsigned long long* total;
    const int32_t* root;
    const uint32_t* pre2;
    const unsigned short* rank2;
    const uint32_t* next2;
    const uint4* hot;
    const uint4* cold;
    const uint4* chains;
    const unsigned char* tails;
    uint32_t next2_bytes;       // multiple of 16 (copied to smem when next2_hot)
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
    int halo;                   // staged halo, multiple of 16, >= 16
    int in_aligned;             // in is 16-byte aligned
    int out_aligned;            // out is 16-byte aligned
};

// tables as the walker sees them (shared-memory copies where available)
struct Tables {
    const int* root;            // smem
    const uint32_t* pre2;       // smem
    const unsigned short* rank2;  // smem
    const uint32_t* next2;      // smem or global
    const uint4* hot;           // smem
    const uint4* cold;          // global
    const uint4* chains;        // smem or global
    const unsigned char* tails; // smem or global
    uint32_t hot_buckets, cold_buckets, mul;
    int hot_depth, num_final;
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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
// s_fixed: root 1 KB | pre2 8 KB | rank2 4 KB;  s_var: [next2][hot buckets][chains][tails]
constexpr int kFixedTableBytes = 1024 + 8192 + 4096;
__device__ __forceinline__ Tables stage_tables(const KParams& p, unsigned char* s_fixed, unsigned char* s_var,
                                               int tid, int nthreads) {
    int* root = reinterpret_cast<int*>(s_fixed);
    uint4* pre2 = reinterpret_cast<uint4*>(s_fixed + 1024);
    uint4* rank2 = reinterpret_cast<uint4*>(s_fixed + 1024 + 8192);
    for (int i = tid; i < 256; i += nthreads) root[i] = p.root[i];
    for (int i = tid; i < 8192 / 16; i += nthreads) pre2[i] = reinterpret_cast<const uint4*>(p.pre2)[i];
    for (int i = tid; i < 4096 / 16; i += nthreads) rank2[i] = reinterpret_cast<const uint4*>(p.rank2)[i];
    Tables t;
    t.root = root;
    t.pre2 = reinterpret_cast<const uint32_t*>(pre2);
    t.rank2 = reinterpret_cast<const unsigned short*>(rank2);
    t.next2 = p.next2;
    uint4* var = reinterpret_cast<uint4*>(s_var);
    if (p.next2_hot) {
        for (uint32_t i = tid; i < p.next2_bytes / 16; i += nthreads)
            var[i] = reinterpret_cast<const uint4*>(p.next2)[i];
        t.next2 = reinterpret_cast<const uint32_t*>(var);
        var += p.next2_bytes / 16;
    }
    for (uint32_t i = tid; i < p.hot_buckets; i += nthreads) var[i] = p.hot[i];
    t.hot = var;
    var += p.hot_buckets;
    t.cold = p.cold;
    t.chains = p.chains;
    t.tails = p.tails;
    if (p.chains_hot) {
        uint4* st = var + p.chain_bytes / 16;
        for (uint32_t i = tid; i < p.chain_bytes / 16; i += nthreads) var[i] = p.chains[i];
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
    return t;
}

// 16 consecutive start positions at inb[lb..]: one LDS.128 + one LDS.32 of text, one prefilter
// LDS per position.  Bit q of the result = position lb+q survives.
__device__ __forceinline__ uint32_t prefilter16(const unsigned char* inb, int lb, const uint32_t* s_pre2) {
    uint32_t w[5];
    const uint4 v = *reinterpret_cast<const uint4*>(inb + lb);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    w[4] = *reinterpret_cast<const uint32_t*>(inb + lb + 16);
    uint32_t cand = 0;
#pragma unroll
    for (int k = 3; k >= 0; k--) {
#pragma unroll
        for (int j = 3; j >= 0; j--) {
            const uint32_t x = (j == 0) ? w[k] : __funnelshift_r(w[k], w[k + 1], 8 * j);  // c0 | c1<<8 | ..
            const uint32_t word = s_pre2[(x >> 5) & 0x7FFu];
            // the table stores bit idx at position 31-(idx&31): one shift brings it to bit 31,
            // one funnel shift moves it into cand from the right
            const uint32_t t = word << (x & 31u);
            cand = __funnelshift_l(t, cand, 1);
        }
    }
    return cand;
}

// exclusive offset of this lane's survivors in the warp queue + warp total; pushes positions
__device__ __forceinline__ int push_survivors(uint32_t cand, int lb, unsigned short* q16, int lane) {
    const int cnt = __popc(cand);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
    int off = incl - cnt;
    while (cand) {
        const int b = __ffs(cand) - 1;
        cand &= cand - 1;
        q16[off++] = static_cast<unsigned short>(lb + b);
    }
    return wtotal;
}

// Walk the queued survivors of one warp.  inb: staged text of the tile (stage_bytes bytes,
// local position 0 = first byte), gin: the same bytes in global memory (for walks that run past
// the staged halo), tile_rem: real input bytes from local position 0 to the end of the input.
// DENSE: wres[local position] = id (non-zero only); else wres[queue slot] = id.
//
// Per survivor: root row (is c0 alone a match?) -> next2[rank of (c0,c1)] (the walk after two
// bytes, no hashing) -> then per step either a chain (tail compared 4 bytes at a time) or a
// hash probe (hot rows in shared memory below hot_depth, cold rows through L1/L2).
template <bool DENSE>
__device__ __forceinline__ void walk_queue(const Tables& T, const unsigned char* inb, int stage_bytes,
                                           const unsigned char* __restrict__ gin, int tile_rem,
                                           const unsigned short* q16, int wtotal, int* wres, int lane) {
    const unsigned lt_mask = (1u << lane) - 1u;
    int head = 0;
    bool active = false;
    int pl = 0, d = 0, limit = 0, best = 0, slot = 0;
    uint32_t v = kEmpty;
    auto text_byte = [&](int at) -> uint32_t { return (at < stage_bytes) ? inb[at] : gin[at]; };
    // n (1..4) text bytes from `at`, little-endian, possibly with junk above byte n-1
    auto text_word = [&](int at, int n) -> uint32_t {
        if (at + 4 <= stage_bytes) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(inb + (at & ~3));
            return __funnelshift_r(w[0], w[1], (at & 3) * 8);
        }
        uint32_t x = 0;
        for (int k = 0; k < n; k++) x |= text_byte(at + k) << (8 * k);
        return x;
    };
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !active);
        if (need) {
            const int my = head + __popc(need & lt_mask);
            if (!active && my < wtotal) {
                slot = my;
                pl = q16[my];
                const uint32_t c0 = inb[pl], c1 = inb[pl + 1];  // pl+1 is always staged (halo >= 16)
                const int r = T.root[c0];                        // valid: the prefilter bit is set
                best = (r <= T.num_final) ? r : 0;
                limit = tile_rem - pl;                           // real input bytes from this position
                v = kEmpty;
                if (limit >= 2) {
                    const uint32_t idx = c0 | (c1 << 8);
                    const uint32_t word = T.pre2[idx >> 5];
                    const uint32_t before = __popc(word & ~(0xFFFFFFFFu >> (idx & 31u)));
                    v = T.next2[T.rank2[idx >> 5] + before];
                }
                d = 1;
                active = true;
            }
            head += __popc(need);
        }
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
            // v = the transition out of depth d (d bytes consumed): trap, chain or plain state
            bool done = false;
            uint32_t s = 0;
            if (v == kEmpty) {
                done = true;
            } else if (v & kChainBit) {
                const uint4 rec = T.chains[v & ~kChainBit];  // {tail offset, len, end|leaf, first 4 bytes}
                const int len = static_cast<int>(rec.y);
                if (d + 1 + len > limit) {
                    done = true;  // cut off by the end of the input: nothing more to report
                } else {
                    const int at0 = pl + d + 1;
                    uint32_t tw = rec.w;
                    int i = 0;
                    for (;;) {
                        const int n = len - i;
                        const uint32_t mask = (n >= 4) ? 0xFFFFFFFFu : ((1u << (8 * n)) - 1u);
                        if ((text_word(at0 + i, n) ^ tw) & mask) { done = true; break; }
                        i += 4;
                        if (i >= len) break;
                        tw = *reinterpret_cast<const uint32_t*>(T.tails + rec.x + i);
                    }
                    if (!done) {
                        s = rec.z & ~kChainBit;
                        d += 1 + len;
                        if (static_cast<int>(s) <= T.num_final) best = static_cast<int>(s);
                        if (rec.z & kChainBit) done = true;  // leaf: no out-edges
                    }
                }
            } else {
                s = v;
                if (static_cast<int>(s) <= T.num_final) best = static_cast<int>(s);
                d++;
            }
            if (!done) {
                if (d >= limit) {
                    done = true;
                } else {
                    const uint32_t key = (s << 8) | text_byte(pl + d);
                    v = (d < T.hot_depth) ? probe_hot(T.hot, T.hot_buckets, T.mul, key)
                                          : probe_cold(T.cold, T.cold_buckets, T.mul, key);
                    if (v == kEmpty) done = true;
                }
            }
            if (done) {
                if (DENSE) {
                    if (best) wres[pl] = best;
                } else {
                    wres[slot] = best;
                }
                active = false;
            }
        }
    }
}

// =================================================================================================
// Dense kernel: one persistent CTA of 32 autonomous warps per SM.
// shared memory: [mbarriers 32*NSTAGE*8][root 1K | pre2 8K | rank2 4K][per warp: queue 1K | res 2K |
// NSTAGE input stages][next2][hot buckets][chains][tails]
// =================================================================================================
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

template <int NSTAGE>
__global__ void __launch_bounds__(kDenseThreads, 1) pfac_dense_kernel(const KParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int stage = kWarpTile + p.halo;
    const int per_warp = kWarpTile * 2 + kWarpTile * 4 + NSTAGE * stage;
    constexpr int kBarBytes = ((kDenseWarps * NSTAGE * 8 + 127) / 128) * 128;
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem);
    unsigned char* s_fixed = smem + kBarBytes;
    unsigned char* s_warp = s_fixed + kFixedTableBytes;
    unsigned char* s_var = s_warp + kDenseWarps * per_warp;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    // warp-uniform by construction, so addresses derived from it can live in uniform registers
    const uint32_t warp = __shfl_sync(0xffffffffu, static_cast<uint32_t>(tid) >> 5, 0);

    const Tables T = stage_tables(p, s_fixed, s_var, tid, kDenseThreads);
    unsigned long long* bar = s_bar + warp * NSTAGE;
    if (lane == 0) {
        for (int i = 0; i < NSTAGE; i++) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier

    unsigned char* mine = s_warp + warp * per_warp;
    unsigned short* q16 = reinterpret_cast<unsigned short*>(mine);
    int* wres = reinterpret_cast<int*>(mine + kWarpTile * 2);
    unsigned char* s_in = mine + kWarpTile * 2 + kWarpTile * 4;

    // consecutive warps of a CTA take consecutive 512-byte tiles; tiles advance by the grid
    const uint32_t num_tiles = static_cast<uint32_t>(p.num_tiles);
    const uint32_t full_tiles = static_cast<uint32_t>(p.n_owned / kWarpTile);  // tiles with 512 owned positions
    const uint32_t tstride = gridDim.x * kDenseWarps;
    uint32_t tile = blockIdx.x * kDenseWarps + warp;

    auto issue_load = [&](uint32_t t, int st) {  // one elected lane; tiles beyond bulk_tiles are copied at use
        if (t < p.bulk_tiles) {
            mbar_arrive_expect_tx(&bar[st], static_cast<uint32_t>(stage));
            tma_load_1d(s_in + st * stage, p.in + static_cast<size_t>(t) * kWarpTile, static_cast<uint32_t>(stage),
                        &bar[st]);
        }
    };

    if (elect_one()) {
#pragma unroll
        for (int i = 0; i < NSTAGE; i++) issue_load(tile + i * tstride, i);
    }

    uint32_t parity = 0u;  // bit s = phase of bar[s]
    int st = 0;
    for (; tile < num_tiles; tile += tstride) {
        unsigned char* inb = s_in + st * stage;
        const size_t start = static_cast<size_t>(tile) * kWarpTile;
        if (tile < p.bulk_tiles) {
            mbar_wait(&bar[st], (parity >> st) & 1u);
            parity ^= 1u << st;
        } else {
            // odd pointers and tail tiles: guarded copy, zero fill past the end of the input
            for (int i = lane; i < stage; i += 32) {
                const long long g = static_cast<long long>(start) + i;
                inb[i] = (g < p.n_total) ? p.in[g] : static_cast<unsigned char>(0);
            }
            __syncwarp();
        }
        const long long total_left = p.n_total - static_cast<long long>(start);
        const int tile_rem = total_left > 0x7fffffffLL ? 0x7fffffff : static_cast<int>(total_left);
        const bool full = tile < full_tiles;

        const int lb = lane * kPosPerThread;
        uint32_t cand = prefilter16(inb, lb, T.pre2);
        int valid = kWarpTile;
        if (!full) {  // tail tile: drop positions we do not own
            valid = static_cast<int>(p.n_owned - static_cast<long long>(start));
            int nv = valid - lb;
            nv = nv < 0 ? 0 : (nv > kPosPerThread ? kPosPerThread : nv);
            cand &= (1u << nv) - 1u;
        }
        const int wtotal = push_survivors(cand, lb, q16, lane);

        // the previous bulk store of this warp must have finished reading wres
        if (elect_one()) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; k++) reinterpret_cast<uint4*>(wres)[lane + 32 * k] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();

        walk_queue<true>(T, inb, stage, p.in + start, tile_rem, q16, wtotal, wres, lane);

        int* gout = p.out + start;
        if (p.out_aligned && full) {
            fence_proxy_async();
            __syncwarp();
            if (elect_one()) {
                tma_store_1d(gout, wres, kWarpTile * 4);
                tma_store_commit();
                issue_load(tile + NSTAGE * tstride, st);  // every lane is done with this stage
            }
        } else {
            __syncwarp();
            for (int i = lane; i < valid; i += 32) gout[i] = wres[i];
            __syncwarp();
            if (elect_one()) issue_load(tile + NSTAGE * tstride, st);
        }
        st = (st + 1 == NSTAGE) ? 0 : st + 1;
    }
    if (elect_one()) tma_store_wait_all();  // shared memory must outlive the bulk stores
}

// =================================================================================================
// Reduce kernel: 256-thread CTAs, 4096-position tiles handed out by a global ticket counter.
// shared memory: [bars 16][base 8][misc][root 1K | pre2 8K | rank2 4K][queue 8K][ids 16K]
// [2 input stages][next2][hot][chains][tails]
// =================================================================================================
constexpr int kROffBar = 0;        // 2 x uint64 mbarrier
constexpr int kROffBase = 16;      // uint64 tile base
constexpr int kROffWcount = 32;    // int[8]
constexpr int kROffWoff = 64;      // int[8]
constexpr int kROffTicket = 96;    // long long[2]
constexpr int kROffFixed = 128;    // root | pre2 | rank2
constexpr int kROffQueue = kROffFixed + kFixedTableBytes;
constexpr int kROffIds = kROffQueue + kRedWarps * kWarpTile * 2;
constexpr int kROffIn = kROffIds + kRedTile * 4;

template <bool POS64>
__global__ void __launch_bounds__(kRedThreads) pfac_reduce_kernel(const KParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + kROffBar);
    unsigned long long* s_base = reinterpret_cast<unsigned long long*>(smem + kROffBase);
    int* s_wcount = reinterpret_cast<int*>(smem + kROffWcount);
    int* s_woff = reinterpret_cast<int*>(smem + kROffWoff);
    long long* s_ticket = reinterpret_cast<long long*>(smem + kROffTicket);
    unsigned short* s_queue = reinterpret_cast<unsigned short*>(smem + kROffQueue);
    int* s_ids = reinterpret_cast<int*>(smem + kROffIds);
    const int stage = kRedTile + p.halo;
    unsigned char* s_in = smem + kROffIn;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    const Tables T = stage_tables(p, smem + kROffFixed, smem + kROffIn + 2 * stage, tid, kRedThreads);
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_is_bulk = [&](long long t) -> bool {
        return p.in_aligned && (t * kRedTile + stage <= p.n_total);
    };
    auto load_tile = [&](long long t, int buf) {
        if (t >= p.num_tiles) return;
        unsigned char* dst = s_in + buf * stage;
        const long long start = t * kRedTile;
        if (tile_is_bulk(t)) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&s_bar[buf], static_cast<uint32_t>(stage));
                tma_load_1d(dst, p.in + start, static_cast<uint32_t>(stage), &s_bar[buf]);
            }
        } else {
            for (int i = tid; i < stage; i += kRedThreads) {
                const long long g = start + i;
                dst[i] = (g < p.n_total) ? p.in[g] : static_cast<unsigned char>(0);
            }
        }
    };

    // Tiles are tickets from a global counter, so a tile index is only ever held by a running
    // CTA and the look-back below can never wait on a CTA that is not resident.
    long long tile, next_tile;
    if (tid == 0) {
        s_ticket[0] = static_cast<long long>(atomicAdd(p.ticket, 1ull));
        s_ticket[1] = static_cast<long long>(atomicAdd(p.ticket, 1ull));
    }
    __syncthreads();
    tile = s_ticket[0];
    next_tile = s_ticket[1];
    __syncthreads();
    load_tile(tile, 0);
    load_tile(next_tile, 1);
    __syncthreads();

    uint32_t parity = 0u;
    unsigned short* q16 = s_queue + warp * kWarpTile;
    int* wids = s_ids + warp * kWarpTile;  // ids by queue slot

    for (int it = 0; tile < p.num_tiles; ++it) {
        unsigned long long my_ticket = 0;
        if (tid == 0) my_ticket = atomicAdd(p.ticket, 1ull);  // consumed just before (A)
        const int buf = it & 1;
        const unsigned char* inb = s_in + buf * stage;
        const long long start = tile * kRedTile;
        if (tile_is_bulk(tile)) {
            mbar_wait(&s_bar[buf], (parity >> buf) & 1u);
            parity ^= 1u << buf;
        }
        const long long owned_left = p.n_owned - start;
        const int valid = owned_left < kRedTile ? static_cast<int>(owned_left) : kRedTile;
        const long long total_left = p.n_total - start;
        const int tile_rem = total_left > 0x7fffffffLL ? 0x7fffffff : static_cast<int>(total_left);

        const int lb = tid * kPosPerThread;
        uint32_t cand = prefilter16(inb, lb, T.pre2);
        if (valid < kRedTile) {
            int nv = valid - lb;
            nv = nv < 0 ? 0 : (nv > kPosPerThread ? kPosPerThread : nv);
            cand &= (1u << nv) - 1u;
        }
        const int wtotal = push_survivors(cand, lb, q16, lane);
        __syncwarp();
        walk_queue<false>(T, inb, stage, p.in + start, tile_rem, q16, wtotal, wids, lane);
        __syncwarp();

        // ---- ordered compaction: warp count -> CTA scan -> look-back -> write pairs -----------
        int nmatch = 0;
        for (int base = 0; base < wtotal; base += 32) {
            const int i = base + lane;
            const int id = (i < wtotal) ? wids[i] : 0;
            nmatch += __popc(__ballot_sync(0xffffffffu, id != 0));
        }
        if (lane == 0) s_wcount[warp] = nmatch;
        if (tid == 0) s_ticket[0] = static_cast<long long>(my_ticket);
        __syncthreads();  // (A) all walks done: s_in[buf] is free, counts are published
        const long long future_tile = s_ticket[0];
        load_tile(future_tile, buf);
        if (warp == 0) {
            const int c = (lane < kRedWarps) ? s_wcount[lane] : 0;
            int inc = c;
#pragma unroll
            for (int dd = 1; dd < kRedWarps; dd <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, inc, dd);
                if (lane >= dd) inc += o;
            }
            if (lane < kRedWarps) s_woff[lane] = inc - c;
            const unsigned long long ttotal =
                static_cast<unsigned long long>(__shfl_sync(0xffffffffu, inc, kRedWarps - 1));
            unsigned long long base = 0;
            if (tile > 0) {
                if (lane == 0) st_relaxed_u64(p.desc + tile, kStatusAgg | ttotal);
                long long t = tile - 1;
                for (;;) {
                    const long long idx = t - lane;
                    unsigned long long v = (idx >= 0) ? ld_relaxed_u64(p.desc + idx) : kStatusIncl;
                    while (__any_sync(0xffffffffu, (v >> 62) == 0)) {
                        if ((v >> 62) == 0) v = ld_relaxed_u64(p.desc + idx);
                    }
                    const unsigned incl_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                    const int first = incl_mask ? (__ffs(incl_mask) - 1) : 31;
                    unsigned long long part = (lane <= first) ? (v & kValueMask) : 0ull;
#pragma unroll
                    for (int dd = 16; dd > 0; dd >>= 1) part += __shfl_xor_sync(0xffffffffu, part, dd);
                    base += part;
                    if (incl_mask) break;
                    t -= 32;
                }
            }
            if (lane == 0) {
                st_relaxed_u64(p.desc + tile, kStatusIncl | (base + ttotal));
                *s_base = base;
                if (tile == p.num_tiles - 1) *p.total = base + ttotal;
            }
        }
        __syncthreads();  // (B)
        unsigned long long obase = *s_base + static_cast<unsigned long long>(s_woff[warp]);
        for (int base = 0; base < wtotal; base += 32) {
            const int i = base + lane;
            const int id = (i < wtotal) ? wids[i] : 0;
            const unsigned m = __ballot_sync(0xffffffffu, id != 0);
            if (id != 0) {
                const unsigned long long o = obase + __popc(m & lt_mask);
                const long long gpos = p.pos_base + start + q16[i];
                p.out_id[o] = id;
                if (POS64) reinterpret_cast<long long*>(p.out_pos)[o] = gpos;
                else reinterpret_cast<int*>(p.out_pos)[o] = static_cast<int>(gpos);
            }
            obase += __popc(m);
        }
        tile = next_tile;
        next_tile = future_tile;
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

int denseStages(int halo) { return halo > 256 ? 2 : 3; }

size_t denseFixedBytes(int halo) {
    const int nst = denseStages(halo);
    const size_t bar = size_t((kDenseWarps * nst * 8 + 127) / 128) * 128;
    return bar + kFixedTableBytes +
           size_t(kDenseWarps) * (kWarpTile * 2 + kWarpTile * 4 + nst * (kWarpTile + halo));
}

size_t tableSmemBytes(const DeviceTable& t) {
    return (t.next2Hot ? size_t(t.next2Bytes) : 0) + size_t(t.hotBuckets) * 16 +
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
    p.next2 = t.next2;
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
    p.halo = halo;
    p.in_aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    return p;
}

}  // namespace

size_t tableSmemBudget(int maxPatternLen) {
    const size_t fixed = denseFixedBytes(roundHalo(maxPatternLen, kDenseMaxHalo));
    return fixed < size_t(kMaxSmem) ? size_t(kMaxSmem) - fixed : 0;
}

size_t reduceWorkspaceWords(size_t n_owned) { return (n_owned + kRedTile - 1) / kRedTile + 1; }

unsigned long long kernelLaunchCount() { return g_launches.load(); }

cudaError_t launchMatchDense(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                             size_t n_owned, size_t n_total, int* out, cudaStream_t stream) {
    if (n_owned == 0) return cudaSuccess;
    const int halo = roundHalo(t.maxPatternLen, kDenseMaxHalo);
    KParams p = baseParams(t, in, n_owned, n_total, halo, kWarpTile);
    if (p.num_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;  // 1 TiB per launch
    p.out = out;
    p.out_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    {   // tile t is staged by TMA iff the pointer is 16-byte aligned and t*512 + stage <= n_total
        const long long stage = kWarpTile + halo;
        long long bulk = 0;
        if (p.in_aligned && p.n_total >= stage) bulk = (p.n_total - stage) / kWarpTile + 1;
        if (bulk > p.num_tiles) bulk = p.num_tiles;
        p.bulk_tiles = static_cast<uint32_t>(bulk);
    }
    const size_t smem = denseFixedBytes(halo) + tableSmemBytes(t);
    if (smem > size_t(kMaxSmem)) return cudaErrorInvalidConfiguration;
    const int nst = denseStages(halo);
    auto kernel = (nst == 3) ? pfac_dense_kernel<3> : pfac_dense_kernel<2>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    const long long ctaTiles = (p.num_tiles + kDenseWarps - 1) / kDenseWarps;
    long long grid = cfg.numSMs;
    if (grid > ctaTiles) grid = ctaTiles;
    kernel<<<int(grid), kDenseThreads, smem, stream>>>(p);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launchMatchReduce(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                              size_t n_owned, size_t n_total, long long pos_base, int* out_id,
                              void* out_pos, bool pos64, unsigned long long* desc,
                              unsigned long long* d_total, cudaStream_t stream) {
    if (n_owned == 0) return cudaSuccess;
    const int halo = roundHalo(t.maxPatternLen, kRedMaxHalo);
    KParams p = baseParams(t, in, n_owned, n_total, halo, kRedTile);
    p.out_id = out_id;
    p.out_pos = out_pos;
    p.pos_base = pos_base;
    p.desc = desc;
    p.ticket = desc + p.num_tiles;  // last workspace word
    p.total = d_total;
    const size_t smem = size_t(kROffIn) + 2 * size_t(kRedTile + halo) + tableSmemBytes(t);
    if (smem > size_t(kMaxSmem)) return cudaErrorInvalidConfiguration;
    auto kernel = pos64 ? pfac_reduce_kernel<true> : pfac_reduce_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int perSM = cfg.ctasPerSM;
    if (perSM <= 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kRedThreads, smem);
        if (e != cudaSuccess) return e;
        if (perSM < 1) return cudaErrorLaunchOutOfResources;
    }
    long long grid = static_cast<long long>(cfg.numSMs) * perSM;
    if (grid > p.num_tiles) grid = p.num_tiles;
    kernel<<<int(grid), kRedThreads, smem, stream>>>(p);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace pfac
