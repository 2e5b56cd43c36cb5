// pfac_kernels.cu -- sm_100a kernels of the PFAC matching path.
//
// What is computed is the reference's per-position failureless walk
// (reference PFAC/src/PFAC_CPU.cpp:60-100 is the spec; PFAC_kernel.cu:377-458 and
// PFAC_reduce_kernel.cu:639-867 are the 2012 GPU forms being replaced).  How it is computed
// is new -- see DESIGN.md:
//   * persistent CTAs, one 4096-byte input tile (+ halo) per iteration, staged global->shared
//     by one 1-D TMA bulk copy (cp.async.bulk + mbarrier), double buffered;
//   * a 64-Kbit two-byte prefilter in shared memory rejects most start positions with one
//     LDS; survivors are compacted into a per-warp queue;
//   * lanes pull survivors from the queue and walk them (refill on early exit), root row and
//     shallow ("hot") hash rows in shared memory, deep ("cold") rows through L1/L2;
//   * dense mode: results staged in shared memory, each warp ships its 2 KB with a TMA bulk
//     store; reduce mode: ordered warp/CTA compaction + decoupled look-back across tiles
//     writes (id, position) pairs in one pass.
#include "pfac_kernels.h"

#include <atomic>

namespace pfac {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPosPerThread = 16;
constexpr int kWarpTile = 32 * kPosPerThread;    // 512 positions per warp per iteration
constexpr int kTile = kThreads * kPosPerThread;  // 4096 positions per CTA per iteration
constexpr int kMaxHalo = 1024;                   // staged halo cap; longer walks read global
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

constexpr int kModeDense = 0;
constexpr int kModeReduce = 1;

// look-back descriptor: [63:62] status, [61:0] value
constexpr unsigned long long kStatusAgg = 1ull << 62;
constexpr unsigned long long kStatusIncl = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

// shared memory map (bytes)
constexpr int kOffBar = 0;                         // 2 x uint64 mbarrier
constexpr int kOffBase = 16;                       // uint64 tile base (reduce)
constexpr int kOffWcount = 32;                     // int[8]
constexpr int kOffTicket = 96;                     // long long[2]
constexpr int kOffWoff = 64;                       // int[8]
constexpr int kOffRoot = 128;                      // int[256]
constexpr int kOffPre2 = kOffRoot + 1024;          // uint32[2048]
constexpr int kOffQueue = kOffPre2 + 8192;         // uint16[kWarps][512]
constexpr int kOffRes = kOffQueue + kWarps * kWarpTile * 2;  // int[kTile]: dense results / reduce ids
constexpr int kOffIn = kOffRes + kTile * 4;        // 2 x stage bytes, then hot buckets

struct KParams {
    const unsigned char* in;
    long long n_owned;
    long long n_total;
    long long num_tiles;
    int* out;                   // dense
    int* out_id;                // reduce
    void* out_pos;              // reduce
    long long pos_base;
    unsigned long long* desc;
    unsigned long long* ticket;  // tile ticket counter (zeroed with desc)
    unsigned long long* total;
    const int32_t* root;
    const uint32_t* pre2;
    const uint4* hot;
    const uint4* cold;
    uint32_t hot_buckets;
    uint32_t cold_buckets;
    uint32_t mul;
    int hot_depth;
    int num_final;
    int halo;                   // multiple of 16, >= 16
    int in_aligned;             // in is 16-byte aligned
    int out_aligned;            // out is 16-byte aligned
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

// hot rows: shared memory
__device__ __forceinline__ int probe_hot(const uint4* tab, uint32_t nb, uint32_t mul, uint32_t key) {
    uint32_t b = home_bucket(key, mul, nb);
    for (;;) {
        const uint4 e = tab[b];
        if (e.x == key) return static_cast<int>(e.y);
        if (e.z == key) return static_cast<int>(e.w);
        if (e.z == kEmpty) return -1;
        b = (b + 1 == nb) ? 0 : b + 1;
    }
}
// cold rows: global memory through the read-only path (L1/L2 resident)
__device__ __forceinline__ int probe_cold(const uint4* __restrict__ tab, uint32_t nb, uint32_t mul, uint32_t key) {
    uint32_t b = home_bucket(key, mul, nb);
    for (;;) {
        const uint4 e = __ldg(tab + b);
        if (e.x == key) return static_cast<int>(e.y);
        if (e.z == key) return static_cast<int>(e.w);
        if (e.z == kEmpty) return -1;
        b = (b + 1 == nb) ? 0 : b + 1;
    }
}

template <int MODE, bool POS64>
__global__ void __launch_bounds__(kThreads) pfac_match_kernel(const KParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + kOffBar);
    unsigned long long* s_base = reinterpret_cast<unsigned long long*>(smem + kOffBase);
    int* s_wcount = reinterpret_cast<int*>(smem + kOffWcount);
    int* s_woff = reinterpret_cast<int*>(smem + kOffWoff);
    long long* s_ticket = reinterpret_cast<long long*>(smem + kOffTicket);
    int* s_root = reinterpret_cast<int*>(smem + kOffRoot);
    uint32_t* s_pre2 = reinterpret_cast<uint32_t*>(smem + kOffPre2);
    unsigned short* s_queue = reinterpret_cast<unsigned short*>(smem + kOffQueue);
    int* s_res = reinterpret_cast<int*>(smem + kOffRes);
    const int stage = kTile + p.halo;
    unsigned char* s_in = smem + kOffIn;
    const uint4* s_hot = reinterpret_cast<const uint4*>(smem + kOffIn + 2 * stage);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    // ---- one-time setup: tables -> shared memory, mbarriers ---------------------------------
    for (int i = tid; i < 256; i += kThreads) s_root[i] = p.root[i];
    for (int i = tid; i < 2048 / 4; i += kThreads)
        reinterpret_cast<uint4*>(s_pre2)[i] = reinterpret_cast<const uint4*>(p.pre2)[i];
    for (uint32_t i = tid; i < p.hot_buckets; i += kThreads)
        const_cast<uint4*>(s_hot)[i] = p.hot[i];
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // stage one tile (+halo) into s_in[buf]; TMA when aligned and fully inside the input,
    // else a guarded cooperative copy that zero-fills past n_total (tail tiles, odd pointers)
    auto tile_is_bulk = [&](long long t) -> bool {
        return p.in_aligned && (t * kTile + stage <= p.n_total);
    };
    auto load_tile = [&](long long t, int buf) {
        if (t >= p.num_tiles) return;
        unsigned char* dst = s_in + buf * stage;
        const long long start = t * kTile;
        if (tile_is_bulk(t)) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&s_bar[buf], static_cast<uint32_t>(stage));
                tma_load_1d(dst, p.in + start, static_cast<uint32_t>(stage), &s_bar[buf]);
            }
        } else {
            for (int i = tid; i < stage; i += kThreads) {
                const long long g = start + i;
                dst[i] = (g < p.n_total) ? p.in[g] : static_cast<unsigned char>(0);
            }
        }
    };

    // Tile assignment.  Dense: static striding (tiles are independent).  Reduce: tickets from a
    // global counter, so a tile index is only ever held by a running CTA and the look-back
    // below can never wait on a CTA that is not resident (two concurrent reduce launches on
    // one GPU would otherwise be able to deadlock each other).
    long long tile, next_tile;
    if (MODE == kModeReduce) {
        if (tid == 0) {
            s_ticket[0] = static_cast<long long>(atomicAdd(p.ticket, 1ull));
            s_ticket[1] = static_cast<long long>(atomicAdd(p.ticket, 1ull));
        }
        __syncthreads();
        tile = s_ticket[0];
        next_tile = s_ticket[1];
        __syncthreads();
    } else {
        tile = blockIdx.x;
        next_tile = tile + gridDim.x;
    }
    load_tile(tile, 0);
    load_tile(next_tile, 1);
    __syncthreads();

    uint32_t parity = 0u;  // bit b = phase of s_bar[b]
    unsigned short* q16 = s_queue + warp * kWarpTile;
    int* wres = s_res + warp * kWarpTile;  // dense: this warp's 512 results; reduce: ids by queue slot

    for (int it = 0; tile < p.num_tiles; ++it) {
        long long future_tile = next_tile + gridDim.x;  // dense; reduce overwrites below
        unsigned long long my_ticket = 0;
        if (MODE == kModeReduce) {
            if (tid == 0) my_ticket = atomicAdd(p.ticket, 1ull);  // consumed just before (A)
        }
        const int buf = it & 1;
        const unsigned char* inb = s_in + buf * stage;
        const long long start = tile * kTile;
        if (tile_is_bulk(tile)) {
            mbar_wait(&s_bar[buf], (parity >> buf) & 1u);
            parity ^= 1u << buf;
        }
        const long long owned_left = p.n_owned - start;             // > 0
        const int valid = owned_left < kTile ? static_cast<int>(owned_left) : kTile;
        const long long total_left = p.n_total - start;             // >= owned_left
        const int tile_rem = total_left > 0x7fffffffLL ? 0x7fffffff : static_cast<int>(total_left);

        // ---- prefilter: 16 consecutive start positions per thread ---------------------------
        const int lb = tid * kPosPerThread;
        uint32_t w[5];
        {
            const uint4 v = *reinterpret_cast<const uint4*>(inb + lb);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            w[4] = *reinterpret_cast<const uint32_t*>(inb + lb + 16);
        }
        uint32_t cand = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t x = (j == 0) ? w[k] : __funnelshift_r(w[k], w[k + 1], 8 * j);
                const uint32_t idx = x & 0xFFFFu;  // c0 | c1<<8
                const uint32_t word = s_pre2[idx >> 5];
                cand |= ((word >> (idx & 31u)) & 1u) << (4 * k + j);
            }
        }
        if (valid < kTile) {  // tail tile: drop positions we do not own
            int nv = valid - lb;
            nv = nv < 0 ? 0 : (nv > kPosPerThread ? kPosPerThread : nv);
            cand &= (1u << nv) - 1u;
        }

        // ---- ordered push of survivors into the warp queue -----------------------------------
        const int cnt = __popc(cand);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
        {
            int off = incl - cnt;
            while (cand) {
                const int b = __ffs(cand) - 1;
                cand &= cand - 1;
                q16[off++] = static_cast<unsigned short>(lb + b);
            }
        }

        if (MODE == kModeDense) {
            // this warp's previous bulk store must have finished reading wres
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; k++)
                reinterpret_cast<uint4*>(wres)[lane + 32 * k] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();

        // ---- walk the survivors; a lane that finishes pulls the next queue entry -------------
        {
            int head = 0;
            bool active = false;
            int pl = 0, d = 0, limit = 0, s = 0, best = 0, slot = 0;
            for (;;) {
                const unsigned need = __ballot_sync(0xffffffffu, !active);
                if (need) {
                    const int my = head + __popc(need & lt_mask);
                    if (!active && my < wtotal) {
                        slot = my;
                        pl = q16[my];
                        s = s_root[inb[pl]];
                        best = (s <= p.num_final) ? s : 0;
                        d = 1;
                        limit = tile_rem - pl;  // bytes of real input from this position
                        active = true;
                    }
                    head += __popc(need);
                }
                if (!__any_sync(0xffffffffu, active)) break;
                if (active) {
                    bool done = (d >= limit) || (s < 0);
                    if (!done) {
                        const int at = pl + d;
                        const uint32_t c = (at < stage) ? inb[at] : p.in[start + at];
                        const uint32_t key = (static_cast<uint32_t>(s) << 8) | c;
                        const int nx = (d < p.hot_depth) ? probe_hot(s_hot, p.hot_buckets, p.mul, key)
                                                         : probe_cold(p.cold, p.cold_buckets, p.mul, key);
                        if (nx < 0) {
                            done = true;
                        } else {
                            s = nx;
                            if (s <= p.num_final) best = s;
                            d++;
                        }
                    }
                    if (done) {
                        if (MODE == kModeDense) {
                            if (best) wres[pl - warp * kWarpTile] = best;
                        } else {
                            wres[slot] = best;
                        }
                        active = false;
                    }
                }
            }
        }

        if (MODE == kModeDense) {
            // ---- ship this warp's 512 results -------------------------------------------------
            int vw = valid - warp * kWarpTile;
            vw = vw < 0 ? 0 : (vw > kWarpTile ? kWarpTile : vw);
            int* gout = p.out + start + warp * kWarpTile;
            if (p.out_aligned && vw == kWarpTile) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_1d(gout, wres, kWarpTile * 4);
                    tma_store_commit();
                }
            } else {
                __syncwarp();
                for (int i = lane; i < vw; i += 32) gout[i] = wres[i];
            }
            __syncthreads();  // every warp is done reading s_in[buf]
            load_tile(future_tile, buf);
        } else {
            // ---- ordered compaction: warp count -> CTA scan -> look-back -> write pairs -------
            __syncwarp();
            int nmatch = 0;
            for (int base = 0; base < wtotal; base += 32) {
                const int i = base + lane;
                const int id = (i < wtotal) ? wres[i] : 0;
                nmatch += __popc(__ballot_sync(0xffffffffu, id != 0));
            }
            if (lane == 0) s_wcount[warp] = nmatch;
            if (tid == 0) s_ticket[0] = static_cast<long long>(my_ticket);
            __syncthreads();  // (A) all walks done: s_in[buf] is free, counts are published
            future_tile = s_ticket[0];
            load_tile(future_tile, buf);
            if (warp == 0) {
                const int c = (lane < kWarps) ? s_wcount[lane] : 0;
                int inc = c;
#pragma unroll
                for (int dd = 1; dd < kWarps; dd <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, inc, dd);
                    if (lane >= dd) inc += o;
                }
                if (lane < kWarps) s_woff[lane] = inc - c;
                const unsigned long long ttotal =
                    static_cast<unsigned long long>(__shfl_sync(0xffffffffu, inc, kWarps - 1));
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
                const int id = (i < wtotal) ? wres[i] : 0;
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
        }
        tile = next_tile;
        next_tile = future_tile;
    }
    if (MODE == kModeDense) {
        if (lane == 0) tma_store_wait_all();  // smem must outlive the bulk stores
    }
}

std::atomic<unsigned long long> g_launches{0};

int haloFor(int maxPatternLen) {
    int h = maxPatternLen - 1;
    if (h < 4) h = 4;                 // the prefilter reads one word past the tile
    h = (h + 15) & ~15;
    if (h > kMaxHalo) h = kMaxHalo;
    return h;
}

size_t smemBytes(const DeviceTable& t, int halo) {
    return size_t(kOffIn) + 2 * size_t(kTile + halo) + size_t(t.hotBuckets) * 16;
}

template <typename K>
cudaError_t prepare(K kernel, size_t smem, const LaunchConfig& cfg, long long numTiles, int* gridOut) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int perSM = cfg.ctasPerSM;
    if (perSM <= 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, kThreads, smem);
        if (e != cudaSuccess) return e;
        if (perSM < 1) return cudaErrorLaunchOutOfResources;
    }
    long long g = static_cast<long long>(cfg.numSMs) * perSM;
    if (g > numTiles) g = numTiles;
    *gridOut = int(g);
    return cudaSuccess;
}

KParams baseParams(const DeviceTable& t, const unsigned char* in, size_t n_owned, size_t n_total, int halo) {
    KParams p{};
    p.in = in;
    p.n_owned = (long long)n_owned;
    p.n_total = (long long)n_total;
    p.num_tiles = ((long long)n_owned + kTile - 1) / kTile;
    p.root = t.root;
    p.pre2 = t.pre2;
    p.hot = t.hot;
    p.cold = t.cold;
    p.hot_buckets = t.hotBuckets;
    p.cold_buckets = t.coldBuckets;
    p.mul = t.mul;
    p.hot_depth = t.hotDepth;
    p.num_final = t.numFinal;
    p.halo = halo;
    p.in_aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    return p;
}

}  // namespace

size_t reduceWorkspaceWords(size_t n_owned) { return (n_owned + kTile - 1) / kTile + 1; }

unsigned long long kernelLaunchCount() { return g_launches.load(); }

cudaError_t launchMatchDense(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                             size_t n_owned, size_t n_total, int* out, cudaStream_t stream) {
    if (n_owned == 0) return cudaSuccess;
    const int halo = haloFor(t.maxPatternLen);
    KParams p = baseParams(t, in, n_owned, n_total, halo);
    p.out = out;
    p.out_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const size_t smem = smemBytes(t, halo);
    auto kernel = pfac_match_kernel<kModeDense, false>;
    int grid = 0;
    cudaError_t e = prepare(kernel, smem, cfg, p.num_tiles, &grid);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kThreads, smem, stream>>>(p);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launchMatchReduce(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                              size_t n_owned, size_t n_total, long long pos_base, int* out_id,
                              void* out_pos, bool pos64, unsigned long long* desc,
                              unsigned long long* d_total, cudaStream_t stream) {
    if (n_owned == 0) return cudaSuccess;
    const int halo = haloFor(t.maxPatternLen);
    KParams p = baseParams(t, in, n_owned, n_total, halo);
    p.out_id = out_id;
    p.out_pos = out_pos;
    p.pos_base = pos_base;
    p.desc = desc;
    p.ticket = desc + p.num_tiles;  // last workspace word
    p.total = d_total;
    const size_t smem = smemBytes(t, halo);
    int grid = 0;
    cudaError_t e;
    if (pos64) {
        auto kernel = pfac_match_kernel<kModeReduce, true>;
        e = prepare(kernel, smem, cfg, p.num_tiles, &grid);
        if (e != cudaSuccess) return e;
        kernel<<<grid, kThreads, smem, stream>>>(p);
    } else {
        auto kernel = pfac_match_kernel<kModeReduce, false>;
        e = prepare(kernel, smem, cfg, p.num_tiles, &grid);
        if (e != cudaSuccess) return e;
        kernel<<<grid, kThreads, smem, stream>>>(p);
    }
    g_launches++;
    return cudaGetLastError();
}

}  // namespace pfac
