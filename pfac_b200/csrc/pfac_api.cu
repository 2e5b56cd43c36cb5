// pfac_api.cu -- the C ABI of libpfac.so (include/PFAC.h, include/PFAC_ext.h): handle
// lifecycle, argument checking in the reference's order, table upload, host-buffer pipelines.
//
// Mirrors the behaviour of reference PFAC/src/PFAC.cpp (each entry point cites its lines);
// the implementation is new.  There is no CPU matcher in this library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <emmintrin.h>  // SSE2 streaming stores (x86-64 baseline)

#include "PFAC.h"
#include "PFAC_ext.h"
#include "pfac_kernels.h"
#include "pfac_table.h"

namespace {

static_assert(pfac::kKernelHashFilterMul == pfac::kHashFilterMul &&
                  pfac::kKernelHashFilterMul2 == pfac::kHashFilterMul2 &&
                  pfac::kKernelHashFilterMul3 == pfac::kHashFilterMul3 &&
                  pfac::kKernelHashFilterWords == pfac::kHashFilterWords,
              "kernels and table compiler disagree on the hashed filter");

constexpr size_t kFilenameLen = 256;            // reference PFAC_P.h FILENAME_LEN
constexpr size_t kInt32Limit = size_t(1) << 31;

size_t envBytes(const char* name, size_t dflt, size_t unit) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt * unit;
    char* end = nullptr;
    unsigned long long x = strtoull(v, &end, 10);
    if (end == v) return dflt * unit;
    return size_t(x) * unit;
}

// ---- pageable host buffers -------------------------------------------------------------------
// Callers of the reference pass malloc'ed memory (reference test/simple_example.cpp).  cudaMemcpyAsync
// on pageable memory goes through the driver's small bounce buffers and does not overlap: measured
// 3.3 GB/s through PFAC_matchFromHost on 1 GiB, against 13.2 GB/s with pinned buffers.  So the
// library owns pinned staging buffers and moves user memory to/from them with a few host threads
// while the DMA and the kernel of the neighbouring chunks run.
class CopyPool {
public:
    struct Job {  // one memcpy split over the pool; finish() before the job or its buffers go away
        std::atomic<int> left{0};
    };

    explicit CopyPool(int workers) {
        for (int i = 0; i < workers; i++) threads_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread& t : threads_) t.join();
    }

    // streamDst: write the destination with streaming stores (results going to user memory that no
    // core reads again soon; measured 8.0 -> 9.0 GB/s on the dense path).  Staging copies that the
    // DMA engine reads next are better off in the cache hierarchy (40 vs 35 GB/s on the reduce path).
    void start(Job& job, void* dst, const void* src, size_t bytes, bool streamDst = false) {
        if (!bytes) return;
        const size_t kMinSlice = size_t(1) << 20;
        const size_t parts = std::min(threads_.size() + 1, (bytes + kMinSlice - 1) / kMinSlice);
        const size_t per = (((bytes + parts - 1) / parts) + 63) & ~size_t(63);
        {
            std::lock_guard<std::mutex> lock(m_);
            for (size_t o = 0; o < bytes; o += per) {
                job.left.fetch_add(1, std::memory_order_relaxed);
                q_.push_back(Task{static_cast<char*>(dst) + o, static_cast<const char*>(src) + o,
                                  std::min(per, bytes - o), &job, streamDst});
            }
        }
        cv_.notify_all();
    }

    void finish(Job& job) {  // the calling thread copies too, then waits for the stragglers
        while (job.left.load(std::memory_order_acquire) > 0) {
            Task t{};
            bool have = false;
            {
                std::lock_guard<std::mutex> lock(m_);
                if (!q_.empty()) {
                    t = q_.front();
                    q_.pop_front();
                    have = true;
                }
            }
            if (have) exec(t);
            else std::this_thread::yield();
        }
    }

    // zero `bytes` bytes at dst with streaming stores, split over the pool (src == nullptr marks a fill)
    void startZero(Job& job, void* dst, size_t bytes) {
        if (!bytes) return;
        const size_t kMinSlice = size_t(1) << 20;
        const size_t parts = std::min(threads_.size() + 1, (bytes + kMinSlice - 1) / kMinSlice);
        const size_t per = (((bytes + parts - 1) / parts) + 63) & ~size_t(63);
        {
            std::lock_guard<std::mutex> lock(m_);
            for (size_t o = 0; o < bytes; o += per) {
                job.left.fetch_add(1, std::memory_order_relaxed);
                q_.push_back(Task{static_cast<char*>(dst) + o, nullptr, std::min(per, bytes - o), &job, true});
            }
        }
        cv_.notify_all();
    }

    void copy(void* dst, const void* src, size_t bytes, bool streamDst = false) {
        Job job;
        start(job, dst, src, bytes, streamDst);
        finish(job);
    }

private:
    struct Task {
        char* dst;
        const char* src;
        size_t bytes;
        Job* job;
        bool stream;
    };
    // streaming stores skip the read-for-ownership of the destination lines
    static void streamCopy(char* dst, const char* src, size_t n) {
        if (n < (size_t(64) << 10)) {
            memcpy(dst, src, n);
            return;
        }
        const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
        memcpy(dst, src, head);
        dst += head;
        src += head;
        n -= head;
        const size_t blocks = n / 64;
        for (size_t i = 0; i < blocks; i++, src += 64, dst += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 48), d);
        }
        _mm_sfence();
        memcpy(dst, src, n - blocks * 64);
    }
    static void streamZero(char* dst, size_t n) {
        if (n < (size_t(64) << 10)) {
            memset(dst, 0, n);
            return;
        }
        const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
        memset(dst, 0, head);
        dst += head;
        n -= head;
        const size_t blocks = n / 64;
        const __m128i z = _mm_setzero_si128();
        for (size_t i = 0; i < blocks; i++, dst += 64) {
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst), z);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 16), z);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 32), z);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 48), z);
        }
        _mm_sfence();
        memset(dst, 0, n - blocks * 64);
    }
    static void exec(const Task& t) {
        if (!t.src) {
            if (streamingStores()) streamZero(t.dst, t.bytes);
            else memset(t.dst, 0, t.bytes);
        } else if (t.stream && streamingStores()) streamCopy(t.dst, t.src, t.bytes);
        else memcpy(t.dst, t.src, t.bytes);
        t.job->left.fetch_sub(1, std::memory_order_release);
    }
    static bool streamingStores() {
        static const bool on = envBytes("PFAC_B200_COPY_STREAM", 1, 1) != 0;
        return on;
    }
    void run() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;  // stop_
                t = q_.front();
                q_.pop_front();
            }
            exec(t);
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Task> q_;
    bool stop_ = false;
};

// one pool per process, shared by the handles that stage; joined when the last of them goes
std::shared_ptr<CopyPool> acquireCopyPool() {
    static std::mutex m;
    static std::weak_ptr<CopyPool> weak;
    std::lock_guard<std::mutex> lock(m);
    std::shared_ptr<CopyPool> sp = weak.lock();
    if (!sp) {
        const unsigned hw = std::thread::hardware_concurrency();
        const size_t dflt = std::min<size_t>(8, std::max<size_t>(1, hw / 2));
        const size_t n = std::max<size_t>(1, envBytes("PFAC_B200_COPY_THREADS", dflt, 1));
        sp = std::make_shared<CopyPool>(int(n) - 1);  // the caller is the n-th copier
        weak = sp;
    }
    return sp;
}

bool isPageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

struct HostPipe {  // cached buffers of the matchFromHost* pipelines
    cudaStream_t stream[2] = {nullptr, nullptr};
    unsigned char* d_in[2] = {nullptr, nullptr};
    int* d_out[2] = {nullptr, nullptr};  // dense results, or ids (reduce)
    int* d_pos[2] = {nullptr, nullptr};  // reduce positions
    size_t chunk = 0;                    // owned bytes per chunk
    size_t inCap = 0;
    size_t posBytes = 0;                 // bytes per position in d_pos (0: none, 4: int, 8: long long)
    // pinned staging for pageable user buffers (allocated on first use)
    unsigned char* s_in[3] = {nullptr, nullptr, nullptr};
    int* s_out[2] = {nullptr, nullptr};
    size_t stageChunk = 0;               // owned bytes per staged chunk (<= chunk)
    size_t stageInCap = 0;
    int* s_list[2] = {nullptr, nullptr};  // pinned (ids | positions) of the sparse result path: the kernel writes them
    size_t listCap = 0;                   // entries per slot
    unsigned long long* h_cnt = nullptr;  // pinned: match count of the chunk in each slot (written by the kernel)
    cudaEvent_t evH2D[2] = {nullptr, nullptr}, evDone[2] = {nullptr, nullptr};
    std::shared_ptr<CopyPool> pool;
    size_t lastH2D = 0, lastD2H = 0;      // bytes the last matchFromHost* call moved over PCIe
};

}  // namespace

struct PFAC_context {
    int device = 0;
    pfac::LaunchConfig launch;
    int platform = PFAC_PLATFORM_GPU;
    int textureMode = PFAC_AUTOMATIC;
    int perfMode = PFAC_TIME_DRIVEN;
    bool patternsReady = false;
    cudaStream_t stream = nullptr;  // legacy default stream, as the reference
    char patternFile[kFilenameLen] = {0};

    pfac::Machine machine;
    // one compiled layout per kernel: they differ only in the hot/cold split (smem budgets differ)
    pfac::DeviceLayout layout;        // dense kernel
    pfac::DeviceLayout layoutReduce;  // reduce kernel
    pfac::DeviceTable table;
    pfac::DeviceTable tableReduce;
    std::vector<void*> d_arrays;      // the device slab holding every array of both tables
    unsigned char* d_slab = nullptr;  // (one allocation: one L2 access-policy window covers all of them)
    size_t slabBytes = 0, slabUsed = 0;
    bool l2Window = false;            // a persisting-L2 window over the slab is set on the handle's streams

    std::mutex pipeMu;  // held for a whole matchFromHost* call (host pipeline buffers)
    std::mutex mu;      // guards the reduce workspace (several host threads may share a handle)
    unsigned long long* d_ws = nullptr;
    size_t wsWords = 0;
    unsigned long long* d_park = nullptr;  // reduce kernel: per-warp spill rings
    unsigned long long* d_total = nullptr;
    unsigned long long* h_total = nullptr;  // pinned
    HostPipe pipe;
};

// Cross-GPU state of one rank (PFAC_ext.h "count exchange"): one device block = mailbox (4 KB: two
// parities x 16 count words, two parities x 16 "placed" words, a ticket) followed by this rank's
// region of the global (id, position) list; every rank maps every rank's block (CUDA IPC between
// processes, peer access inside one process).
struct PFAC_comm {
    int rank = 0, world = 1, device = 0;
    unsigned epoch = 0;        // count exchanges issued so far (all ranks call in the same order)
    unsigned placeEpoch = 0;   // run placements issued so far
    unsigned char* block = nullptr;
    size_t blockBytes = 0;
    size_t listCap = 0;        // entries of the list region
    unsigned char* peer[pfac::kKernelCommMaxRanks] = {};
    size_t peerCap[pfac::kKernelCommMaxRanks] = {};   // list capacity of every rank's block (read from its header)
    bool ipcOpened[pfac::kKernelCommMaxRanks] = {};
    unsigned long long* d_scan = nullptr;   // 4 words, used when the caller passes no d_scan
    unsigned long long* h_scan = nullptr;   // pinned, 4 words
};

namespace {

constexpr size_t kCommMailboxBytes = 4096;
constexpr int kCommPlacedWord = 2 * pfac::kKernelCommMaxRanks;   // u64 index of placed[0][0]
constexpr int kCommTicketWord = 4 * pfac::kKernelCommMaxRanks;
constexpr int kCommCapWord = 4 * pfac::kKernelCommMaxRanks + 1;  // this block's list capacity, for the peers

size_t commIdsBytes(size_t cap) { return ((cap * sizeof(int) + 255) / 256) * 256; }
int* commListIds(unsigned char* block) { return reinterpret_cast<int*>(block + kCommMailboxBytes); }
long long* commListPos(unsigned char* block, size_t cap) {
    return reinterpret_cast<long long*>(block + kCommMailboxBytes + commIdsBytes(cap));
}

void freeDeviceTable(PFAC_handle_t h) {
    for (void* p : h->d_arrays) cudaFree(p);
    h->d_arrays.clear();
    h->d_slab = nullptr;
    h->slabBytes = h->slabUsed = 0;
    h->l2Window = false;
    h->table = pfac::DeviceTable();
    h->tableReduce = pfac::DeviceTable();
}

void freePatterns(PFAC_handle_t h) {  // reference PFAC_freeResource, PFAC.cpp:221-296
    freeDeviceTable(h);
    h->machine = pfac::Machine();
    h->layout = pfac::DeviceLayout();
    h->layoutReduce = pfac::DeviceLayout();
    h->patternsReady = false;
}

void freeStage(HostPipe& p) {
    for (int i = 0; i < 3; i++) {
        if (p.s_in[i]) cudaFreeHost(p.s_in[i]);
        p.s_in[i] = nullptr;
    }
    for (int i = 0; i < 2; i++) {
        if (p.s_out[i]) cudaFreeHost(p.s_out[i]);
        p.s_out[i] = nullptr;
        if (p.s_list[i]) cudaFreeHost(p.s_list[i]);
        p.s_list[i] = nullptr;
    }
    if (p.h_cnt) cudaFreeHost(p.h_cnt);
    p.h_cnt = nullptr;
    p.stageChunk = p.stageInCap = p.listCap = 0;
}

void freePipe(HostPipe& p) {
    freeStage(p);
    for (int i = 0; i < 2; i++) {
        if (p.d_in[i]) cudaFree(p.d_in[i]);
        if (p.d_out[i]) cudaFree(p.d_out[i]);
        if (p.d_pos[i]) cudaFree(p.d_pos[i]);
        if (p.stream[i]) cudaStreamDestroy(p.stream[i]);
        if (p.evH2D[i]) cudaEventDestroy(p.evH2D[i]);
        if (p.evDone[i]) cudaEventDestroy(p.evDone[i]);
    }
    p = HostPipe();
}

size_t paddedTableBytes(size_t bytes) { return ((bytes + 15) / 16) * 16 + 256; }  // kernels copy tables in 16-byte pieces

PFAC_status_t uploadArray(PFAC_handle_t h, const void** dst, const void* src, size_t bytes) {
    const size_t padded = paddedTableBytes(bytes);
    if (h->slabUsed + padded > h->slabBytes) return PFAC_STATUS_INTERNAL_ERROR;
    void* d = h->d_slab + h->slabUsed;
    h->slabUsed += (padded + 255) & ~size_t(255);
    if (cudaMemset(d, 0xFF, padded) != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
    if (bytes && cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
        return PFAC_STATUS_INTERNAL_ERROR;
    *dst = d;
    return PFAC_STATUS_SUCCESS;
}

// shared-memory bytes the table compiler may fill: whatever the kernel's per-warp pipelines
// leave free (PFAC_B200_HOT_KB caps it; PFAC_SPACE_DRIVEN = nothing in smem)
size_t hotBudget(PFAC_handle_t h, bool reduceKernel) {
    if (h->perfMode == PFAC_SPACE_DRIVEN) return 0;
    const size_t avail = pfac::tableSmemBudget(h->machine.maxPatternLen, reduceKernel);
    const size_t cap = envBytes("PFAC_B200_HOT_KB", avail / 1024 + 1, 1024);
    return cap < avail ? cap : avail;
}

PFAC_status_t uploadLayout(PFAC_handle_t h, const pfac::DeviceLayout& L, pfac::DeviceTable& t) {
    PFAC_status_t st;
    const void* d = nullptr;
#define PFAC_UP(field, type, src, bytes)                                                   \
    if ((st = uploadArray(h, &d, src, bytes)) != PFAC_STATUS_SUCCESS) return st;           \
    t.field = static_cast<type>(d);
    PFAC_UP(root, const int32_t*, L.root, sizeof(L.root))
    PFAC_UP(pre2, const uint32_t*, L.pre2.data(), L.pre2.size() * 4)
    PFAC_UP(rank2, const unsigned short*, L.rank2.data(), L.rank2.size() * 2)
    PFAC_UP(lut, const unsigned char*, L.lut, sizeof(L.lut))
    PFAC_UP(next2, const uint32_t*, L.next2.data(), L.next2.size() * 4)
    PFAC_UP(best2, const uint32_t*, L.best2.data(), L.best2.size() * 4)
    PFAC_UP(chk2, const unsigned short*, L.chk2.data(), L.chk2.size() * 2)
    PFAC_UP(hfilt, const uint32_t*, L.hfilt.data(), L.hfilt.size() * 4)
    t.hfiltBytes = uint32_t(L.hfilt.size() * 4);
    t.hfiltK = L.hfiltK;
    t.chk2Bytes = uint32_t(((L.chk2.size() * 2 + 15) / 16) * 16);
    t.hasBest2 = !L.best2.empty();
    t.codeBits = L.codeBits;
    t.gramLen = L.gramLen;
    t.codeShift = L.codeShift;
    PFAC_UP(hot, const uint4*, L.hot.data(), L.hot.size() * 4)
    PFAC_UP(cold, const uint4*, L.cold.data(), L.cold.size() * 4)
    PFAC_UP(chains, const uint4*, L.chains.data(), L.chains.size() * 4)
    PFAC_UP(tails, const unsigned char*, L.tails.data(), L.tails.size())
#undef PFAC_UP
    t.next2Bytes = uint32_t(((L.next2.size() * 4 + 15) / 16) * 16);
    t.next2Hot = L.next2Hot;
    t.chainBytes = uint32_t(L.chains.size() * 4);
    t.tailBytes = uint32_t(L.tails.size());
    t.chainsHot = L.chainsHot;
    t.hotBuckets = L.hotBuckets;
    t.coldBuckets = L.coldBuckets;
    t.mul = L.mul;
    t.hotDepth = L.hotDepth;
    t.numFinal = h->machine.numFinal;
    t.maxPatternLen = h->machine.maxPatternLen;
    return PFAC_STATUS_SUCCESS;
}

// PFAC_B200_FILTER=exact|hash overrides the table compiler's choice of first prefilter stage
int filterPolicy() {
    const char* v = getenv("PFAC_B200_FILTER");
    if (v && !strcmp(v, "exact")) return pfac::kFilterExact;
    if (v && !strcmp(v, "hash")) return pfac::kFilterHashed;
    if (v && !strcmp(v, "nopair")) return pfac::kFilterNoPair;
    return pfac::kFilterAuto;
}

// The pair filter halves the first stage's instructions: +18 % for the reduce kernel, which is bound by
// instruction issue (C2: 1,333 -> 1,575 GB/s per call), but its twice as many survivors cost the dense
// kernel, which is bound by HBM, 5-7 % (0.918 -> 0.97 ms per GiB): the dense layout keeps the
// per-position filter unless PFAC_B200_FILTER says otherwise.
int filterPolicyFor(bool reduceKernel) {
    const int p = filterPolicy();
    return (!reduceKernel && p == pfac::kFilterAuto) ? pfac::kFilterNoPair : p;
}

// compile the device layouts for the current perf mode and upload them (reference
// PFAC_bindTable, PFAC.cpp:321-343, which picks the dense 2-D table or the hash table)
PFAC_status_t uploadTables(PFAC_handle_t h) {
    freeDeviceTable(h);
    // one slab for all arrays of both layouts (14 arrays each, 512 B of padding and alignment per array)
    h->slabBytes = h->layout.deviceBytes() + h->layoutReduce.deviceBytes() + 2 * 14 * 768 + 4096;
    if (cudaMalloc(reinterpret_cast<void**>(&h->d_slab), h->slabBytes) != cudaSuccess) {
        h->d_slab = nullptr;
        h->slabBytes = 0;
        return PFAC_STATUS_CUDA_ALLOC_FAILED;
    }
    h->d_arrays.push_back(h->d_slab);
    PFAC_status_t st = uploadLayout(h, h->layout, h->table);
    if (st != PFAC_STATUS_SUCCESS) return st;
    st = uploadLayout(h, h->layoutReduce, h->tableReduce);
    if (st != PFAC_STATUS_SUCCESS) return st;
    const cudaError_t e = pfac::prepareKernels(h->table, h->tableReduce);
    return e == cudaSuccess ? PFAC_STATUS_SUCCESS : (e == cudaErrorMemoryAllocation ? PFAC_STATUS_CUDA_ALLOC_FAILED : PFAC_STATUS_INTERNAL_ERROR);
}

// Tables that do not fit shared memory are read through L2 by the walkers while gigabytes of text and
// results stream through the same cache.  PFAC_B200_L2_PERSIST=1 asks for a persisting access-policy
// window over the table slab on `stream` (and, once per device, for a persisting carve-out to hold
// it); measured on C3 in profiles/r2_history.md.  Off by default: it changes a device-wide limit.
void applyL2Window(PFAC_handle_t h, cudaStream_t stream) {
    static const bool want = envBytes("PFAC_B200_L2_PERSIST", 0, 1) != 0;
    if (!want || !h->d_slab || h->slabUsed < (size_t(256) << 10)) return;   // small tables live in shared memory
    int maxWin = 0, maxPersist = 0;
    cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
    cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
    if (maxWin <= 0 || maxPersist <= 0) return;
    const size_t bytes = std::min<size_t>(h->slabUsed, size_t(maxWin));
    size_t cur = 0;
    cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
    const size_t need = std::min<size_t>(((bytes + (size_t(1) << 20) - 1) >> 20) << 20, size_t(maxPersist));
    if (cur < need) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, need);
    cudaStreamAttrValue v{};
    v.accessPolicyWindow.base_ptr = h->d_slab;
    v.accessPolicyWindow.num_bytes = bytes;
    v.accessPolicyWindow.hitRatio = 1.0f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
}

PFAC_status_t bindTable(PFAC_handle_t h) {
    // the two layouts are independent functions of the (read-only) automaton: compile them side by side.
    // Large dictionaries allocate by the number of states: an allocation failure must come back as
    // PFAC_STATUS_ALLOC_FAILED (as the reference does), not unwind through the C ABI or a std::thread.
    const int policy = filterPolicyFor(true);
    const size_t budgetReduce = hotBudget(h, true);
    std::atomic<bool> failed{false};
    auto reduceSide = [&] {
        try {
            pfac::compileLayout(h->machine, budgetReduce, h->layoutReduce, policy);
        } catch (...) {
            failed = true;
        }
    };
    std::thread other;
    try {
        other = std::thread(reduceSide);
    } catch (...) {  // no thread to be had: one after the other
    }
    try {
        pfac::compileLayout(h->machine, hotBudget(h, false), h->layout, filterPolicyFor(false));
    } catch (...) {
        failed = true;
    }
    if (other.joinable()) other.join();
    else reduceSide();
    if (failed) return PFAC_STATUS_ALLOC_FAILED;
    return uploadTables(h);
}

PFAC_status_t loadImage(PFAC_handle_t h, const char* image, size_t size) {
    int st;
    try {
        st = pfac::buildMachine(image, size, h->machine);
    } catch (...) {
        st = PFAC_STATUS_ALLOC_FAILED;
    }
    if (st != PFAC_STATUS_SUCCESS) { freePatterns(h); return PFAC_status_t(st); }
    h->patternsReady = true;
    PFAC_status_t bs = bindTable(h);
    if (bs != PFAC_STATUS_SUCCESS) { freePatterns(h); return bs; }
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t readWholeFile(const char* filename, std::string& out) {
    FILE* fp = fopen(filename, "rb");
    if (!fp) return PFAC_STATUS_FILE_OPEN_ERROR;
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    rewind(fp);
    if (sz < 0) { fclose(fp); return PFAC_STATUS_FILE_OPEN_ERROR; }
    out.resize(size_t(sz));
    size_t got = sz ? fread(&out[0], 1, size_t(sz), fp) : 0;
    fclose(fp);
    out.resize(got);
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t cudaToStatus(cudaError_t e) {
    if (e == cudaSuccess) return PFAC_STATUS_SUCCESS;
    if (e == cudaErrorMemoryAllocation) return PFAC_STATUS_CUDA_ALLOC_FAILED;
    return PFAC_STATUS_INTERNAL_ERROR;
}

// workspace for one reduce launch; caller holds h->mu
PFAC_status_t ensureReduceWorkspace(PFAC_handle_t h, size_t words) {
    if (!h->d_total) {
        if (cudaMalloc(reinterpret_cast<void**>(&h->d_total), 16) != cudaSuccess) return PFAC_STATUS_CUDA_ALLOC_FAILED;
        if (cudaMallocHost(reinterpret_cast<void**>(&h->h_total), 128) != cudaSuccess) return PFAC_STATUS_ALLOC_FAILED;
        memset(h->h_total, 0, 128);   // [0]: count read-back, [8..15]: watchdog report of the reduce kernel
    }
    if (!h->d_park &&
        cudaMalloc(reinterpret_cast<void**>(&h->d_park), pfac::reduceParkWords(h->launch) * 8) != cudaSuccess)
        return PFAC_STATUS_CUDA_ALLOC_FAILED;
    if (words > h->wsWords) {
        cudaFree(h->d_ws);
        h->d_ws = nullptr;
        h->wsWords = 0;
        size_t want = words + words / 4 + 1024;
        if (cudaMalloc(reinterpret_cast<void**>(&h->d_ws), want * 8) != cudaSuccess) return PFAC_STATUS_CUDA_ALLOC_FAILED;
        h->wsWords = want;
    }
    return PFAC_STATUS_SUCCESS;
}

// enqueue one fused match+compaction over a device shard on `stream` (workspace reset + kernel); the
// count lands in h->d_total.  Caller holds h->mu.  comm != nullptr: the cross-GPU count exchange and
// scan run inside the same kernel.
PFAC_status_t reduceShardEnqueue(PFAC_handle_t h, const unsigned char* d_in, size_t n_owned, size_t n_total,
                                 long long pos_base, int* d_id, void* d_pos, bool pos64, cudaStream_t stream,
                                 const pfac::CommLaunch* comm, unsigned long long capacity = ~0ull,
                                 unsigned long long* total_out = nullptr) {
    const size_t words = pfac::reduceWorkspaceWords(n_owned);
    PFAC_status_t st = ensureReduceWorkspace(h, words);
    if (st != PFAC_STATUS_SUCCESS) return st;
    applyL2Window(h, stream);
    if (cudaMemsetAsync(h->d_ws, 0, words * 8, stream) != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
    if (cudaMemsetAsync(h->d_total, 0, 8, stream) != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
    return cudaToStatus(pfac::launchMatchReduce(h->tableReduce, h->launch, d_in, n_owned, n_total, pos_base, d_id,
                                                d_pos, pos64, h->d_ws, h->d_park, total_out ? total_out : h->d_total, stream,
                                                comm, h->h_total + 8, capacity));
}

// one fused match+compaction over a device shard; synchronous (returns the count)
PFAC_status_t reduceShard(PFAC_handle_t h, const unsigned char* d_in, size_t n_owned, size_t n_total,
                          long long pos_base, int* d_id, void* d_pos, bool pos64, cudaStream_t stream,
                          unsigned long long* count, unsigned long long capacity = ~0ull) {
    std::lock_guard<std::mutex> lock(h->mu);
    PFAC_status_t st = reduceShardEnqueue(h, d_in, n_owned, n_total, pos_base, d_id, d_pos, pos64, stream, nullptr, capacity);
    if (st != PFAC_STATUS_SUCCESS) return st;
    if (cudaMemcpyAsync(h->h_total, h->d_total, 8, cudaMemcpyDeviceToHost, stream) != cudaSuccess)
        return PFAC_STATUS_INTERNAL_ERROR;
    if (cudaStreamSynchronize(stream) != cudaSuccess) {
        const unsigned long long* d = h->h_total + 8;
        if (d[0])   // the kernel's watchdog fired: say where (pfac_kernels.cu, PFAC_SPIN_GUARD sites)
            fprintf(stderr, "libpfac: reduce kernel wait watchdog: site %llu block %llu thread %llu values %lld %lld %lld\n",
                    d[1], d[2] >> 32, d[2] & 0xFFFFFFFFull, (long long)d[3], (long long)d[4], (long long)d[5]);
        return PFAC_STATUS_INTERNAL_ERROR;
    }
    *count = *h->h_total;
    return PFAC_STATUS_SUCCESS;
}

// host pipeline buffers; caller holds h->pipeMu
PFAC_status_t ensurePipe(PFAC_handle_t h, size_t posBytes) {
    HostPipe& p = h->pipe;
    size_t chunk = envBytes("PFAC_B200_HOST_CHUNK_MB", 32, size_t(1) << 20);
    chunk = std::min(std::max(chunk, size_t(1) << 16), size_t(1) << 30);  // 64 KB .. 1 GiB (int positions per chunk)
    const size_t halo = size_t(h->machine.maxPatternLen > 1 ? h->machine.maxPatternLen - 1 : 0);
    const size_t inCap = ((chunk + halo + 255) / 256) * 256;
    if (p.chunk == chunk && p.inCap >= inCap && p.posBytes >= posBytes) return PFAC_STATUS_SUCCESS;
    posBytes = std::max(posBytes, p.posBytes);   // a handle that served a 64-bit call keeps the wider buffer
    freePipe(p);
    for (int i = 0; i < 2; i++) {
        if (cudaStreamCreateWithFlags(&p.stream[i], cudaStreamNonBlocking) != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
        if (cudaEventCreateWithFlags(&p.evH2D[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p.evDone[i], cudaEventDisableTiming) != cudaSuccess) { freePipe(p); return PFAC_STATUS_INTERNAL_ERROR; }
        if (cudaMalloc(reinterpret_cast<void**>(&p.d_in[i]), inCap) != cudaSuccess) { freePipe(p); return PFAC_STATUS_CUDA_ALLOC_FAILED; }
        if (cudaMalloc(reinterpret_cast<void**>(&p.d_out[i]), chunk * 4) != cudaSuccess) { freePipe(p); return PFAC_STATUS_CUDA_ALLOC_FAILED; }
        // sized by the position width the caller asked for (4 B for the legacy int calls, 8 B for the
        // 64-bit shard forms); the dense fallback of the sparse path reuses it as int scratch
        if (posBytes && cudaMalloc(reinterpret_cast<void**>(&p.d_pos[i]), chunk * posBytes) != cudaSuccess) { freePipe(p); return PFAC_STATUS_CUDA_ALLOC_FAILED; }
    }
    p.chunk = chunk;
    p.inCap = inCap;
    p.posBytes = posBytes;
    return PFAC_STATUS_SUCCESS;
}

// PFAC_B200_STAGE=0 hands pageable pointers straight to cudaMemcpyAsync (for A/B measurements)
bool stagingEnabled() { return envBytes("PFAC_B200_STAGE", 1, 1) != 0; }

// pinned staging buffers of the pageable paths; after ensurePipe, caller holds h->pipeMu
PFAC_status_t ensureStage(PFAC_handle_t h, bool needIn, bool needOut) {
    HostPipe& p = h->pipe;
    size_t sc = envBytes("PFAC_B200_STAGE_CHUNK_MB", 8, size_t(1) << 20);
    if (sc == 0 || sc > p.chunk) sc = p.chunk;
    const size_t halo = size_t(h->machine.maxPatternLen > 1 ? h->machine.maxPatternLen - 1 : 0);
    const size_t inCap = ((sc + halo + 255) / 256) * 256;
    if (p.stageChunk != sc || p.stageInCap < inCap) {
        freeStage(p);
        p.stageChunk = sc;
        p.stageInCap = inCap;
    }
    if (!p.pool) p.pool = acquireCopyPool();
    if (needIn && !p.s_in[0])
        for (int i = 0; i < 3; i++)
            if (cudaMallocHost(reinterpret_cast<void**>(&p.s_in[i]), inCap) != cudaSuccess) { freeStage(p); return PFAC_STATUS_ALLOC_FAILED; }
    if (needOut && !p.s_out[0])
        for (int i = 0; i < 2; i++)
            if (cudaMallocHost(reinterpret_cast<void**>(&p.s_out[i]), sc * 4) != cudaSuccess) { freeStage(p); return PFAC_STATUS_ALLOC_FAILED; }
    return PFAC_STATUS_SUCCESS;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* PFAC_versionString(void) { return "pfac-b200 sm_100a"; }

// reference PFAC.cpp:133-204.  No per-arch module to dlopen: the sm_100a kernels are in this
// library.  Without a usable device the raw CUDA error is returned, as the reference does.
PFAC_status_t PFAC_create(PFAC_handle_t* handle) {
    if (!handle) return PFAC_STATUS_INVALID_PARAMETER;
    *handle = nullptr;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return PFAC_status_t(e);
    int major = 0, sms = 0;
    if ((e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device)) != cudaSuccess) return PFAC_status_t(e);
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return PFAC_status_t(e);
    int minor = 0;
    if ((e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device)) != cudaSuccess) return PFAC_status_t(e);
    if (major != 10 || minor != 0) return PFAC_STATUS_ARCH_MISMATCH;  // kernels are built for sm_100a only
    PFAC_context* h = new (std::nothrow) PFAC_context();
    if (!h) return PFAC_STATUS_ALLOC_FAILED;
    h->device = device;
    h->launch.numSMs = sms;
    *handle = h;
    return PFAC_STATUS_SUCCESS;
}

// reference PFAC.cpp:207-218
PFAC_status_t PFAC_destroy(PFAC_handle_t handle) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    freePatterns(handle);
    freePipe(handle->pipe);
    cudaFree(handle->d_ws);
    cudaFree(handle->d_park);
    cudaFree(handle->d_total);
    if (handle->h_total) cudaFreeHost(handle->h_total);
    delete handle;
    return PFAC_STATUS_SUCCESS;
}

// reference PFAC.cpp:741-757
PFAC_status_t PFAC_setPlatform(PFAC_handle_t handle, PFAC_platform_t platform) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (platform != PFAC_PLATFORM_GPU && platform != PFAC_PLATFORM_CPU && platform != PFAC_PLATFORM_CPU_OMP)
        return PFAC_STATUS_INVALID_PARAMETER;
    handle->platform = int(platform);
    return PFAC_STATUS_SUCCESS;
}

// reference PFAC.cpp:764-780
PFAC_status_t PFAC_setTextureMode(PFAC_handle_t handle, PFAC_textureMode_t mode) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (mode != PFAC_AUTOMATIC && mode != PFAC_TEXTURE_ON && mode != PFAC_TEXTURE_OFF)
        return PFAC_STATUS_INVALID_PARAMETER;
    handle->textureMode = int(mode);
    return PFAC_STATUS_SUCCESS;
}

// reference PFAC.cpp:782-817: a mode change with patterns loaded rebuilds the device table
PFAC_status_t PFAC_setPerfMode(PFAC_handle_t handle, PFAC_perfMode_t mode) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (mode != PFAC_TIME_DRIVEN && mode != PFAC_SPACE_DRIVEN) return PFAC_STATUS_INVALID_PARAMETER;
    const bool rebuild = handle->patternsReady && (int(mode) != handle->perfMode);
    handle->perfMode = int(mode);
    if (rebuild) {
        std::lock_guard<std::mutex> lock(handle->mu);
        cudaDeviceSynchronize();  // no kernel may still be reading the old table
        PFAC_status_t st = bindTable(handle);
        if (st != PFAC_STATUS_SUCCESS) { freePatterns(handle); return st; }  // no tables: patterns are not ready
    }
    return PFAC_STATUS_SUCCESS;
}

// reference PFAC.cpp:1131-1185 (same text: callers print these)
const char* PFAC_getErrorString(PFAC_status_t status) {
    if (status == PFAC_STATUS_SUCCESS) return "PFAC_STATUS_SUCCESS: operation is successful";
    if (status < PFAC_STATUS_BASE) return cudaGetErrorString(cudaError_t(status));
    static const char* const text[] = {
        "PFAC_STATUS_ALLOC_FAILED: allocation fails on host memory",
        "PFAC_STATUS_CUDA_ALLOC_FAILED: allocation fails on device memory",
        "PFAC_STATUS_INVALID_HANDLE: handle is invalid (NULL)",
        "PFAC_STATUS_INVALID_PARAMETER: parameter is invalid",
        "PFAC_STATUS_PATTERNS_NOT_READY: please call PFAC_readPatternFromFile() first",
        "PFAC_STATUS_FILE_OPEN_ERROR: pattern file does not exist",
        "PFAC_STATUS_LIB_NOT_EXIST: cannot find PFAC library, please check LD_LIBRARY_PATH",
        "PFAC_STATUS_ARCH_MISMATCH: sm1.0 is not supported",
        "PFAC_STATUS_MUTEX_ERROR: please report bugs. Workaround: choose non-texture mode.",
    };
    const int idx = int(status) - int(PFAC_STATUS_ALLOC_FAILED);
    if (idx >= 0 && idx < int(sizeof(text) / sizeof(text[0]))) return text[idx];
    return "PFAC_STATUS_INTERNAL_ERROR: please report bugs";
}

// reference PFAC.cpp:1188-1246
PFAC_status_t PFAC_dumpTransitionTable(PFAC_handle_t handle, FILE* fp) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!fp) fp = stdout;
    pfac::dumpMachine(handle->machine, fp);
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_dumpTransitionTableToFile(PFAC_handle_t handle, const char* filename) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!filename) return PFAC_STATUS_INVALID_PARAMETER;
    FILE* fp = fopen(filename, "w");
    if (!fp) return PFAC_STATUS_FILE_OPEN_ERROR;
    pfac::dumpMachine(handle->machine, fp);
    fclose(fp);
    return PFAC_STATUS_SUCCESS;
}

// reference PFAC.cpp:653-735
PFAC_status_t PFAC_readPatternFromFile(PFAC_handle_t handle, char* filename) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!filename) return PFAC_STATUS_INVALID_PARAMETER;
    std::lock_guard<std::mutex> lock(handle->mu);
    if (handle->patternsReady) {
        cudaDeviceSynchronize();
        freePatterns(handle);
    }
    if (strlen(filename) >= kFilenameLen) return PFAC_STATUS_INTERNAL_ERROR;  // as :668-672
    strcpy(handle->patternFile, filename);
    std::string image;
    PFAC_status_t st = readWholeFile(filename, image);
    if (st != PFAC_STATUS_SUCCESS) { freePatterns(handle); return st; }
    return loadImage(handle, image.data(), image.size());
}

PFAC_status_t PFAC_readPatternFromMemory(PFAC_handle_t handle, const char* image, size_t size) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!image && size) return PFAC_STATUS_INVALID_PARAMETER;
    std::lock_guard<std::mutex> lock(handle->mu);
    if (handle->patternsReady) {
        cudaDeviceSynchronize();
        freePatterns(handle);
    }
    handle->patternFile[0] = 0;
    return loadImage(handle, image, size);
}

PFAC_status_t PFAC_readPatternFromArrays(PFAC_handle_t handle, const char* const* patterns, const size_t* lengths,
                                         size_t num_patterns) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    std::lock_guard<std::mutex> lock(handle->mu);
    if (handle->patternsReady) {
        cudaDeviceSynchronize();
        freePatterns(handle);
    }
    handle->patternFile[0] = 0;
    int st;
    try {
        st = pfac::buildMachineFromArrays(patterns, lengths, num_patterns, handle->machine);
    } catch (...) {
        st = PFAC_STATUS_ALLOC_FAILED;
    }
    if (st != PFAC_STATUS_SUCCESS) { freePatterns(handle); return PFAC_status_t(st); }
    handle->patternsReady = true;
    PFAC_status_t bs = bindTable(handle);
    if (bs != PFAC_STATUS_SUCCESS) { freePatterns(handle); return bs; }
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_setStream(PFAC_handle_t handle, void* cuda_stream) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    handle->stream = static_cast<cudaStream_t>(cuda_stream);
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_matchShardFromDevice(PFAC_handle_t handle, const char* d_in, size_t n_owned,
                                        size_t n_total, int* d_out) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (!d_in || !d_out) return PFAC_STATUS_INVALID_PARAMETER;
    if (n_total < n_owned) return PFAC_STATUS_INVALID_PARAMETER;
    if (n_owned == 0) return PFAC_STATUS_SUCCESS;
    applyL2Window(handle, handle->stream);
    return cudaToStatus(pfac::launchMatchDense(handle->table, handle->launch,
                                               reinterpret_cast<const unsigned char*>(d_in), n_owned,
                                               n_total, d_out, handle->stream));
}

// reference PFAC.cpp:843-876: same checks in the same order; platform is immaterial
PFAC_status_t PFAC_matchFromDevice(PFAC_handle_t handle, char* d_in, size_t size, int* d_out) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (!d_in) return PFAC_STATUS_INVALID_PARAMETER;
    if (!d_out) return PFAC_STATUS_INVALID_PARAMETER;
    if (size == 0) return PFAC_STATUS_SUCCESS;
    applyL2Window(handle, handle->stream);
    return cudaToStatus(pfac::launchMatchDense(handle->table, handle->launch,
                                               reinterpret_cast<const unsigned char*>(d_in), size, size,
                                               d_out, handle->stream));
}

// PFAC_B200_HOST_RESULT=dense: always bring the 4-byte-per-position array back over PCIe
static bool sparseHostResult() {
    const char* v = getenv("PFAC_B200_HOST_RESULT");
    return !(v && !strcmp(v, "dense"));
}

// PFAC_matchFromHost, sparse result path.  The dense result is almost all zeros, and returning it
// costs 4 bytes of PCIe per input byte (13.7 GB/s of input at best).  Instead every chunk goes
// through the fused match + compaction kernel, only its (id, position) pairs come back, and the
// host writes the dense array itself: the copy pool zero-fills the chunk's slice of h_out with
// streaming stores (about 200 GB/s on the 16-core boxes) while the chunk is on the GPU, then the
// pairs are scattered into it.  A chunk with more than one match per 16 positions falls back to the
// dense kernel + a plain D2H of its slice.  Works the same for pinned and pageable h_out.
static PFAC_status_t hostDenseSparse(PFAC_handle_t handle, const char* h_in, size_t n_owned, size_t n_total,
                                     int* h_out) {
    std::lock_guard<std::mutex> lock(handle->pipeMu);
    {
        PFAC_status_t st = ensurePipe(handle, sizeof(int));
        if (st != PFAC_STATUS_SUCCESS) return st;
    }
    HostPipe& p = handle->pipe;
    const bool stIn = stagingEnabled() && isPageable(h_in);
    {
        PFAC_status_t st = ensureStage(handle, stIn, false);
        if (st != PFAC_STATUS_SUCCESS) return st;
    }
    const size_t chunk = stIn ? p.stageChunk : p.chunk;
    const size_t cap = std::max<size_t>(chunk / 16, 1024);  // pairs a chunk may return before it goes dense
    if (p.listCap < cap) {
        for (int i = 0; i < 2; i++) {
            if (p.s_list[i]) cudaFreeHost(p.s_list[i]);
            p.s_list[i] = nullptr;
            if (cudaMallocHost(reinterpret_cast<void**>(&p.s_list[i]), cap * 8) != cudaSuccess) {
                p.listCap = 0;
                return PFAC_STATUS_ALLOC_FAILED;
            }
        }
        p.listCap = cap;
    }
    if (!p.h_cnt && cudaMallocHost(reinterpret_cast<void**>(&p.h_cnt), 128) != cudaSuccess) return PFAC_STATUS_ALLOC_FAILED;
    const size_t nchunks = (n_owned + chunk - 1) / chunk;
    const size_t halo = size_t(handle->machine.maxPatternLen > 1 ? handle->machine.maxPatternLen - 1 : 0);
    auto ownedOf = [&](size_t c) { return (n_owned - c * chunk < chunk) ? n_owned - c * chunk : chunk; };
    auto totalOf = [&](size_t c) {
        const size_t off = c * chunk, owned = ownedOf(c);
        return (n_total - off < owned + halo) ? n_total - off : owned + halo;
    };
    // Nothing in the loop blocks on the GPU except the wait for a chunk's own kernel:
    //   copy stream    : H2D of chunk c (after the kernel of chunk c-2 released the slot's input buffer)
    //   compute stream : fused match + compaction of chunk c; the kernel stores the (id, position) pairs
    //                    and the count straight into pinned host memory (zero-copy: a few bytes per match)
    //   copy pool      : zero fill of the h_out slice of chunk c+2, host copy of pageable chunk c+3
    //   this thread    : helps with the fill of chunk c, waits for its kernel, scatters its pairs,
    //                    enqueues chunk c+2
    // The kernels run one after the other on one stream, so they share the handle's reduce workspace;
    // handle->mu is held for the whole call so that no other thread's reduce call gets in between.
    std::lock_guard<std::mutex> wsLock(handle->mu);
    cudaStream_t copyS = p.stream[0], compS = p.stream[1];
    CopyPool::Job jobs[3];
    CopyPool::Job zero[2];  // the fill runs two chunks ahead, so the workers never wait for the GPU
    auto hostStage = [&](size_t c) {
        if (stIn && c < nchunks) p.pool->start(jobs[c % 3], p.s_in[c % 3], h_in + c * chunk, totalOf(c));
    };
    auto zeroFill = [&](size_t c) {
        if (c < nchunks) p.pool->startZero(zero[c % 2], h_out + c * chunk, ownedOf(c) * sizeof(int));
    };
    auto bail = [&](PFAC_status_t st) {
        if (stIn) for (int j = 0; j < 3; j++) p.pool->finish(jobs[j]);
        p.pool->finish(zero[0]);
        p.pool->finish(zero[1]);
        cudaDeviceSynchronize();
        return st;
    };
    auto enqueue = [&](size_t c) -> PFAC_status_t {
        const int slot = int(c % 2);
        if (c >= 2 && cudaStreamWaitEvent(copyS, p.evDone[slot], 0) != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
        const void* src = h_in + c * chunk;
        if (stIn) {
            p.pool->finish(jobs[c % 3]);
            src = p.s_in[c % 3];
        }
        p.lastH2D += totalOf(c);
        if (cudaMemcpyAsync(p.d_in[slot], src, totalOf(c), cudaMemcpyHostToDevice, copyS) != cudaSuccess ||
            cudaEventRecord(p.evH2D[slot], copyS) != cudaSuccess ||
            cudaStreamWaitEvent(compS, p.evH2D[slot], 0) != cudaSuccess)
            return PFAC_STATUS_INTERNAL_ERROR;
        PFAC_status_t st = reduceShardEnqueue(handle, p.d_in[slot], ownedOf(c), totalOf(c), 0, p.s_list[slot],
                                              p.s_list[slot] + cap, false, compS, nullptr, cap, p.h_cnt + slot * 8);
        if (st != PFAC_STATUS_SUCCESS) return st;
        return cudaEventRecord(p.evDone[slot], compS) == cudaSuccess ? PFAC_STATUS_SUCCESS : PFAC_STATUS_INTERNAL_ERROR;
    };
    p.lastH2D = p.lastD2H = 0;
    if (!p.pool) p.pool = acquireCopyPool();
    hostStage(0);
    hostStage(1);
    hostStage(2);
    zeroFill(0);
    zeroFill(1);
    for (size_t c = 0; c < 2 && c < nchunks; c++) {
        PFAC_status_t st = enqueue(c);
        if (st != PFAC_STATUS_SUCCESS) return bail(st);
    }
    for (size_t c = 0; c < nchunks; c++) {
        const int slot = int(c % 2);
        const size_t off = c * chunk, owned = ownedOf(c);
        p.pool->finish(zero[slot]);                       // this thread fills too while the GPU works
        if (cudaEventSynchronize(p.evDone[slot]) != cudaSuccess) {
            const unsigned long long* d = handle->h_total + 8;
            if (d[0])
                fprintf(stderr, "libpfac: reduce kernel wait watchdog: site %llu block %llu thread %llu values %lld %lld %lld\n",
                        d[1], d[2] >> 32, d[2] & 0xFFFFFFFFull, (long long)d[3], (long long)d[4], (long long)d[5]);
            return bail(PFAC_STATUS_INTERNAL_ERROR);
        }
        hostStage(c + 3);                                 // its staging slot is that of chunk c, now on the device
        const unsigned long long count = p.h_cnt[slot * 8];
        p.lastD2H += 8;
        if (count <= cap) {
            const int* ids = p.s_list[slot];
            const int* pos = p.s_list[slot] + cap;
            int* dst = h_out + off;
            for (unsigned long long i = 0; i < count; i++) dst[pos[i]] = ids[i];
            p.lastD2H += count * 8;
        } else {
            // dense chunk (more than one match per 16 positions): the dense kernel + a plain D2H of its slice
            int* d_dense = p.d_out[slot];
            cudaError_t e = pfac::launchMatchDense(handle->table, handle->launch, p.d_in[slot], owned, totalOf(c),
                                                   d_dense, compS);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(h_out + off, d_dense, owned * sizeof(int), cudaMemcpyDeviceToHost, compS);
            if (e == cudaSuccess) e = cudaEventRecord(p.evDone[slot], compS);   // the slot's input is in use until here
            if (e == cudaSuccess) e = cudaStreamSynchronize(compS);
            if (e != cudaSuccess) return bail(cudaToStatus(e));
            p.lastD2H += owned * sizeof(int);
        }
        zeroFill(c + 2);
        if (c + 2 < nchunks) {
            PFAC_status_t st = enqueue(c + 2);
            if (st != PFAC_STATUS_SUCCESS) return bail(st);
        }
    }
    if (cudaStreamSynchronize(copyS) != cudaSuccess || cudaStreamSynchronize(compS) != cudaSuccess)
        return PFAC_STATUS_INTERNAL_ERROR;
    return PFAC_STATUS_SUCCESS;
}

// reference PFAC.cpp:879-961.  Chunks of the host input (+ tail halo) go H2D on two private
// streams, each followed by its kernel and the D2H of its 4-byte-per-position results, so
// copy-in, match and copy-out of neighbouring chunks overlap.  Pinned user buffers are DMA'd
// directly; pageable ones go through the pinned staging buffers, copied by the CopyPool while the
// other slot's chunk is on the GPU.  Synchronous, like the reference.
// host shard: results for [0,n_owned), input bytes [0,n_total) (owned + tail halo), both on the host
static PFAC_status_t hostDenseShard(PFAC_handle_t handle, const char* h_in, size_t n_owned, size_t n_total,
                                    int* h_out) {
    if (sparseHostResult()) return hostDenseSparse(handle, h_in, n_owned, n_total, h_out);
    std::lock_guard<std::mutex> lock(handle->pipeMu);
    PFAC_status_t st = ensurePipe(handle, 0);
    if (st != PFAC_STATUS_SUCCESS) return st;
    HostPipe& p = handle->pipe;
    const bool stIn = stagingEnabled() && isPageable(h_in);
    const bool stOut = stagingEnabled() && isPageable(h_out);
    const bool staged = stIn || stOut;
    if (staged && (st = ensureStage(handle, stIn, stOut)) != PFAC_STATUS_SUCCESS) return st;
    const size_t chunk = staged ? p.stageChunk : p.chunk;
    const size_t halo = size_t(handle->machine.maxPatternLen > 1 ? handle->machine.maxPatternLen - 1 : 0);
    struct Pending {
        size_t off = 0, owned = 0;
        bool live = false;
    } pend[2];
    // wait for a slot's chunk; staged results then move from the pinned buffer to the caller's
    auto drain = [&](int sl) -> cudaError_t {
        if (!pend[sl].live) return cudaSuccess;
        pend[sl].live = false;
        cudaError_t r = cudaStreamSynchronize(p.stream[sl]);
        if (r == cudaSuccess && stOut) p.pool->copy(h_out + pend[sl].off, p.s_out[sl], pend[sl].owned * sizeof(int), true);
        return r;
    };
    cudaError_t e = cudaSuccess;
    int slot = 0;
    p.lastH2D = p.lastD2H = 0;
    for (size_t off = 0; off < n_owned && e == cudaSuccess; off += chunk, slot ^= 1) {
        const size_t owned = (n_owned - off < chunk) ? n_owned - off : chunk;
        const size_t total = (n_total - off < owned + halo) ? n_total - off : owned + halo;
        cudaStream_t s = p.stream[slot];
        const void* src = h_in + off;
        if (stIn) {  // this slot's previous chunk was drained one iteration ago
            p.pool->copy(p.s_in[slot], h_in + off, total);
            src = p.s_in[slot];
        }
        e = cudaMemcpyAsync(p.d_in[slot], src, total, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) break;
        p.lastH2D += total;
        p.lastD2H += owned * sizeof(int);
        e = pfac::launchMatchDense(handle->table, handle->launch, p.d_in[slot], owned, total, p.d_out[slot], s);
        if (e != cudaSuccess) break;
        e = cudaMemcpyAsync(stOut ? p.s_out[slot] : h_out + off, p.d_out[slot], owned * sizeof(int),
                            cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) break;
        pend[slot].off = off;
        pend[slot].owned = owned;
        pend[slot].live = true;
        if (staged) e = drain(slot ^ 1);  // the chunk queued just now keeps the GPU busy meanwhile
    }
    cudaError_t e0 = drain(slot);  // older chunk first
    cudaError_t e1 = drain(slot ^ 1);
    cudaStreamSynchronize(p.stream[0]);
    cudaStreamSynchronize(p.stream[1]);
    if (e != cudaSuccess) return cudaToStatus(e);
    if (e0 != cudaSuccess) return cudaToStatus(e0);
    return cudaToStatus(e1);
}

PFAC_status_t PFAC_matchFromHost(PFAC_handle_t handle, char* h_in, size_t size, int* h_out) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (!h_in) return PFAC_STATUS_INVALID_PARAMETER;
    if (!h_out) return PFAC_STATUS_INVALID_PARAMETER;
    if (size == 0) return PFAC_STATUS_SUCCESS;
    return hostDenseShard(handle, h_in, size, size, h_out);
}

PFAC_status_t PFAC_matchShardFromDeviceReduce64(PFAC_handle_t handle, const char* d_in, size_t n_owned,
                                                size_t n_total, long long pos_base, int* d_id,
                                                long long* d_pos, unsigned long long* h_num) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (!d_in || !h_num || !d_pos || !d_id) return PFAC_STATUS_INVALID_PARAMETER;
    if (n_total < n_owned) return PFAC_STATUS_INVALID_PARAMETER;
    *h_num = 0;
    if (n_owned == 0) return PFAC_STATUS_SUCCESS;
    return reduceShard(handle, reinterpret_cast<const unsigned char*>(d_in), n_owned, n_total, pos_base,
                       d_id, d_pos, true, handle->stream, h_num);
}

// the same with the capacity of d_id / d_pos stated: nothing is stored past it, *h_num is the full count
// (larger than `capacity` = the caller's buffers were too small: PFAC_STATUS_INVALID_PARAMETER)
PFAC_status_t PFAC_matchShardFromDeviceReduce64Cap(PFAC_handle_t handle, const char* d_in, size_t n_owned,
                                                   size_t n_total, long long pos_base, int* d_id, long long* d_pos,
                                                   size_t capacity, unsigned long long* h_num) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (!d_in || !h_num || !d_pos || !d_id) return PFAC_STATUS_INVALID_PARAMETER;
    if (n_total < n_owned) return PFAC_STATUS_INVALID_PARAMETER;
    *h_num = 0;
    if (n_owned == 0) return PFAC_STATUS_SUCCESS;
    PFAC_status_t st = reduceShard(handle, reinterpret_cast<const unsigned char*>(d_in), n_owned, n_total, pos_base,
                                   d_id, d_pos, true, handle->stream, h_num, capacity);
    if (st != PFAC_STATUS_SUCCESS) return st;
    return *h_num > capacity ? PFAC_STATUS_INVALID_PARAMETER : PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_matchFromDeviceReduce64(PFAC_handle_t handle, const char* d_in, size_t size, int* d_id,
                                           long long* d_pos, unsigned long long* h_num) {
    return PFAC_matchShardFromDeviceReduce64(handle, d_in, size, size, 0, d_id, d_pos, h_num);
}

// reference PFAC.cpp:964-1008.  NULL checks are the reference's, in its order: handle, input,
// h_num_matched, d_pos (it checks neither isPatternsReady nor d_matched_result; with no
// patterns the reference would dereference a NULL table -- here that is PATTERNS_NOT_READY).
PFAC_status_t PFAC_matchFromDeviceReduce(PFAC_handle_t handle, char* d_in, size_t size, int* d_id,
                                         int* d_pos, int* h_num) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!d_in) return PFAC_STATUS_INVALID_PARAMETER;
    if (!h_num) return PFAC_STATUS_INVALID_PARAMETER;
    if (!d_pos) return PFAC_STATUS_INVALID_PARAMETER;
    if (size == 0) return PFAC_STATUS_SUCCESS;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (!d_id) return PFAC_STATUS_INVALID_PARAMETER;
    if (size >= kInt32Limit) return PFAC_STATUS_INVALID_PARAMETER;  // int positions; use the 64-bit call
    unsigned long long count = 0;
    PFAC_status_t st = reduceShard(handle, reinterpret_cast<const unsigned char*>(d_in), size, size, 0, d_id,
                                   d_pos, false, handle->stream, &count);
    if (st != PFAC_STATUS_SUCCESS) return st;
    *h_num = int(count);
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_reduceOnDevice(PFAC_handle_t handle, char* d_in, size_t size, int* d_id, int* d_pos,
                                  int* h_num) {
    return PFAC_matchFromDeviceReduce(handle, d_in, size, d_id, d_pos, h_num);
}

PFAC_status_t PFAC_reduceInplaceOnDevice(PFAC_handle_t handle, char* d_in, size_t size, int* d_id,
                                         int* d_pos, int* h_num) {
    return PFAC_matchFromDeviceReduce(handle, d_in, size, d_id, d_pos, h_num);
}

// Host buffers, reduced result (reference PFAC.cpp:1010-1128).  Chunked: the host copy of chunk
// c+2 into pinned staging (pageable input only) and the H2D of chunk c+1 overlap the fused
// match+compaction of chunk c; only 8 bytes per match come back.  Positions are global (chunk
// offset added on the device), lists are appended in chunk order, so the result is ascending in
// position.
// host shard, reduced: appends (id, position) for owned positions; position = pos_base + local.
// h_pos is int* (pos64 == false) or long long* (pos64 == true).
static PFAC_status_t hostReduceShard(PFAC_handle_t handle, const char* h_in, size_t n_owned, size_t n_total,
                                     long long pos_base, int* h_id, void* h_pos, bool pos64,
                                     unsigned long long* h_num) {
    std::lock_guard<std::mutex> lock(handle->pipeMu);
    {
        PFAC_status_t st = ensurePipe(handle, pos64 ? sizeof(long long) : sizeof(int));
        if (st != PFAC_STATUS_SUCCESS) return st;
    }
    HostPipe& p = handle->pipe;
    const bool stIn = stagingEnabled() && isPageable(h_in);
    if (stIn) {
        PFAC_status_t st = ensureStage(handle, true, false);
        if (st != PFAC_STATUS_SUCCESS) return st;
    }
    const size_t chunk = stIn ? p.stageChunk : p.chunk;
    const size_t nchunks = (n_owned + chunk - 1) / chunk;
    const size_t halo = size_t(handle->machine.maxPatternLen > 1 ? handle->machine.maxPatternLen - 1 : 0);
    const size_t posBytes = pos64 ? 8 : 4;
    size_t written = 0;
    p.lastH2D = p.lastD2H = 0;
    auto ownedOf = [&](size_t c) { return (n_owned - c * chunk < chunk) ? n_owned - c * chunk : chunk; };
    auto totalOf = [&](size_t c) {
        const size_t off = c * chunk, owned = ownedOf(c);
        return (n_total - off < owned + halo) ? n_total - off : owned + halo;
    };
    // three stages in flight: [host copy of chunk c+2 into pinned staging] || [H2D of chunk c+1] ||
    // [fused match+compaction of chunk c].  Pinned input skips the first.
    CopyPool::Job jobs[3];
    auto hostStage = [&](size_t c) {
        if (stIn && c < nchunks) p.pool->start(jobs[c % 3], p.s_in[c % 3], h_in + c * chunk, totalOf(c));
    };
    auto h2d = [&](size_t c) -> cudaError_t {
        const void* src = h_in + c * chunk;
        if (stIn) {
            p.pool->finish(jobs[c % 3]);
            src = p.s_in[c % 3];
        }
        p.lastH2D += totalOf(c);
        return cudaMemcpyAsync(p.d_in[c % 2], src, totalOf(c), cudaMemcpyHostToDevice, p.stream[c % 2]);
    };
    auto bail = [&](PFAC_status_t st) {
        if (stIn) for (int j = 0; j < 3; j++) p.pool->finish(jobs[j]);
        cudaDeviceSynchronize();
        return st;
    };
    hostStage(0);
    hostStage(1);
    if (h2d(0) != cudaSuccess) return bail(PFAC_STATUS_INTERNAL_ERROR);
    for (size_t c = 0; c < nchunks; c++) {
        const int slot = int(c % 2);
        // slot^1's previous chunk (c-1) has been reduced and read back: its buffers are free
        if (c + 1 < nchunks && h2d(c + 1) != cudaSuccess) return bail(PFAC_STATUS_INTERNAL_ERROR);
        hostStage(c + 2);  // staging slot of chunk c-1, whose H2D finished before its reduce did
        unsigned long long count = 0;
        PFAC_status_t st = reduceShard(handle, p.d_in[slot], ownedOf(c), totalOf(c), pos_base + (long long)(c * chunk),
                                       p.d_out[slot], p.d_pos[slot], pos64, p.stream[slot], &count);
        if (st != PFAC_STATUS_SUCCESS) return bail(st);
        if (count) {
            if (cudaMemcpyAsync(h_id + written, p.d_out[slot], count * 4, cudaMemcpyDeviceToHost, p.stream[slot]) != cudaSuccess ||
                cudaMemcpyAsync(static_cast<char*>(h_pos) + written * posBytes, p.d_pos[slot], count * posBytes,
                                cudaMemcpyDeviceToHost, p.stream[slot]) != cudaSuccess)
                return bail(PFAC_STATUS_INTERNAL_ERROR);
            written += count;
        }
        p.lastD2H += 8 + count * (4 + posBytes);
        // the slot is reused two chunks later: its D2H must have drained before the next H2D into it
        if (cudaStreamSynchronize(p.stream[slot]) != cudaSuccess) return bail(PFAC_STATUS_INTERNAL_ERROR);
    }
    if (cudaStreamSynchronize(p.stream[0]) != cudaSuccess || cudaStreamSynchronize(p.stream[1]) != cudaSuccess)
        return PFAC_STATUS_INTERNAL_ERROR;
    *h_num = written;
    return PFAC_STATUS_SUCCESS;
}

// reference PFAC.cpp:1010-1128: same NULL checks; the work is hostReduceShard above
PFAC_status_t PFAC_matchFromHostReduce(PFAC_handle_t handle, char* h_in, size_t size, int* h_id, int* h_pos,
                                       int* h_num) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (!h_in) return PFAC_STATUS_INVALID_PARAMETER;
    if (!h_id) return PFAC_STATUS_INVALID_PARAMETER;
    if (!h_pos) return PFAC_STATUS_INVALID_PARAMETER;
    if (!h_num) return PFAC_STATUS_INVALID_PARAMETER;
    if (size == 0) return PFAC_STATUS_SUCCESS;
    if (size >= kInt32Limit) return PFAC_STATUS_INVALID_PARAMETER;
    unsigned long long n = 0;
    PFAC_status_t st = hostReduceShard(handle, h_in, size, size, 0, h_id, h_pos, false, &n);
    if (st != PFAC_STATUS_SUCCESS) return st;
    *h_num = int(n);
    return PFAC_STATUS_SUCCESS;
}

// reference PFAC_memoryUsage (PFAC.cpp:1250-1306, not in the public header): table sizes to stdout
PFAC_status_t PFAC_memoryUsage(PFAC_handle_t handle) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    const pfac::Machine& m = handle->machine;
    const pfac::DeviceLayout& L = handle->layout;
    const double dense2d = double(m.numStates) * 256.0 * 4.0;
    printf("%s: symbol code %d bits x %d, prefilter 8192 bytes (%d of 65536 bits set)\n",
           handle->perfMode == PFAC_SPACE_DRIVEN ? "space-driven" : "time-driven", L.codeBits, L.gramLen,
           L.pre2BitsSet);
    printf("next2 = %zu entries, hash rows = %d entries (%u hot + %u cold buckets), chains = %d, tails = %zu bytes\n",
           L.next2.size(), L.hashEdges, L.hotBuckets, L.coldBuckets, L.numChains, L.tails.size());
    printf("total amount = %7.2f MB\n", double(L.deviceBytes()) / 1024. / 1024.);
    printf("(device layout)/(2-D table) = %5.3f\n", double(L.deviceBytes()) / dense2d);
    printf("S = number of states (ignore s0) = %d \n", m.numStates - 1);
    printf("F = number of final states = %d \n", m.numFinal);
    printf("L = number of leaf nodes = %d\n", m.numLeaves);
    return PFAC_STATUS_SUCCESS;
}

// ---- cross-GPU count exchange + global list (PFAC_ext.h) -------------------------------------------
// The path's only inter-GPU step: an exclusive scan of the per-GPU match counts.
// It runs on the device, inside the reduce kernel, over NVLink peer memory: no NCCL call, no host
// round trip.  The reference's only inter-GPU mechanism is peer access to caller buffers
// (reference test/UVA.cpp:137); its multi-GPU program stitches results on the host
// (test/omp_PFAC.cpp:351-394).
static PFAC_status_t commAlloc(PFAC_comm* c, size_t list_capacity) {
    c->listCap = list_capacity;
    c->blockBytes = kCommMailboxBytes + commIdsBytes(list_capacity) + list_capacity * sizeof(long long) + 256;
    if (cudaMalloc(reinterpret_cast<void**>(&c->block), c->blockBytes) != cudaSuccess) return PFAC_STATUS_CUDA_ALLOC_FAILED;
    if (cudaMemset(c->block, 0, kCommMailboxBytes) != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
    const unsigned long long cap64 = list_capacity;
    if (cudaMemcpy(c->block + kCommCapWord * 8, &cap64, 8, cudaMemcpyHostToDevice) != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
    c->peerCap[c->rank] = list_capacity;
    if (cudaMalloc(reinterpret_cast<void**>(&c->d_scan), 32) != cudaSuccess) return PFAC_STATUS_CUDA_ALLOC_FAILED;
    if (cudaMallocHost(reinterpret_cast<void**>(&c->h_scan), 32) != cudaSuccess) return PFAC_STATUS_ALLOC_FAILED;
    if (cudaDeviceSynchronize() != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
    c->peer[c->rank] = c->block;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_commCreate(PFAC_comm_t* comm, int rank, int world, size_t list_capacity, void* ipc_handle_out) {
    if (!comm) return PFAC_STATUS_INVALID_PARAMETER;
    *comm = nullptr;
    if (world < 1 || world > pfac::kKernelCommMaxRanks || rank < 0 || rank >= world) return PFAC_STATUS_INVALID_PARAMETER;
    if (world > 1 && !ipc_handle_out) return PFAC_STATUS_INVALID_PARAMETER;
    PFAC_comm* c = new (std::nothrow) PFAC_comm();
    if (!c) return PFAC_STATUS_ALLOC_FAILED;
    c->rank = rank;
    c->world = world;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e != cudaSuccess) { delete c; return PFAC_status_t(e); }
    PFAC_status_t st = commAlloc(c, list_capacity);
    if (st == PFAC_STATUS_SUCCESS && ipc_handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == PFAC_COMM_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t hd;
        e = cudaIpcGetMemHandle(&hd, c->block);
        if (e != cudaSuccess) st = PFAC_status_t(e);
        else memcpy(ipc_handle_out, &hd, sizeof(hd));
    }
    if (st != PFAC_STATUS_SUCCESS) { PFAC_commDestroy(c); return st; }
    *comm = c;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_commConnect(PFAC_comm_t comm, const void* all_handles) {
    if (!comm) return PFAC_STATUS_INVALID_HANDLE;
    if (comm->world > 1 && !all_handles) return PFAC_STATUS_INVALID_PARAMETER;
    for (int r = 0; r < comm->world; r++) {
        if (r == comm->rank || comm->peer[r]) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, static_cast<const char*>(all_handles) + size_t(r) * sizeof(hd), sizeof(hd));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return PFAC_status_t(e);
        comm->peer[r] = static_cast<unsigned char*>(ptr);
        comm->ipcOpened[r] = true;
        unsigned long long cap64 = 0;   // where that rank's position array starts depends on ITS capacity
        e = cudaMemcpy(&cap64, comm->peer[r] + kCommCapWord * 8, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return PFAC_status_t(e);
        comm->peerCap[r] = size_t(cap64);
    }
    return PFAC_STATUS_SUCCESS;
}

// one process driving several GPUs: blocks are mapped through peer access instead of IPC
PFAC_status_t PFAC_commCreateLocal(PFAC_comm_t* comms, const int* devices, int num_devices, size_t list_capacity) {
    if (!comms || !devices || num_devices < 1 || num_devices > pfac::kKernelCommMaxRanks) return PFAC_STATUS_INVALID_PARAMETER;
    int saved = 0;
    cudaGetDevice(&saved);
    for (int i = 0; i < num_devices; i++) comms[i] = nullptr;
    PFAC_status_t st = PFAC_STATUS_SUCCESS;
    for (int i = 0; i < num_devices && st == PFAC_STATUS_SUCCESS; i++) {
        cudaError_t e = cudaSetDevice(devices[i]);
        if (e != cudaSuccess) { st = PFAC_status_t(e); break; }
        PFAC_comm* c = new (std::nothrow) PFAC_comm();
        if (!c) { st = PFAC_STATUS_ALLOC_FAILED; break; }
        comms[i] = c;
        c->rank = i;
        c->world = num_devices;
        c->device = devices[i];
        st = commAlloc(c, list_capacity);
        for (int j = 0; j < num_devices && st == PFAC_STATUS_SUCCESS; j++) {
            if (devices[j] == devices[i]) continue;
            e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) st = PFAC_status_t(e);
        }
    }
    for (int i = 0; i < num_devices && st == PFAC_STATUS_SUCCESS; i++)
        for (int j = 0; j < num_devices; j++) {
            comms[i]->peer[j] = comms[j]->block;
            comms[i]->peerCap[j] = comms[j]->listCap;
        }
    cudaSetDevice(saved);
    if (st != PFAC_STATUS_SUCCESS)
        for (int i = 0; i < num_devices; i++) { if (comms[i]) PFAC_commDestroy(comms[i]); comms[i] = nullptr; }
    return st;
}

PFAC_status_t PFAC_commDestroy(PFAC_comm_t comm) {
    if (!comm) return PFAC_STATUS_INVALID_HANDLE;
    int saved = 0;
    cudaGetDevice(&saved);
    cudaSetDevice(comm->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < comm->world; r++)
        if (comm->ipcOpened[r]) cudaIpcCloseMemHandle(comm->peer[r]);
    cudaFree(comm->block);
    cudaFree(comm->d_scan);
    if (comm->h_scan) cudaFreeHost(comm->h_scan);
    cudaSetDevice(saved);
    delete comm;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_commGlobalList(PFAC_comm_t comm, int** d_ids, long long** d_pos, size_t* capacity) {
    if (!comm) return PFAC_STATUS_INVALID_HANDLE;
    if (d_ids) *d_ids = commListIds(comm->block);
    if (d_pos) *d_pos = commListPos(comm->block, comm->listCap);
    if (capacity) *capacity = comm->listCap;
    return PFAC_STATUS_SUCCESS;
}

// host copy of entries [first, first + n) of this rank's list region (synchronous)
PFAC_status_t PFAC_commReadGlobalList(PFAC_comm_t comm, size_t first, size_t n, int* h_ids, long long* h_pos) {
    if (!comm) return PFAC_STATUS_INVALID_HANDLE;
    if (first + n > comm->listCap || (n && (!h_ids || !h_pos))) return PFAC_STATUS_INVALID_PARAMETER;
    if (n == 0) return PFAC_STATUS_SUCCESS;
    int saved = 0;
    cudaGetDevice(&saved);
    cudaSetDevice(comm->device);
    cudaError_t e = cudaMemcpy(h_ids, commListIds(comm->block) + first, n * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        e = cudaMemcpy(h_pos, commListPos(comm->block, comm->listCap) + first, n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaSetDevice(saved);
    return cudaToStatus(e);
}

// PFAC_matchShardFromDeviceReduce64 + the exclusive scan of the ranks' counts, one kernel.  Collective:
// every rank of the comm calls it, in the same order.  d_scan (device, 3 words, may be NULL) receives
// {this rank's offset into the global list, total matches, this rank's count}.  h_scan == NULL: nothing
// is read back and the call does not synchronise (the scan stays on the device for the kernels that
// follow on the handle's stream); else one stream synchronisation and the same three words on the host.
PFAC_status_t PFAC_matchShardFromDeviceReduce64Global(PFAC_handle_t handle, PFAC_comm_t comm, const char* d_in,
                                                      size_t n_owned, size_t n_total, long long pos_base, int* d_id,
                                                      long long* d_pos, size_t capacity, unsigned long long* d_scan,
                                                      unsigned long long* h_scan) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!comm) return PFAC_STATUS_INVALID_HANDLE;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (n_total < n_owned) return PFAC_STATUS_INVALID_PARAMETER;
    if (n_owned && (!d_in || !d_pos || !d_id)) return PFAC_STATUS_INVALID_PARAMETER;
    if (comm->device != handle->device) return PFAC_STATUS_INVALID_PARAMETER;
    for (int r = 0; r < comm->world; r++) if (!comm->peer[r]) return PFAC_STATUS_INVALID_PARAMETER;  // not connected
    std::lock_guard<std::mutex> lock(handle->mu);
    pfac::CommLaunch cl;
    for (int r = 0; r < comm->world; r++) cl.peer[r] = reinterpret_cast<unsigned long long*>(comm->peer[r]);
    cl.scan = d_scan ? d_scan : comm->d_scan;
    cl.world = comm->world;
    cl.rank = comm->rank;
    cl.epoch = ++comm->epoch;
    PFAC_status_t st = reduceShardEnqueue(handle, reinterpret_cast<const unsigned char*>(d_in), n_owned, n_total, pos_base,
                                          d_id, d_pos, true, handle->stream, &cl, capacity);
    if (st != PFAC_STATUS_SUCCESS) return st;
    if (h_scan) {
        if (cudaMemcpyAsync(comm->h_scan, cl.scan, 24, cudaMemcpyDeviceToHost, handle->stream) != cudaSuccess)
            return PFAC_STATUS_INTERNAL_ERROR;
        if (cudaStreamSynchronize(handle->stream) != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
        h_scan[0] = comm->h_scan[0];
        h_scan[1] = comm->h_scan[1];
        h_scan[2] = comm->h_scan[2];
        if (h_scan[2] > capacity) return PFAC_STATUS_INVALID_PARAMETER;   // the run did not fit d_id / d_pos
    }
    return PFAC_STATUS_SUCCESS;
}

// Optional second step: one global list on rank dst_rank.  Collective.  Every rank copies its run
// d_id / d_pos [0, count) to the list region of dst_rank's comm block at its scanned offset (count and
// offset are read from d_scan on the device) with stores to peer memory, then raises its "placed"
// word there; dst_rank waits on the device until all ranks have.  Enqueued on the handle's stream;
// synchronize != 0 waits for the stream.  The list is PFAC_commGlobalList of dst_rank, [0, total).
PFAC_status_t PFAC_commGatherRuns(PFAC_handle_t handle, PFAC_comm_t comm, int dst_rank, const int* d_id,
                                  const long long* d_pos, const unsigned long long* d_scan, int synchronize) {
    if (!handle || !comm) return PFAC_STATUS_INVALID_HANDLE;
    if (dst_rank < 0 || dst_rank >= comm->world || !d_id || !d_pos) return PFAC_STATUS_INVALID_PARAMETER;
    if (comm->device != handle->device) return PFAC_STATUS_INVALID_PARAMETER;
    for (int r = 0; r < comm->world; r++) if (!comm->peer[r]) return PFAC_STATUS_INVALID_PARAMETER;
    const unsigned ep = ++comm->placeEpoch;
    const int par = int(ep & 1u) * pfac::kKernelCommMaxRanks;
    unsigned char* dst = comm->peer[dst_rank];
    unsigned long long* dstWords = reinterpret_cast<unsigned long long*>(dst);
    unsigned long long* ownWords = reinterpret_cast<unsigned long long*>(comm->block);
    cudaError_t e = pfac::launchPlaceRun(d_id, d_pos, d_scan ? d_scan : comm->d_scan, commListIds(dst),
                                         commListPos(dst, comm->peerCap[dst_rank]), comm->peerCap[dst_rank],
                                         dstWords + kCommPlacedWord + par + comm->rank, ownWords + kCommTicketWord, ep,
                                         handle->launch.numSMs, handle->stream);
    if (e == cudaSuccess && comm->rank == dst_rank)
        e = pfac::launchWaitPlaced(ownWords + kCommPlacedWord + par, comm->world, ep, handle->stream);
    if (e != cudaSuccess) return cudaToStatus(e);
    if (synchronize && cudaStreamSynchronize(handle->stream) != cudaSuccess) return PFAC_STATUS_INTERNAL_ERROR;
    return PFAC_STATUS_SUCCESS;
}

// ---- multi-GPU driver (one process, one host thread per GPU) ---------------------------------------
// What reference test/omp_PFAC.cpp:257-394 builds by hand: one handle per GPU, contiguous shards with
// a tail halo of maxPatternLen-1 bytes, results stitched on the host.  For the reduced form the
// per-GPU lists are written at provisional offsets and moved down to the exclusive scan of the
// per-GPU counts (the one-process counterpart of the NCCL count all-gather in pfac_b200/sharding.py).
struct PFAC_mgpu {
    std::vector<int> devices;
    std::vector<PFAC_handle_t> handles;
};

static bool mgpuReady(const PFAC_mgpu* mg) {  // a load that failed on one GPU leaves that handle without tables
    if (mg->handles.empty()) return false;
    for (PFAC_handle_t h : mg->handles)
        if (!h->patternsReady) return false;
    return true;
}

static void shardBounds(size_t size, int world, int rank, size_t halo, size_t* start, size_t* owned, size_t* total) {
    size_t per = (size + size_t(world) - 1) / size_t(world);
    per = (per + 4095) / 4096 * 4096;
    const size_t s = std::min(size_t(rank) * per, size);
    const size_t e = std::min(s + per, size);
    *start = s;
    *owned = e - s;
    *total = std::min(e + halo, size) - s;
}

PFAC_status_t PFAC_mgpuCreate(PFAC_mgpu_t* mg, const int* devices, int num_devices) {
    if (!mg || !devices || num_devices <= 0) return PFAC_STATUS_INVALID_PARAMETER;
    *mg = nullptr;
    int saved = 0;
    cudaGetDevice(&saved);
    PFAC_mgpu* m = new (std::nothrow) PFAC_mgpu();
    if (!m) return PFAC_STATUS_ALLOC_FAILED;
    PFAC_status_t st = PFAC_STATUS_SUCCESS;
    for (int i = 0; i < num_devices && st == PFAC_STATUS_SUCCESS; i++) {
        cudaError_t e = cudaSetDevice(devices[i]);
        if (e != cudaSuccess) { st = PFAC_status_t(e); break; }
        PFAC_handle_t h = nullptr;
        st = PFAC_create(&h);
        if (st == PFAC_STATUS_SUCCESS) { m->devices.push_back(devices[i]); m->handles.push_back(h); }
    }
    cudaSetDevice(saved);
    if (st != PFAC_STATUS_SUCCESS) { PFAC_mgpuDestroy(m); return st; }
    *mg = m;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_mgpuDestroy(PFAC_mgpu_t mg) {
    if (!mg) return PFAC_STATUS_INVALID_HANDLE;
    int saved = 0;
    cudaGetDevice(&saved);
    for (size_t i = 0; i < mg->handles.size(); i++) {
        cudaSetDevice(mg->devices[i]);
        PFAC_destroy(mg->handles[i]);
    }
    cudaSetDevice(saved);
    delete mg;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_mgpuReadPatternFromFile(PFAC_mgpu_t mg, char* filename) {
    if (!mg) return PFAC_STATUS_INVALID_HANDLE;
    int saved = 0;
    cudaGetDevice(&saved);
    PFAC_status_t st = PFAC_STATUS_SUCCESS;
    for (size_t i = 0; i < mg->handles.size() && st == PFAC_STATUS_SUCCESS; i++) {
        cudaSetDevice(mg->devices[i]);
        st = PFAC_readPatternFromFile(mg->handles[i], filename);
    }
    if (st != PFAC_STATUS_SUCCESS)   // all or nothing: no GPU keeps a dictionary the others lack
        for (size_t i = 0; i < mg->handles.size(); i++) {
            cudaSetDevice(mg->devices[i]);
            std::lock_guard<std::mutex> lock(mg->handles[i]->mu);
            if (mg->handles[i]->patternsReady) cudaDeviceSynchronize();
            freePatterns(mg->handles[i]);
        }
    cudaSetDevice(saved);
    return st;
}

PFAC_status_t PFAC_mgpuMatchFromHost(PFAC_mgpu_t mg, char* h_in, size_t size, int* h_out) {
    if (!mg) return PFAC_STATUS_INVALID_HANDLE;
    if (!h_in || !h_out) return PFAC_STATUS_INVALID_PARAMETER;
    if (!mgpuReady(mg)) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (size == 0) return PFAC_STATUS_SUCCESS;
    const int G = int(mg->handles.size());
    const size_t halo = size_t(std::max(mg->handles[0]->machine.maxPatternLen - 1, 0));
    std::vector<PFAC_status_t> status(size_t(G), PFAC_STATUS_SUCCESS);
    std::vector<std::thread> workers;
    for (int g = 0; g < G; g++) {
        workers.emplace_back([&, g]() {
            size_t s, owned, total;
            shardBounds(size, G, g, halo, &s, &owned, &total);
            if (owned == 0) return;
            if (cudaSetDevice(mg->devices[size_t(g)]) != cudaSuccess) { status[size_t(g)] = PFAC_STATUS_INTERNAL_ERROR; return; }
            status[size_t(g)] = hostDenseShard(mg->handles[size_t(g)], h_in + s, owned, total, h_out + s);
        });
    }
    for (std::thread& w : workers) w.join();
    for (PFAC_status_t st : status) if (st != PFAC_STATUS_SUCCESS) return st;
    return PFAC_STATUS_SUCCESS;
}

// h_id / h_pos must hold `size` entries (as the reference requires of its reduce buffers, user guide
// r1.2 p.30): each GPU first writes its run at its shard's start index.
PFAC_status_t PFAC_mgpuMatchFromHostReduce64(PFAC_mgpu_t mg, char* h_in, size_t size, int* h_id, long long* h_pos,
                                             unsigned long long* h_num) {
    if (!mg) return PFAC_STATUS_INVALID_HANDLE;
    if (!h_in || !h_id || !h_pos || !h_num) return PFAC_STATUS_INVALID_PARAMETER;
    if (!mgpuReady(mg)) return PFAC_STATUS_PATTERNS_NOT_READY;
    *h_num = 0;
    if (size == 0) return PFAC_STATUS_SUCCESS;
    const int G = int(mg->handles.size());
    const size_t halo = size_t(std::max(mg->handles[0]->machine.maxPatternLen - 1, 0));
    std::vector<PFAC_status_t> status(size_t(G), PFAC_STATUS_SUCCESS);
    std::vector<unsigned long long> counts(size_t(G), 0);
    std::vector<size_t> starts(size_t(G), 0);
    std::vector<std::thread> workers;
    for (int g = 0; g < G; g++) {
        workers.emplace_back([&, g]() {
            size_t s, owned, total;
            shardBounds(size, G, g, halo, &s, &owned, &total);
            starts[size_t(g)] = s;
            if (owned == 0) return;
            if (cudaSetDevice(mg->devices[size_t(g)]) != cudaSuccess) { status[size_t(g)] = PFAC_STATUS_INTERNAL_ERROR; return; }
            status[size_t(g)] = hostReduceShard(mg->handles[size_t(g)], h_in + s, owned, total, (long long)s, h_id + s,
                                                h_pos + s, true, &counts[size_t(g)]);
        });
    }
    for (std::thread& w : workers) w.join();
    for (PFAC_status_t st : status) if (st != PFAC_STATUS_SUCCESS) return st;
    // exclusive scan of the per-GPU counts = each run's place in the global list; runs only move down
    unsigned long long off = 0;
    for (int g = 0; g < G; g++) {
        const unsigned long long c = counts[size_t(g)];
        if (c && off != starts[size_t(g)]) {
            memmove(h_id + off, h_id + starts[size_t(g)], c * sizeof(int));
            memmove(h_pos + off, h_pos + starts[size_t(g)], c * sizeof(long long));
        }
        off += c;
    }
    *h_num = off;
    return PFAC_STATUS_SUCCESS;
}

// ---- host-only table compiler --------------------------------------------------------------
struct PFAC_table {
    pfac::Machine machine;
    pfac::DeviceLayout layout;
    size_t budget = 0;   // what the layout was compiled for (stored in compiled-table files)
    int policy = pfac::kFilterAuto;
};

static void fillInfo(const pfac::Machine& m, const pfac::DeviceLayout& L, PFAC_tableInfo_t* info) {
    info->num_patterns = m.numPatterns;
    info->num_states = m.numStates;
    info->initial_state = m.initialState;
    info->max_pattern_len = m.maxPatternLen;
    info->num_leaves = m.numLeaves;
    info->num_edges = L.numEdges;
    info->hash_edges = L.hashEdges;
    info->num_chains = L.numChains;
    info->tail_bytes = int(L.tails.size());
    info->chains_hot = L.chainsHot ? 1 : 0;
    info->next2_hot = L.next2Hot ? 1 : 0;
    info->code_bits = L.codeBits;
    info->gram_len = L.gramLen;
    info->has_best2 = L.best2.empty() ? 0 : 1;
    info->has_chk2 = L.chk2.empty() ? 0 : 1;
    info->max_depth = L.maxDepth;
    info->hot_depth = L.hotDepth;
    info->hot_buckets = L.hotBuckets;
    info->cold_buckets = L.coldBuckets;
    info->hash_mul = L.mul;
    info->hot_max_probe = L.hotMaxProbe;
    info->cold_max_probe = L.coldMaxProbe;
    info->pre2_bits_set = L.pre2BitsSet;
    info->root_fanout = L.rootFanout;
    info->hashed_filter = L.hfilt.empty() ? 0 : L.hfiltK;
    info->hfilt_bits_set = L.hfiltBitsSet;
    info->code_shift = L.codeShift;
    info->device_bytes = L.deviceBytes();
    info->hfilt_words = unsigned(L.hfilt.size());
}

PFAC_status_t PFAC_tableCompile(const char* image, size_t size, size_t hot_budget_bytes, PFAC_table_t* table) {
    if (!table) return PFAC_STATUS_INVALID_PARAMETER;
    *table = nullptr;
    PFAC_table* t = new (std::nothrow) PFAC_table();
    if (!t) return PFAC_STATUS_ALLOC_FAILED;
    int st;
    try {
        st = pfac::buildMachine(image, size, t->machine);
        if (st == PFAC_STATUS_SUCCESS) {
            t->budget = hot_budget_bytes;
            t->policy = filterPolicy();
            pfac::compileLayout(t->machine, hot_budget_bytes, t->layout, t->policy);
        }
    } catch (...) {
        st = PFAC_STATUS_ALLOC_FAILED;
    }
    if (st != PFAC_STATUS_SUCCESS) { delete t; return PFAC_status_t(st); }
    *table = t;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_tableCompileArrays(const char* const* patterns, const size_t* lengths, size_t num_patterns,
                                      size_t hot_budget_bytes, PFAC_table_t* table) {
    if (!table) return PFAC_STATUS_INVALID_PARAMETER;
    *table = nullptr;
    PFAC_table* t = new (std::nothrow) PFAC_table();
    if (!t) return PFAC_STATUS_ALLOC_FAILED;
    int st;
    try {
        st = pfac::buildMachineFromArrays(patterns, lengths, num_patterns, t->machine);
        if (st == PFAC_STATUS_SUCCESS) {
            t->budget = hot_budget_bytes;
            t->policy = filterPolicy();
            pfac::compileLayout(t->machine, hot_budget_bytes, t->layout, t->policy);
        }
    } catch (...) {
        st = PFAC_STATUS_ALLOC_FAILED;
    }
    if (st != PFAC_STATUS_SUCCESS) { delete t; return PFAC_status_t(st); }
    *table = t;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_tableCompileFile(const char* filename, size_t hot_budget_bytes, PFAC_table_t* table) {
    if (!filename || !table) return PFAC_STATUS_INVALID_PARAMETER;
    std::string image;
    PFAC_status_t st = readWholeFile(filename, image);
    if (st != PFAC_STATUS_SUCCESS) return st;
    return PFAC_tableCompile(image.data(), image.size(), hot_budget_bytes, table);
}

// ---- compiled-table files (pfac_table.h saveCompiled / loadCompiled) ------------------------------
PFAC_status_t PFAC_tableSave(PFAC_table_t table, const char* filename) {
    if (!table) return PFAC_STATUS_INVALID_HANDLE;
    if (!filename) return PFAC_STATUS_INVALID_PARAMETER;
    pfac::CompiledLayout c;
    c.budget = table->budget;
    c.policy = table->policy;
    c.layout = table->layout;
    return pfac::saveCompiled(filename, table->machine, {&c}) ? PFAC_STATUS_SUCCESS : PFAC_STATUS_FILE_OPEN_ERROR;
}

PFAC_status_t PFAC_tableLoad(const char* filename, PFAC_table_t* table) {
    if (!filename || !table) return PFAC_STATUS_INVALID_PARAMETER;
    *table = nullptr;
    FILE* fp = fopen(filename, "rb");
    if (!fp) return PFAC_STATUS_FILE_OPEN_ERROR;
    fclose(fp);
    PFAC_table* t = new (std::nothrow) PFAC_table();
    if (!t) return PFAC_STATUS_ALLOC_FAILED;
    std::vector<pfac::CompiledLayout> ls;
    if (!pfac::loadCompiled(filename, t->machine, ls) || ls.empty()) { delete t; return PFAC_STATUS_INVALID_PARAMETER; }
    t->layout = std::move(ls[0].layout);
    t->budget = size_t(ls[0].budget);
    t->policy = ls[0].policy;
    *table = t;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_saveCompiledPatterns(PFAC_handle_t handle, const char* filename) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    if (!filename) return PFAC_STATUS_INVALID_PARAMETER;
    std::lock_guard<std::mutex> lock(handle->mu);
    pfac::CompiledLayout a, b;
    a.budget = hotBudget(handle, false);
    b.budget = hotBudget(handle, true);
    a.policy = filterPolicyFor(false);
    b.policy = filterPolicyFor(true);
    a.layout = handle->layout;
    b.layout = handle->layoutReduce;
    return pfac::saveCompiled(filename, handle->machine, {&a, &b}) ? PFAC_STATUS_SUCCESS
                                                                    : PFAC_STATUS_FILE_OPEN_ERROR;
}

// Same end state as PFAC_readPatternFromFile on the pattern file the image was compiled from.  A
// layout stored for another shared-memory budget or filter policy is not used: that one is
// recompiled from the stored automaton.
PFAC_status_t PFAC_loadCompiledPatterns(PFAC_handle_t handle, const char* filename) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!filename) return PFAC_STATUS_INVALID_PARAMETER;
    FILE* fp = fopen(filename, "rb");
    if (!fp) return PFAC_STATUS_FILE_OPEN_ERROR;
    fclose(fp);
    pfac::Machine m;
    std::vector<pfac::CompiledLayout> ls;
    if (!pfac::loadCompiled(filename, m, ls)) return PFAC_STATUS_INVALID_PARAMETER;
    std::lock_guard<std::mutex> lock(handle->mu);
    if (handle->patternsReady) {
        cudaDeviceSynchronize();
        freePatterns(handle);
    }
    handle->patternFile[0] = 0;
    handle->machine = std::move(m);
    handle->patternsReady = true;
    auto take = [&](bool reduceKernel, pfac::DeviceLayout& dst) {
        const size_t budget = hotBudget(handle, reduceKernel);
        const int policy = filterPolicyFor(reduceKernel);
        for (pfac::CompiledLayout& c : ls)
            if (c.budget == budget && c.policy == policy && !c.layout.pre2.empty()) {
                dst = c.layout;
                return;
            }
        pfac::compileLayout(handle->machine, budget, dst, policy);
    };
    take(false, handle->layout);
    take(true, handle->layoutReduce);
    PFAC_status_t st = uploadTables(handle);
    if (st != PFAC_STATUS_SUCCESS) { freePatterns(handle); return st; }
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_tableDestroy(PFAC_table_t table) {
    if (!table) return PFAC_STATUS_INVALID_HANDLE;
    delete table;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_tableDump(PFAC_table_t table, FILE* fp) {
    if (!table) return PFAC_STATUS_INVALID_HANDLE;
    pfac::dumpMachine(table->machine, fp ? fp : stdout);
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_tableDumpToFile(PFAC_table_t table, const char* filename) {
    if (!table) return PFAC_STATUS_INVALID_HANDLE;
    if (!filename) return PFAC_STATUS_INVALID_PARAMETER;
    FILE* fp = fopen(filename, "w");
    if (!fp) return PFAC_STATUS_FILE_OPEN_ERROR;
    pfac::dumpMachine(table->machine, fp);
    fclose(fp);
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_tableGetInfo(PFAC_table_t table, PFAC_tableInfo_t* info) {
    if (!table) return PFAC_STATUS_INVALID_HANDLE;
    if (!info) return PFAC_STATUS_INVALID_PARAMETER;
    fillInfo(table->machine, table->layout, info);
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_tableGetLayout(PFAC_table_t table, const int** root, const unsigned** pre2,
                                  const unsigned short** rank2, const unsigned** next2,
                                  const unsigned** hot, const unsigned** cold, const unsigned** chains,
                                  const unsigned char** tails) {
    if (!table) return PFAC_STATUS_INVALID_HANDLE;
    if (root) *root = table->layout.root;
    if (pre2) *pre2 = table->layout.pre2.data();
    if (rank2) *rank2 = table->layout.rank2.data();
    if (next2) *next2 = table->layout.next2.data();
    if (hot) *hot = table->layout.hot.data();
    if (cold) *cold = table->layout.cold.data();
    if (chains) *chains = table->layout.chains.data();
    if (tails) *tails = table->layout.tails.data();
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_tableGetFilter(PFAC_table_t table, const unsigned** hfilt) {
    if (!table) return PFAC_STATUS_INVALID_HANDLE;
    if (!hfilt) return PFAC_STATUS_INVALID_PARAMETER;
    *hfilt = table->layout.hfilt.empty() ? nullptr : table->layout.hfilt.data();
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_tableGetLayout2(PFAC_table_t table, const unsigned char** lut, const unsigned** best2,
                                   const unsigned short** chk2) {
    if (!table) return PFAC_STATUS_INVALID_HANDLE;
    if (lut) *lut = table->layout.lut;
    if (best2) *best2 = table->layout.best2.empty() ? nullptr : table->layout.best2.data();
    if (chk2) *chk2 = table->layout.chk2.empty() ? nullptr : table->layout.chk2.data();
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_getTableInfo(PFAC_handle_t handle, PFAC_tableInfo_t* info) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!info) return PFAC_STATUS_INVALID_PARAMETER;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    fillInfo(handle->machine, handle->layout, info);
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_getTableInfoReduce(PFAC_handle_t handle, PFAC_tableInfo_t* info) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    if (!info) return PFAC_STATUS_INVALID_PARAMETER;
    if (!handle->patternsReady) return PFAC_STATUS_PATTERNS_NOT_READY;
    fillInfo(handle->machine, handle->layoutReduce, info);
    return PFAC_STATUS_SUCCESS;
}

// bench.py reports how many of this library's kernels ran inside its timed region
PFAC_status_t PFAC_hostCopy(void* dst, const void* src, size_t bytes) {
    if (bytes && (!dst || !src)) return PFAC_STATUS_INVALID_PARAMETER;
    acquireCopyPool()->copy(dst, src, bytes, true);  // destination written with streaming stores
    return PFAC_STATUS_SUCCESS;
}

// The matchFromHost* pipelines keep their device and pinned buffers between calls (the reference
// allocates and frees per call, PFAC.cpp:915-961); a caller that is done with host calls can give
// them back without destroying the handle.
PFAC_status_t PFAC_releaseHostBuffers(PFAC_handle_t handle) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    std::lock_guard<std::mutex> lock(handle->pipeMu);
    cudaDeviceSynchronize();
    freePipe(handle->pipe);
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_lastHostTransfer(PFAC_handle_t handle, size_t* h2d_bytes, size_t* d2h_bytes) {
    if (!handle) return PFAC_STATUS_INVALID_HANDLE;
    std::lock_guard<std::mutex> lock(handle->pipeMu);
    if (h2d_bytes) *h2d_bytes = handle->pipe.lastH2D;
    if (d2h_bytes) *d2h_bytes = handle->pipe.lastD2H;
    return PFAC_STATUS_SUCCESS;
}

PFAC_status_t PFAC_hostZero(void* dst, size_t bytes) {
    if (bytes && !dst) return PFAC_STATUS_INVALID_PARAMETER;
    std::shared_ptr<CopyPool> pool = acquireCopyPool();
    CopyPool::Job job;
    pool->startZero(job, dst, bytes);
    pool->finish(job);
    return PFAC_STATUS_SUCCESS;
}

unsigned long long PFAC_kernelLaunchCount(void) { return pfac::kernelLaunchCount(); }

}  // extern "C"
#pragma GCC visibility pop
