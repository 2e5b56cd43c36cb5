// pfac_kernels.h -- launch interface of the sm_100a matching kernels (pfac_kernels.cu).
#pragma once
#include <cstddef>
#include <cstdint>

#include <cuda_runtime.h>

namespace pfac {

// hashed 4-gram filter parameters the kernels are built with (pfac_api.cu checks them against
// the table compiler's, pfac_table.h)
constexpr uint32_t kKernelHashFilterMul = 0x9E3779B1u;
constexpr uint32_t kKernelHashFilterMul2 = 0x85EBCA6Bu;
constexpr uint32_t kKernelHashFilterMul3 = 0xC2B2AE35u;
constexpr int kKernelHashFilterWords = 8192;
constexpr int kKernelCommMaxRanks = 16;   // GPUs of one NVLink domain taking part in a count exchange

// Device-resident compiled table (uploaded by the handle).
struct DeviceTable {
    const int32_t* root = nullptr;    // 256
    const uint32_t* pre2 = nullptr;   // 2048
    const unsigned short* rank2 = nullptr;  // 2048
    const unsigned char* lut = nullptr;     // 256: symbol code | 0x80
    const uint32_t* next2 = nullptr;  // next2Bytes / 4 (padded to 16 bytes)
    const uint32_t* best2 = nullptr;  // parallel to next2 (valid when hasBest2)
    const unsigned short* chk2 = nullptr;  // second prefilter stage (chk2Bytes > 0), always staged in smem
    uint32_t chk2Bytes = 0;           // multiple of 16; 0 = stage off
    const uint32_t* hfilt = nullptr;  // hashed 4-gram first stage (hfiltBytes > 0), always staged in smem
    uint32_t hfiltBytes = 0;          // 0, 32 KB, or 64 KB (hfiltK == 2, byte alphabets: row-indexed)
    int hfiltK = 0;                   // bits tested per lookup (1 or 2)
    uint32_t next2Bytes = 0;
    bool next2Hot = false;            // kernels copy next2 (+ best2) into shared memory
    bool hasBest2 = false;
    int codeBits = 8;                 // bits per symbol of the prefilter index
    int gramLen = 2;                  // symbols covered by the prefilter
    int codeShift = -1;               // 2-bit alphabets with a hashed first stage: code = (byte >> codeShift) & 3
    const uint4* hot = nullptr;       // hotBuckets (copied into shared memory by each CTA)
    const uint4* cold = nullptr;      // coldBuckets (read through L1/L2)
    const uint4* chains = nullptr;    // chain records (16 B each)
    const unsigned char* tails = nullptr;
    uint32_t hotBuckets = 0;
    uint32_t coldBuckets = 1;
    uint32_t chainBytes = 0;          // multiple of 16
    uint32_t tailBytes = 0;           // multiple of 16
    bool chainsHot = false;           // kernels also copy chains + tails into shared memory
    uint32_t mul = 0;
    int hotDepth = 1;
    int numFinal = 0;
    int maxPatternLen = 0;
};

struct LaunchConfig {
    int numSMs = 148;
};

// Shared-memory bytes the dense (or reduce) kernel can spare for next2 / hash rows / chains /
// tails, given the longest pattern (which fixes the staged halo).  The table compiler is run
// once per kernel with its budget; the two layouts differ in what is marked hot and, for sparse
// dictionaries, in the first stage (pair filter: reduce kernel only).
size_t tableSmemBudget(int maxPatternLen, bool reduceKernel);

// Look-back descriptor words needed by the reduce kernel for an n_owned-byte shard.
size_t reduceWorkspaceWords(size_t n_owned);
// Words of the per-warp spill rings (matches that do not fit the shared-memory parking; about 75 MB,
// L2-resident in use); needs no initialisation.
size_t reduceParkWords(const LaunchConfig& cfg);

// Load the kernels the two tables run on now rather than at the first match call.
cudaError_t prepareKernels(const DeviceTable& dense, const DeviceTable& reduce);

// Dense result: out[i] for i in [0,n_owned); walks may read in[0,n_total).
cudaError_t launchMatchDense(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                             size_t n_owned, size_t n_total, int* out, cudaStream_t stream);

// Cross-GPU count exchange fused into the reduce kernel: every rank's mailbox (2 x kKernelCommMaxRanks
// words, peer-mapped), the call's epoch, and where {exclusive offset, total, own count} go.
struct CommLaunch {
    unsigned long long* peer[kKernelCommMaxRanks];
    unsigned long long* scan;
    int world = 0, rank = 0;
    unsigned epoch = 0;
};

// Fused match + ordered compaction.  desc: reduceWorkspaceWords(n_owned) zeroed uint64 words; park:
// reduceParkWords(cfg) uint64 words; dbg: 8 zeroed host-mapped words where a wait that never ends
// reports before the kernel traps (nullptr: trap without a report).  out_cap: entries the output
// arrays hold; matches beyond it are counted but not stored.
// d_total: one uint64 receiving the match count.  pos64 selects long long vs int positions.
cudaError_t launchMatchReduce(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                              size_t n_owned, size_t n_total, long long pos_base, int* out_id,
                              void* out_pos, bool pos64, unsigned long long* desc, unsigned long long* park,
                              unsigned long long* d_total, cudaStream_t stream, const CommLaunch* comm = nullptr,
                              unsigned long long* dbg = nullptr, unsigned long long out_cap = ~0ull);

// Copy this rank's run (count and offset read from `scan` on the device) into the list on the
// destination rank and raise `placed`; the destination waits for all ranks with launchWaitPlaced.
cudaError_t launchPlaceRun(const int* ids, const long long* pos, const unsigned long long* scan, int* g_ids,
                           long long* g_pos, unsigned long long capacity, unsigned long long* placed,
                           unsigned long long* ticket, unsigned epoch, int numSMs, cudaStream_t stream);
cudaError_t launchWaitPlaced(const unsigned long long* placed, int world, unsigned epoch, cudaStream_t stream);

// Number of kernels launched by this library since load (bench.py reports it).
unsigned long long kernelLaunchCount();

}  // namespace pfac
