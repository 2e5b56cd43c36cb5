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
    uint32_t hfiltBytes = 0;          // 0 or kHashFilterWords * 4
    int hfiltK = 0;                   // bits tested per lookup (1 or 2)
    uint32_t next2Bytes = 0;
    bool next2Hot = false;            // kernels copy next2 (+ best2) into shared memory
    bool hasBest2 = false;
    int codeBits = 8;                 // bits per symbol of the prefilter index
    int gramLen = 2;                  // symbols covered by the prefilter
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
// once per kernel with its budget; the two layouts differ only in what is marked hot.
size_t tableSmemBudget(int maxPatternLen, bool reduceKernel);

// Look-back descriptor words needed by the reduce kernel for an n_owned-byte shard.
size_t reduceWorkspaceWords(size_t n_owned);

// Dense result: out[i] for i in [0,n_owned); walks may read in[0,n_total).
cudaError_t launchMatchDense(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                             size_t n_owned, size_t n_total, int* out, cudaStream_t stream);

// Fused match + ordered compaction.  desc: reduceWorkspaceWords(n_owned) zeroed uint64 words.
// d_total: one uint64 receiving the match count.  pos64 selects long long vs int positions.
cudaError_t launchMatchReduce(const DeviceTable& t, const LaunchConfig& cfg, const unsigned char* in,
                              size_t n_owned, size_t n_total, long long pos_base, int* out_id,
                              void* out_pos, bool pos64, unsigned long long* desc,
                              unsigned long long* d_total, cudaStream_t stream);

// Number of kernels launched by this library since load (bench.py reports it).
unsigned long long kernelLaunchCount();

}  // namespace pfac
