// devgen.cu -- the synthetic workloads of workloads/synth.py regenerated on the GPU (test and bench
// tooling; not part of libpfac.so).  Text byte i is a pure function of (seed, i), and every plant a
// pure function of its block / boundary index, so a rank materialises its own shard of a 32 GiB
// stream in milliseconds instead of minutes of numpy.
// Must stay byte-equal to synth.random_bytes / ascii_weighted_bytes / dna_bytes / plant
// (tests/test_devgen.py compares windows on the GPU box).
#include <cstdint>

#include <cuda_runtime.h>

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t x) {  // synth._splitmix64 / _mix_scalar
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// kind 0: uniform bytes; 1: half forced into printable ASCII; 2: ACGT.  One thread per 8-byte word
// of the stream (word index = absolute position / 8), clipped to [start, start + n).
__global__ void text_kernel(int kind, uint64_t seed, uint64_t start, uint64_t n, unsigned char* out) {
    const uint64_t w0 = start / 8;
    const uint64_t nwords = (start + n + 7) / 8 - w0;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < nwords;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t w = w0 + i;
        const uint64_t raw = mix64(w + seed);
        uint64_t val = raw;
        if (kind == 1) {
            const uint64_t sel = mix64(w + (seed ^ 0xA5A5A5A5ull));
            val = 0;
            for (int k = 0; k < 8; k++) {
                const uint32_t r = uint32_t(raw >> (8 * k)) & 0xFFu;
                const uint32_t s = uint32_t(sel >> (8 * k)) & 0xFFu;
                const uint32_t printable = ((r * 95u) >> 8) + 0x20u;
                val |= uint64_t(s < 128u ? printable : r) << (8 * k);
            }
        } else if (kind == 2) {
            val = 0;
            for (int k = 0; k < 8; k++) {
                const uint32_t r = uint32_t(raw >> (8 * k)) & 3u;
                val |= uint64_t("ACGT"[r]) << (8 * k);
            }
        }
        const uint64_t pos = w * 8;  // absolute position of byte 0 of this word
        if (pos >= start && pos + 8 <= start + n && ((reinterpret_cast<uintptr_t>(out) + (pos - start)) & 7) == 0) {
            *reinterpret_cast<uint64_t*>(out + (pos - start)) = val;
        } else {
            for (int k = 0; k < 8; k++) {
                const uint64_t q = pos + k;
                if (q >= start && q < start + n) out[q - start] = (unsigned char)(val >> (8 * k));
            }
        }
    }
}

struct PlantArgs {
    unsigned char* text;
    uint64_t start, n, total_len;
    const unsigned char* pat_bytes;
    const uint64_t* pat_off;  // P + 1
    uint64_t P;
    uint64_t seed, every, boundary;
    uint64_t first, count;    // index range of this kernel
};

__device__ inline void put(const PlantArgs& a, uint64_t abs_pos, const unsigned char* pat, uint64_t len) {
    const uint64_t end = a.start + a.n;
    const uint64_t lo = abs_pos > a.start ? abs_pos : a.start;
    uint64_t hi = abs_pos + len;
    if (hi > end) hi = end;
    if (hi > a.total_len) hi = a.total_len;
    for (uint64_t q = lo; q < hi; q++) a.text[q - a.start] = pat[q - abs_pos];
}

// one pattern per `every`-byte block at a hash-derived offset (synth.plant, first loop)
__global__ void plant_blocks_kernel(PlantArgs a) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < a.count;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t blk = a.first + i;
        const uint64_t h = mix64((a.seed ^ 0x1234567ull) + blk);
        const uint64_t pi = h % a.P;
        const uint64_t len = a.pat_off[pi + 1] - a.pat_off[pi];
        if (len >= a.every) continue;
        const uint64_t span = a.every - len;
        put(a, blk * a.every + (h >> 20) % span, a.pat_bytes + a.pat_off[pi], len);
    }
}

// one pattern straddling every `boundary` multiple (second loop)
__global__ void plant_boundaries_kernel(PlantArgs a) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < a.count;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t k = a.first + i;
        const uint64_t h = mix64((a.seed ^ 0x7654321ull) + k);
        const uint64_t pi = h % a.P;
        const uint64_t len = a.pat_off[pi + 1] - a.pat_off[pi];
        if (len < 2) continue;
        const uint64_t cut = 1 + (h >> 24) % (len - 1);
        put(a, k * a.boundary - cut, a.pat_bytes + a.pat_off[pi], len);
    }
}

// the stream ends in a proper prefix of the longest pattern (must not be reported)
__global__ void plant_tail_kernel(PlantArgs a, uint64_t longest) {
    const uint64_t len = a.pat_off[longest + 1] - a.pat_off[longest];
    if (threadIdx.x == 0 && blockIdx.x == 0 && len >= 2 && a.total_len >= len)
        put(a, a.total_len - (len - 1), a.pat_bytes + a.pat_off[longest], len - 1);
}

int gridFor(uint64_t items) {
    uint64_t g = (items + 255) / 256;
    if (g > 148ull * 16) g = 148ull * 16;
    return g ? int(g) : 1;
}

}  // namespace

extern "C" {

// d_out[0..n) = stream bytes [start, start + n)
int pfac_devgen_text(int kind, unsigned long long seed, unsigned long long start, unsigned long long n,
                     unsigned char* d_out, void* stream) {
    if (kind < 0 || kind > 2) return 1;
    if (n == 0) return 0;
    text_kernel<<<gridFor((n + 7) / 8 + 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(kind, seed, start, n, d_out);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// synth.plant on the device.  d_pat_off: P + 1 byte offsets into d_pat_bytes; maxlen, longest: as
// synth.plant computes them (longest = index of the first pattern of maximum length).
int pfac_devgen_plant(unsigned char* d_text, unsigned long long start, unsigned long long n,
                      unsigned long long total_len, const unsigned char* d_pat_bytes,
                      const unsigned long long* d_pat_off, unsigned long long P, unsigned long long maxlen,
                      unsigned long long longest, unsigned long long seed, unsigned long long every,
                      unsigned long long boundary, void* stream) {
    if (P == 0 || n == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PlantArgs a{d_text, start, n, total_len, d_pat_bytes, reinterpret_cast<const uint64_t*>(d_pat_off), P, seed, every,
                boundary, 0, 0};
    const uint64_t end = start + n;
    const uint64_t lim = end < total_len ? end : total_len;
    if (every) {
        const uint64_t first_blk = start > maxlen ? (start - maxlen) / every : 0;
        const uint64_t last_blk = (lim + every - 1) / every;
        if (last_blk > first_blk) {
            a.first = first_blk;
            a.count = last_blk - first_blk;
            plant_blocks_kernel<<<gridFor(a.count), 256, 0, s>>>(a);
        }
    }
    if (boundary) {
        const uint64_t first_b = start / boundary > 1 ? start / boundary : 1;
        const uint64_t stop = (lim + maxlen) / boundary + 1;
        if (stop > first_b) {
            a.first = first_b;
            a.count = stop - first_b;
            plant_boundaries_kernel<<<gridFor(a.count), 256, 0, s>>>(a);
        }
    }
    plant_tail_kernel<<<1, 32, 0, s>>>(a, longest);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

}  // extern "C"
