// hbm_mix.cu -- what HBM delivers for the dense kernel's traffic mix (1 byte read : 4 bytes written),
// next to a pure fill and a 1:1 copy, on the sizes of config C2 (1 GiB in, 4 GiB out).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/hbm_mix tools/micro/hbm_mix.cu && build/hbm_mix
// Prints one JSON line.  Plain grid-stride kernels with 16-byte accesses: no claim that these are the
// best possible, only a reference point measured with the simplest code that moves the same bytes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void fill_kernel(uint4* out, size_t n16) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = make_uint4(0u, 0u, 0u, 0u);
}
// every thread reads 16 bytes of input and writes the 64 bytes of "results" that belong to them
__global__ void mix_kernel(const uint4* __restrict__ in, uint4* out, size_t n16, unsigned* sink) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    unsigned acc = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = in[i];
        acc |= v.x & v.y & v.z & v.w;
#pragma unroll
        for (int k = 0; k < 4; k++) out[4 * i + k] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (acc == 0x12345678u) *sink = acc;
}
// the same with warp-contiguous 512-byte output rows (lane l writes 16 bytes of each of four rows)
__global__ void mix_rows_kernel(const uint4* __restrict__ in, uint4* out, size_t n16, unsigned* sink) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    unsigned acc = 0;
    const int lane = threadIdx.x & 31;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = in[i];
        acc |= v.x & v.y & v.z & v.w;
        const size_t base = (i - lane) * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) out[base + 32 * k + lane] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (acc == 0x12345678u) *sink = acc;
}
__device__ __forceinline__ void st256_zero(void* p) {
    unsigned z = 0;
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(z) : "memory");
}
// fill with 32-byte stores (sm_100: STG.E.ENL2.256): a warp instruction writes 1 KB
__global__ void fill256_kernel(unsigned char* out, size_t n32) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n32; i += stride) st256_zero(out + 32 * i);
}
// the dense kernel's shape: a warp reads 1,536 + 64 bytes and writes 6 KB of zeros (twelve 16-byte or six 32-byte
// stores per lane), warps of a CTA on consecutive tiles, tiles advancing by the grid
template <bool V8>
__global__ void mix_tiles_kernel(const unsigned char* __restrict__ in, unsigned char* out, size_t ntiles, unsigned* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    unsigned acc = 0;
    for (size_t t = size_t(blockIdx.x) * wpc + warp; t < ntiles; t += size_t(gridDim.x) * wpc) {
        unsigned char* o = out + t * 6144;
        if (V8) {
#pragma unroll
            for (int k = 0; k < 6; k++) st256_zero(o + (k * 32 + lane) * 32);
        } else {
#pragma unroll
            for (int k = 0; k < 12; k++) reinterpret_cast<uint4*>(o)[k * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
        }
        const uint4* i4 = reinterpret_cast<const uint4*>(in + t * 1536);
#pragma unroll
        for (int k = 0; k < 3; k++) { const uint4 v = i4[k * 32 + lane]; acc |= v.x & v.y & v.z & v.w; }
    }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void copy_kernel(const uint4* __restrict__ in, uint4* out, size_t n16) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = in[i];
}

template <typename F> float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; i++) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; i++) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main() {
    const size_t N = size_t(1) << 30;
    unsigned char *in, *out;
    unsigned* sink;
    CK(cudaMalloc(&in, N + 4 * N));   // copy test: 2.5 GiB -> 2.5 GiB inside the same buffers
    CK(cudaMalloc(&out, 4 * N + N));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 1, 5 * N));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 8, block = 512;
    const int reps = 20;
    const float fill = time_ms([&] { fill_kernel<<<grid, block>>>((uint4*)out, 4 * N / 16); }, reps);
    const float mset = time_ms([&] { cudaMemsetAsync(out, 0, 4 * N); }, reps);
    const float mix = time_ms([&] { mix_kernel<<<grid, block>>>((const uint4*)in, (uint4*)out, N / 16, sink); }, reps);
    const float mixr = time_ms([&] { mix_rows_kernel<<<grid, block>>>((const uint4*)in, (uint4*)out, N / 16, sink); }, reps);
    const float fill256 = time_ms([&] { fill256_kernel<<<grid, block>>>(out, 4 * N / 32); }, reps);
    const size_t ntiles = N / 1536;
    const float mixt = time_ms([&] { mix_tiles_kernel<false><<<sms, 1024>>>(in, out, ntiles, sink); }, reps);
    const float mixt8 = time_ms([&] { mix_tiles_kernel<true><<<sms, 1024>>>(in, out, ntiles, sink); }, reps);
    const float mixt8b = time_ms([&] { mix_tiles_kernel<true><<<sms * 2, 1024>>>(in, out, ntiles, sink); }, reps);
    const float copy = time_ms([&] { copy_kernel<<<grid, block>>>((const uint4*)in, (uint4*)out, (5 * N / 2) / 16); }, reps);
    const float mcpy = time_ms([&] { cudaMemcpyAsync(out, in, 5 * N / 2, cudaMemcpyDeviceToDevice); }, reps);
    CK(cudaGetLastError());
    auto gbs = [](double bytes, float ms) { return bytes / (ms * 1e-3) / 1e9; };
    printf("{\"fill256_4GiB_ms\": %.4f, \"fill256_GBps\": %.1f, \"mix_tiles_ms\": %.4f, \"mix_tiles_GBps\": %.1f, "
           "\"mix_tiles_st256_ms\": %.4f, \"mix_tiles_st256_GBps\": %.1f, \"mix_tiles_st256_2cta_ms\": %.4f, ",
           fill256, gbs(4.0 * N, fill256), mixt, gbs(5.0 * N, mixt), mixt8, gbs(5.0 * N, mixt8), mixt8b);
    printf("\"fill_4GiB_ms\": %.4f, \"fill_GBps\": %.1f, \"memset_4GiB_ms\": %.4f, \"memset_GBps\": %.1f, "
           "\"mix_1r4w_ms\": %.4f, \"mix_GBps\": %.1f, \"mix_rows_ms\": %.4f, \"mix_rows_GBps\": %.1f, "
           "\"copy_2.5GiB_ms\": %.4f, \"copy_GBps\": %.1f, \"memcpy_2.5GiB_ms\": %.4f, \"memcpy_GBps\": %.1f}\n",
           fill, gbs(4.0 * N, fill), mset, gbs(4.0 * N, mset), mix, gbs(5.0 * N, mix), mixr, gbs(5.0 * N, mixr),
           copy, gbs(5.0 * N, copy), mcpy, gbs(5.0 * N, mcpy));
    return 0;
}
