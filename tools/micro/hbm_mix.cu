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
    const float copy = time_ms([&] { copy_kernel<<<grid, block>>>((const uint4*)in, (uint4*)out, (5 * N / 2) / 16); }, reps);
    const float mcpy = time_ms([&] { cudaMemcpyAsync(out, in, 5 * N / 2, cudaMemcpyDeviceToDevice); }, reps);
    CK(cudaGetLastError());
    auto gbs = [](double bytes, float ms) { return bytes / (ms * 1e-3) / 1e9; };
    printf("{\"fill_4GiB_ms\": %.4f, \"fill_GBps\": %.1f, \"memset_4GiB_ms\": %.4f, \"memset_GBps\": %.1f, "
           "\"mix_1r4w_ms\": %.4f, \"mix_GBps\": %.1f, \"mix_rows_ms\": %.4f, \"mix_rows_GBps\": %.1f, "
           "\"copy_2.5GiB_ms\": %.4f, \"copy_GBps\": %.1f, \"memcpy_2.5GiB_ms\": %.4f, \"memcpy_GBps\": %.1f}\n",
           fill, gbs(4.0 * N, fill), mset, gbs(4.0 * N, mset), mix, gbs(5.0 * N, mix), mixr, gbs(5.0 * N, mixr),
           copy, gbs(5.0 * N, copy), mcpy, gbs(5.0 * N, mcpy));
    return 0;
}
