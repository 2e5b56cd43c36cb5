// fill_sweep.cu -- which plain fill pattern reaches cudaMemset's rate on this GPU?  (tools/micro/hbm_mix.cu's follow-up)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/fill_sweep tools/micro/fill_sweep.cu && build/fill_sweep
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ void st32(unsigned char* p) {   // 32 bytes of zeros
    unsigned z = 0;
    if (MODE == 0) { reinterpret_cast<uint4*>(p)[0] = make_uint4(0, 0, 0, 0); reinterpret_cast<uint4*>(p)[1] = make_uint4(0, 0, 0, 0); }
    if (MODE == 1) asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(z) : "memory");
    if (MODE == 2) asm volatile("st.global.cs.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(z) : "memory");
    if (MODE == 3) asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(z) : "memory");
}
// CHUNK: every CTA owns one contiguous piece; else grid-stride
template <int MODE, bool CHUNK>
__global__ void fill(unsigned char* out, size_t n32) {
    if (CHUNK) {
        const size_t per = (n32 + gridDim.x - 1) / gridDim.x;
        const size_t b = per * blockIdx.x, e = (b + per < n32) ? b + per : n32;
        for (size_t i = b + threadIdx.x; i < e; i += blockDim.x) st32<MODE>(out + 32 * i);
    } else {
        const size_t stride = size_t(gridDim.x) * blockDim.x;
        for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n32; i += stride) st32<MODE>(out + 32 * i);
    }
}
// persistent grid-stride loop whose CTAs start at different phases of the sweep (rot = a per-CTA rotation of the
// iteration order): at any moment the resident CTAs write all over the buffer instead of one compact window
template <int MODE>
__global__ void fill_rot(unsigned char* out, size_t n32) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    const size_t iters = (n32 + stride - 1) / stride;
    const size_t rot = (size_t(blockIdx.x) * 7919u) % iters;
    for (size_t k = 0; k < iters; k++) {
        size_t kk = k + rot;
        if (kk >= iters) kk -= iters;
        const size_t i = kk * stride + size_t(blockIdx.x) * blockDim.x + threadIdx.x;
        if (i < n32) st32<MODE>(out + 32 * i);
    }
}
// the dense kernel's shape (a warp writes 6 KB of zeros and reads 1,536 bytes per tile, warps of a CTA on consecutive
// tiles, tiles advancing by the grid), with and without the rotation
template <bool ROT>
__global__ void mix_tiles(const unsigned char* __restrict__ in, unsigned char* out, size_t ntiles, unsigned* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const size_t stride = size_t(gridDim.x) * wpc;
    const size_t iters = (ntiles + stride - 1) / stride;
    const size_t rot = ROT ? (size_t(blockIdx.x) * 7919u) % iters : 0;
    unsigned acc = 0;
    for (size_t k = 0; k < iters; k++) {
        size_t kk = k + rot;
        if (kk >= iters) kk -= iters;
        const size_t t = kk * stride + size_t(blockIdx.x) * wpc + warp;
        if (t >= ntiles) continue;
        unsigned char* o = out + t * 6144;
#pragma unroll
        for (int q = 0; q < 6; q++) st32<1>(o + (q * 32 + lane) * 32);
        const uint4* i4 = reinterpret_cast<const uint4*>(in + t * 1536);
#pragma unroll
        for (int q = 0; q < 3; q++) { const uint4 v = i4[q * 32 + lane]; acc |= v.x & v.y & v.z & v.w; }
    }
    if (acc == 0x12345678u) *sink = acc;
}
// the same persistent loop with the CTA's warps kept together (SYNC: __syncthreads every SYNC tiles) or with a
// fence every FENCE tiles: is it the lock-step of a young CTA's warps, or the drain at CTA exit, that makes many
// short CTAs faster than a persistent grid?
template <int SYNC, int FENCE>
__global__ void mix_tiles_sync(const unsigned char* __restrict__ in, unsigned char* out, size_t ntiles, unsigned* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const size_t stride = size_t(gridDim.x) * wpc;
    const size_t iters = (ntiles + stride - 1) / stride;
    unsigned acc = 0;
    for (size_t k = 0; k < iters; k++) {
        const size_t t = k * stride + size_t(blockIdx.x) * wpc + warp;
        if (t < ntiles) {
            unsigned char* o = out + t * 6144;
#pragma unroll
            for (int q = 0; q < 6; q++) st32<1>(o + (q * 32 + lane) * 32);
            const uint4* i4 = reinterpret_cast<const uint4*>(in + t * 1536);
#pragma unroll
            for (int q = 0; q < 3; q++) { const uint4 v = i4[q * 32 + lane]; acc |= v.x & v.y & v.z & v.w; }
        }
        if (SYNC && (k % SYNC) == SYNC - 1) __syncthreads();
        if (FENCE && (k % FENCE) == FENCE - 1) __threadfence();
    }
    if (acc == 0x12345678u) *sink = acc;
}
// GROUP warps of the CTA form a group that synchronises (named barrier) and fences every EVERY tiles
template <int GROUP, int EVERY, bool FENCE_FIRST>
__global__ void mix_tiles_group(const unsigned char* __restrict__ in, unsigned char* out, size_t ntiles, unsigned* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const size_t stride = size_t(gridDim.x) * wpc;
    const size_t iters = (ntiles + stride - 1) / stride;
    unsigned acc = 0;
    for (size_t k = 0; k < iters; k++) {
        const size_t t = k * stride + size_t(blockIdx.x) * wpc + warp;
        if (t < ntiles) {
            unsigned char* o = out + t * 6144;
#pragma unroll
            for (int q = 0; q < 6; q++) st32<1>(o + (q * 32 + lane) * 32);
            const uint4* i4 = reinterpret_cast<const uint4*>(in + t * 1536);
#pragma unroll
            for (int q = 0; q < 3; q++) { const uint4 v = i4[q * 32 + lane]; acc |= v.x & v.y & v.z & v.w; }
        }
        if ((k % EVERY) == EVERY - 1) {
            if (FENCE_FIRST) __threadfence();
            asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / GROUP), "r"(GROUP * 32) : "memory");
            if (!FENCE_FIRST) __threadfence();
        }
    }
    if (acc == 0x12345678u) *sink = acc;
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
    const size_t N = size_t(4) << 30;
    unsigned char* out;
    CK(cudaMalloc(&out, N));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("memset %.1f GB/s\n", N / (time_ms([&] { cudaMemsetAsync(out, 0, N); }, 20) * 1e-3) / 1e9);
    const int grids[] = {sms, 2 * sms, 4 * sms, 8 * sms, 32 * sms};
    const int blocks[] = {256, 1024};
    for (int g : grids) for (int blk : blocks) {
        if (g * blk > 32 * sms * 256 * 4) continue;
        float t[8];
        t[0] = time_ms([&] { fill<0, false><<<g, blk>>>(out, N / 32); }, 10);
        t[1] = time_ms([&] { fill<1, false><<<g, blk>>>(out, N / 32); }, 10);
        t[2] = time_ms([&] { fill<2, false><<<g, blk>>>(out, N / 32); }, 10);
        t[3] = time_ms([&] { fill<3, false><<<g, blk>>>(out, N / 32); }, 10);
        t[4] = time_ms([&] { fill<0, true><<<g, blk>>>(out, N / 32); }, 10);
        t[5] = time_ms([&] { fill<1, true><<<g, blk>>>(out, N / 32); }, 10);
        t[6] = time_ms([&] { fill<2, true><<<g, blk>>>(out, N / 32); }, 10);
        t[7] = time_ms([&] { fill<3, true><<<g, blk>>>(out, N / 32); }, 10);
        printf("grid %5d block %4d | stride: v4x2 %.0f v8 %.0f v8.cs %.0f v8.noalloc %.0f | chunk: v4x2 %.0f v8 %.0f v8.cs %.0f v8.noalloc %.0f GB/s\n", g, blk,
               N / t[0] / 1e6, N / t[1] / 1e6, N / t[2] / 1e6, N / t[3] / 1e6, N / t[4] / 1e6, N / t[5] / 1e6, N / t[6] / 1e6, N / t[7] / 1e6);
    }
    for (int g : {sms, 2 * sms, 4 * sms}) for (int blk : {256, 1024})
        printf("rotated persistent fill, grid %4d block %4d: v8 %.0f GB/s\n", g, blk,
               N / time_ms([&] { fill_rot<1><<<g, blk>>>(out, N / 32); }, 10) / 1e6);
    unsigned char* in;
    unsigned* sink;
    CK(cudaMalloc(&in, size_t(1) << 30));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 1, size_t(1) << 30));
    const size_t ntiles = (size_t(1) << 30) / 1536;
    for (int g : {sms, 2 * sms, 16 * sms}) {
        const float a = time_ms([&] { mix_tiles<false><<<g, 1024>>>(in, out, ntiles, sink); }, 10);
        const float b = time_ms([&] { mix_tiles<true><<<g, 1024>>>(in, out, ntiles, sink); }, 10);
        printf("mix 1 GiB read + 4 GiB written, dense-kernel tiles, grid %4d x 1024: in order %.4f ms (%.0f GB/s), rotated %.4f ms (%.0f GB/s)\n",
               g, a, 5.0 * (1 << 30) / a / 1e6, b, 5.0 * (1 << 30) / b / 1e6);
    }
    {
        const int g = sms;
        printf("persistent mix, grid %d x 1024: sync/1 %.4f ms, sync/4 %.4f, sync/16 %.4f, fence/1 %.4f, fence/8 %.4f, sync/1+fence/1 %.4f\n", g,
               time_ms([&] { mix_tiles_sync<1, 0><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_sync<4, 0><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_sync<16, 0><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_sync<0, 1><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_sync<0, 8><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_sync<1, 1><<<g, 1024>>>(in, out, ntiles, sink); }, 10));
        printf("persistent mix, grid %d x 1024, sync + fence: groups of 32 warps every tile %.4f ms, every 2 %.4f, every 4 %.4f; "
               "groups of 16 %.4f, of 8 %.4f, of 4 %.4f, of 1 %.4f; fence before sync (32) %.4f\n", g,
               time_ms([&] { mix_tiles_group<32, 1, false><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_group<32, 2, false><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_group<32, 4, false><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_group<16, 1, false><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_group<8, 1, false><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_group<4, 1, false><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_group<1, 1, false><<<g, 1024>>>(in, out, ntiles, sink); }, 10),
               time_ms([&] { mix_tiles_group<32, 1, true><<<g, 1024>>>(in, out, ntiles, sink); }, 10));
        for (int gg : {4 * sms, 8 * sms, 16 * sms, 64 * sms})
            printf("non-persistent mix, grid %5d x 1024: %.4f ms;  x 256: %.4f ms\n", gg,
                   time_ms([&] { mix_tiles<false><<<gg, 1024>>>(in, out, ntiles, sink); }, 10),
                   time_ms([&] { mix_tiles<false><<<gg * 4, 256>>>(in, out, ntiles, sink); }, 10));
    }
    CK(cudaGetLastError());
    return 0;
}
