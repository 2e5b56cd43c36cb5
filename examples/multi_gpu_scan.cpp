// multi_gpu_scan.cpp -- the multi-GPU shape of reference test/omp_PFAC.cpp (one process, the input cut into
// contiguous segments with an overlap, one GPU per segment, results stitched and compared element by element
// with a single-GPU run, omp_PFAC.cpp:351-439) written against include/PFAC.h + PFAC_ext.h: the segments stay
// on the devices, every GPU runs the fused match + compaction kernel, the exclusive scan of the per-GPU match
// counts happens inside that kernel over peer memory (PFAC_comm), and the runs are stored into ONE global
// (ID, position) list on GPU 0 by P2P stores.  No host-side stitching.
//
//   g++ -O2 -Iinclude -I/usr/local/cuda/include examples/multi_gpu_scan.cpp -Lpfac_b200/lib -lpfac
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/pfac_b200/lib -o multi_gpu_scan
//   ./multi_gpu_scan <pattern file> <input file> [number of GPUs]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include <PFAC.h>
#include <PFAC_ext.h>

static void check(PFAC_status_t st, const char* what) {
    if (st != PFAC_STATUS_SUCCESS) {
        std::fprintf(stderr, "Error: %s: %s\n", what, PFAC_getErrorString(st));
        std::exit(1);
    }
}
static void cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        std::fprintf(stderr, "Error: %s: %s\n", what, cudaGetErrorString(e));
        std::exit(1);
    }
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <pattern file> <input file> [number of GPUs]\n", argv[0]);
        return 2;
    }
    int G = 0;
    cuda(cudaGetDeviceCount(&G), "cudaGetDeviceCount");
    if (argc > 3) G = std::min(G, std::atoi(argv[3]));
    if (G < 1) { std::fprintf(stderr, "no CUDA device\n"); return 1; }

    FILE* fin = std::fopen(argv[2], "rb");
    if (!fin) { std::perror(argv[2]); return 1; }
    std::fseek(fin, 0, SEEK_END);
    const size_t n = size_t(std::ftell(fin));
    std::rewind(fin);
    std::vector<char> text(n);
    if (std::fread(text.data(), 1, n, fin) != n) { std::fprintf(stderr, "short read\n"); return 1; }
    std::fclose(fin);

    // one handle per GPU, the same dictionary on each
    const size_t ng = static_cast<size_t>(G);
    std::vector<int> devices(ng);
    std::vector<PFAC_handle_t> handle(ng);
    PFAC_tableInfo_t info;
    for (int g = 0; g < G; g++) {
        devices[size_t(g)] = g;
        cuda(cudaSetDevice(g), "cudaSetDevice");
        check(PFAC_create(&handle[size_t(g)]), "PFAC_create");
        check(PFAC_readPatternFromFile(handle[size_t(g)], argv[1]), "PFAC_readPatternFromFile");
    }
    check(PFAC_getTableInfo(handle[0], &info), "PFAC_getTableInfo");
    const size_t halo = size_t(std::max(info.max_pattern_len - 1, 0));

    // every rank's block (mailbox + list region) mapped into every GPU; the global list lives on GPU 0
    std::vector<PFAC_comm_t> comm(ng);
    check(PFAC_commCreateLocal(comm.data(), devices.data(), G, n), "PFAC_commCreateLocal");

    // contiguous shards + tail halo, resident on their GPU
    const size_t per = ((n + ng - 1) / ng + 4095) / 4096 * 4096;
    std::vector<char*> d_in(ng, nullptr);
    std::vector<int*> d_id(ng, nullptr);
    std::vector<long long*> d_pos(ng, nullptr);
    std::vector<unsigned long long*> d_scan(ng, nullptr);
    std::vector<size_t> start(ng), owned(ng), total(ng);
    for (int g = 0; g < G; g++) {
        const size_t s = std::min(size_t(g) * per, n), e = std::min(s + per, n);
        start[size_t(g)] = s;
        owned[size_t(g)] = e - s;
        total[size_t(g)] = std::min(e + halo, n) - s;
        cuda(cudaSetDevice(g), "cudaSetDevice");
        cuda(cudaMalloc(reinterpret_cast<void**>(&d_in[size_t(g)]), total[size_t(g)] + 16), "cudaMalloc");
        cuda(cudaMalloc(reinterpret_cast<void**>(&d_id[size_t(g)]), (owned[size_t(g)] + 1) * sizeof(int)), "cudaMalloc");
        cuda(cudaMalloc(reinterpret_cast<void**>(&d_pos[size_t(g)]), (owned[size_t(g)] + 1) * sizeof(long long)), "cudaMalloc");
        cuda(cudaMalloc(reinterpret_cast<void**>(&d_scan[size_t(g)]), 3 * sizeof(unsigned long long)), "cudaMalloc");
        cuda(cudaMemcpy(d_in[size_t(g)], text.data() + s, total[size_t(g)], cudaMemcpyHostToDevice), "cudaMemcpy");
    }

    // all kernels are enqueued without a host synchronisation: they wait for each other's count on the device
    for (int g = G - 1; g >= 0; g--) {
        cuda(cudaSetDevice(g), "cudaSetDevice");
        check(PFAC_matchShardFromDeviceReduce64Global(handle[size_t(g)], comm[size_t(g)], d_in[size_t(g)], owned[size_t(g)],
                                                      total[size_t(g)], (long long)start[size_t(g)], d_id[size_t(g)],
                                                      d_pos[size_t(g)], owned[size_t(g)], d_scan[size_t(g)], nullptr),
              "PFAC_matchShardFromDeviceReduce64Global");
    }
    for (int g = 0; g < G; g++) {
        cuda(cudaSetDevice(g), "cudaSetDevice");
        check(PFAC_commGatherRuns(handle[size_t(g)], comm[size_t(g)], 0, d_id[size_t(g)], d_pos[size_t(g)], d_scan[size_t(g)], 0),
              "PFAC_commGatherRuns");
    }
    unsigned long long total_matches = 0;
    for (int g = 0; g < G; g++) {
        cuda(cudaSetDevice(g), "cudaSetDevice");
        cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
        unsigned long long scan[3];
        cuda(cudaMemcpy(scan, d_scan[size_t(g)], sizeof(scan), cudaMemcpyDeviceToHost), "cudaMemcpy");
        std::printf("GPU %d: positions [%zu, %zu), %llu matches, offset %llu in the global list\n", g, start[size_t(g)],
                    start[size_t(g)] + owned[size_t(g)], scan[2], scan[0]);
        total_matches = scan[1];
    }
    std::vector<int> ids(total_matches);
    std::vector<long long> pos(total_matches);
    check(PFAC_commReadGlobalList(comm[0], 0, total_matches, ids.data(), pos.data()), "PFAC_commReadGlobalList");

    // self-check against one GPU matching the whole input through the host API (as omp_PFAC.cpp:397-439)
    std::vector<int> dense(n);
    cuda(cudaSetDevice(0), "cudaSetDevice");
    check(PFAC_matchFromHost(handle[0], text.data(), n, dense.data()), "PFAC_matchFromHost");
    size_t k = 0, bad = 0;
    for (size_t i = 0; i < n; i++) {
        if (dense[i] == 0) continue;
        if (k >= total_matches || pos[k] != (long long)i || ids[k] != dense[i]) bad++;
        k++;
    }
    if (k != total_matches) bad++;
    std::printf("number of matched = %llu on %d GPU(s); single-GPU check: %s\n", total_matches, G, bad ? "MISMATCH" : "identical");
    for (size_t i = 0; i < std::min<size_t>(total_matches, 10); i++)
        std::printf("At position %4lld, match pattern %d\n", pos[i], ids[i]);

    for (int g = 0; g < G; g++) {
        cuda(cudaSetDevice(g), "cudaSetDevice");
        cudaFree(d_in[size_t(g)]); cudaFree(d_id[size_t(g)]); cudaFree(d_pos[size_t(g)]); cudaFree(d_scan[size_t(g)]);
        check(PFAC_commDestroy(comm[size_t(g)]), "PFAC_commDestroy");
        check(PFAC_destroy(handle[size_t(g)]), "PFAC_destroy");
    }
    return bad ? 1 : 0;
}
