// simple_example.cpp -- a caller written against include/PFAC.h only, with the call sequence of
// the reference's test/simple_example.cpp (create -> readPatternFromFile -> dumpTransitionTable
// -> matchFromHost -> print non-zeros -> destroy).  Links against either the reference
// libpfac.so or this repo's pfac_b200/lib/libpfac.so without source changes.
//
//   g++ -O2 -Iinclude examples/simple_example.cpp -Lpfac_b200/lib -lpfac \
//       -Wl,-rpath,$PWD/pfac_b200/lib -o simple_example
//   ./simple_example tests/golden/example_pattern tests/golden/example_input [table.txt]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <PFAC.h>

static void check(PFAC_status_t st, const char* what) {
    if (st != PFAC_STATUS_SUCCESS) {
        std::fprintf(stderr, "Error: %s: %s\n", what, PFAC_getErrorString(st));
        std::exit(1);
    }
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <pattern file> <input file> [table dump]\n", argv[0]);
        return 2;
    }
    PFAC_handle_t handle;
    check(PFAC_create(&handle), "PFAC_create");
    check(PFAC_readPatternFromFile(handle, argv[1]), "PFAC_readPatternFromFile");
    if (argc > 3) {
        FILE* fp = std::fopen(argv[3], "w");
        if (!fp) { std::perror(argv[3]); return 1; }
        check(PFAC_dumpTransitionTable(handle, fp), "PFAC_dumpTransitionTable");
        std::fclose(fp);
    }
    FILE* fin = std::fopen(argv[2], "rb");
    if (!fin) { std::perror(argv[2]); return 1; }
    std::fseek(fin, 0, SEEK_END);
    long n = std::ftell(fin);
    std::rewind(fin);
    std::vector<char> input(static_cast<size_t>(n));
    n = static_cast<long>(std::fread(input.data(), 1, input.size(), fin));
    std::fclose(fin);
    std::vector<int> result(static_cast<size_t>(n), 0);
    check(PFAC_matchFromHost(handle, input.data(), static_cast<size_t>(n), result.data()), "PFAC_matchFromHost");
    for (long i = 0; i < n; i++)
        if (result[static_cast<size_t>(i)] != 0)
            std::printf("At position %4ld, match pattern %d\n", i, result[static_cast<size_t>(i)]);
    check(PFAC_destroy(handle), "PFAC_destroy");
    return 0;
}
