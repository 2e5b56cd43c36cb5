/*
 * PFAC_ext.h -- additive entry points of the B200-native libpfac.so.
 *
 * Nothing in the reference binds these; they exist because (a) the legacy ABI is int-indexed
 * (reference src/PFAC_CPU.cpp:43,60; src/PFAC_kernel.cu:378) and BASELINE configs 3-5 exceed
 * 2^31 bytes, (b) multi-GPU shards need "owned positions + tail halo" (what the reference's
 * test/omp_PFAC.cpp:351-383 does by hand), and (c) the table compiler is host-only code that
 * tools and CPU tests can run without a GPU.
 */
#ifndef PFAC_EXT_H_
#define PFAC_EXT_H_

#include "PFAC.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- north-star aliases (BASELINE.json names; the reference has no such symbols, see
 * SURVEY.md section 0.1).  Same signature and result as PFAC_matchFromDeviceReduce
 * (reference PFAC.h:206); "Inplace" is the reference's PFAC_SPACE_DRIVEN flavour
 * (src/PFAC_reduce_inplace_kernel.cu:155) whose outputs are identical. */
PFAC_status_t PFAC_reduceOnDevice(PFAC_handle_t handle, char *d_inputString, size_t size,
                                  int *d_matched_result, int *d_pos, int *h_num_matched);
PFAC_status_t PFAC_reduceInplaceOnDevice(PFAC_handle_t handle, char *d_inputString, size_t size,
                                         int *d_matched_result, int *d_pos, int *h_num_matched);

/* ---- stream control.  The reference only uses the default stream (PFAC/README:135); this
 * keeps that default and lets a caller move the handle's work to its own cudaStream_t. */
PFAC_status_t PFAC_setStream(PFAC_handle_t handle, void *cuda_stream);

/* ---- patterns from memory: same grammar as the file form (reference
 * src/PFAC_reorder_Table.cpp:121-231), image = the bytes the file would hold. */
PFAC_status_t PFAC_readPatternFromMemory(PFAC_handle_t handle, const char *image, size_t size);

/* ---- patterns from arrays: explicit lengths, so any byte may occur in a pattern (the file
 * grammar cannot express 0x0A, user guide r1.2 p.26).  Pattern i gets ID i+1; order, state
 * numbering and longest-match semantics are those of the file form.  Zero-length patterns are
 * rejected (PFAC_STATUS_INVALID_PARAMETER). */
PFAC_status_t PFAC_readPatternFromArrays(PFAC_handle_t handle, const char *const *patterns,
                                         const size_t *lengths, size_t num_patterns);

/* ---- shard form of PFAC_matchFromDevice: results for positions [0,n_owned); walks may read
 * input[0,n_total), n_total >= n_owned (owned bytes followed by the tail halo, which must be
 * the real following bytes; >= maxPatternLen-1 of them unless the stream ends).  This is the
 * kernel-level contract the reference's omp_PFAC.cpp:351-377 builds by copying. */
PFAC_status_t PFAC_matchShardFromDevice(PFAC_handle_t handle, const char *d_inputString,
                                        size_t n_owned, size_t n_total, int *d_matched_result);

/* ---- 64-bit reduce: positions are long long, count is unsigned long long; sizes >= 2^31 ok */
PFAC_status_t PFAC_matchFromDeviceReduce64(PFAC_handle_t handle, const char *d_inputString,
                                           size_t size, int *d_matched_result, long long *d_pos,
                                           unsigned long long *h_num_matched);

/* shard + 64-bit reduce: emitted position = pos_base + local position (pos_base = global
 * offset of the shard's first byte).  Multi-GPU use: each rank calls this on its shard; an
 * exclusive scan of the ranks' counts gives each rank's offset into the global (ID, position)
 * list -- PFAC_matchShardFromDeviceReduce64Global below does that scan inside the kernel. */
PFAC_status_t PFAC_matchShardFromDeviceReduce64(PFAC_handle_t handle, const char *d_inputString,
                                                size_t n_owned, size_t n_total,
                                                long long pos_base, int *d_matched_result,
                                                long long *d_pos,
                                                unsigned long long *h_num_matched);

/* the same with the capacity (entries) of d_matched_result / d_pos stated: nothing is stored past it;
 * *h_num_matched is always the full count, and PFAC_STATUS_INVALID_PARAMETER says it exceeded the
 * capacity.  (The calls without a capacity follow the reference: buffers of n_owned entries.) */
PFAC_status_t PFAC_matchShardFromDeviceReduce64Cap(PFAC_handle_t handle, const char *d_inputString,
                                                   size_t n_owned, size_t n_total, long long pos_base,
                                                   int *d_matched_result, long long *d_pos, size_t capacity,
                                                   unsigned long long *h_num_matched);

/* ---- cross-GPU count exchange, in the library and on the device.
 * The only inter-GPU step of the path is an exclusive scan of the per-GPU match counts.  A PFAC_comm
 * holds one small device block per rank (mailbox + this rank's region of a global list) that every
 * rank maps: through CUDA IPC between processes (one process per GPU: PFAC_commCreate, exchange the
 * 64-byte handles with whatever the host side has -- MPI, torch.distributed, a file --, then
 * PFAC_commConnect), or through peer access inside one process (PFAC_commCreateLocal).  The reduce
 * kernel then publishes its count into every rank's mailbox and scans the counts itself, over NVLink
 * peer memory: no NCCL call and no host round trip sit between the match and the global offsets.
 * (The reference has no counterpart: test/omp_PFAC.cpp:351-394 stitches on the host; its only
 * inter-GPU mechanism is peer access to caller buffers, test/UVA.cpp:137, which this library also
 * accepts for every d_* argument.)  At most 16 ranks (one NVLink domain). */
typedef struct PFAC_comm *PFAC_comm_t;
#define PFAC_COMM_HANDLE_BYTES 64
PFAC_status_t PFAC_commCreate(PFAC_comm_t *comm, int rank, int world, size_t list_capacity,
                              void *ipc_handle_out /* PFAC_COMM_HANDLE_BYTES, may be NULL when world == 1 */);
PFAC_status_t PFAC_commConnect(PFAC_comm_t comm, const void *all_handles /* world x PFAC_COMM_HANDLE_BYTES, rank order */);
PFAC_status_t PFAC_commCreateLocal(PFAC_comm_t *comms /* [num_devices] */, const int *devices, int num_devices,
                                   size_t list_capacity);
PFAC_status_t PFAC_commDestroy(PFAC_comm_t comm);
/* this rank's region of the global list (device pointers) */
PFAC_status_t PFAC_commGlobalList(PFAC_comm_t comm, int **d_ids, long long **d_pos, size_t *capacity);
/* host copy of entries [first, first + n) of this rank's list region (synchronous cudaMemcpy) */
PFAC_status_t PFAC_commReadGlobalList(PFAC_comm_t comm, size_t first, size_t n, int *h_ids, long long *h_pos);

/* PFAC_matchShardFromDeviceReduce64Cap + the scan, one kernel; collective over the comm.  d_scan (device,
 * 3 words; NULL = kept inside the comm) = {this rank's offset into the global list, total matches, this
 * rank's count}.  h_scan == NULL: asynchronous on the handle's stream, nothing is read back; else the
 * call synchronises once and returns the same three words (PFAC_STATUS_INVALID_PARAMETER when this
 * rank's count exceeded `capacity`, the entries d_matched_result / d_pos hold). */
PFAC_status_t PFAC_matchShardFromDeviceReduce64Global(PFAC_handle_t handle, PFAC_comm_t comm,
                                                      const char *d_inputString, size_t n_owned, size_t n_total,
                                                      long long pos_base, int *d_matched_result, long long *d_pos,
                                                      size_t capacity, unsigned long long *d_scan,
                                                      unsigned long long *h_scan);
/* optional second step, collective: every rank stores its run into rank dst_rank's list region at its
 * scanned offset (P2P stores over NVLink; count and offset are read from d_scan on the device), dst_rank
 * waits on the device until all runs have landed.  The list is then PFAC_commGlobalList(dst)[0, total). */
PFAC_status_t PFAC_commGatherRuns(PFAC_handle_t handle, PFAC_comm_t comm, int dst_rank, const int *d_matched_result,
                                  const long long *d_pos, const unsigned long long *d_scan, int synchronize);

/* ---- table compiler, host only (no CUDA calls): pattern image -> reference-numbered trie
 * -> B200 device layout (dense root row, 2-byte prefilter bitmap, hot/cold bucketed hash
 * rows, path-compressed chains). */
typedef struct PFAC_table *PFAC_table_t;

typedef struct {
    int num_patterns;      /* k */
    int num_states;        /* reference numOfStates (counts unused state 0) */
    int initial_state;     /* k+1 */
    int max_pattern_len;
    int num_leaves;        /* final states without out-edges */
    int num_edges;         /* distinct (state,ch) transitions of the matching automaton */
    int hash_edges;        /* hash-row entries left after chain compression */
    int num_chains;        /* compressed single-child runs (tail compared byte-wise) */
    int tail_bytes;        /* bytes of chain tails (padded) */
    int chains_hot;        /* 1: chain records and tails also live in shared memory */
    int next2_hot;         /* 1: the direct K-gram table lives in shared memory */
    int code_bits;         /* b: bits per symbol in the prefilter index (8, 4 or 2) */
    int gram_len;          /* K = 16 / b symbols covered by the prefilter + direct table */
    int has_best2;         /* 1: patterns shorter than K exist (best2 array present) */
    int has_chk2;          /* 1: second prefilter stage (next-symbol sets per K-gram) is on */
    int max_depth;
    int hot_depth;         /* edges whose source state has depth in [1,hot_depth) are "hot" */
    unsigned hot_buckets;  /* 16-byte buckets (2 slots) of the shared-memory hash rows */
    unsigned cold_buckets; /* 16-byte buckets of the global (L2-resident) hash rows */
    unsigned hash_mul;
    int hot_max_probe, cold_max_probe;
    int pre2_bits_set;     /* of 65536: (c0,c1) pairs that survive the prefilter */
    int root_fanout;       /* valid first bytes */
    int hashed_filter;     /* 1 / 2: the per-position test is the hashed 4-gram filter (hfilt_words words), 1 or 2 bits per lookup;
                              3: pair filter, one lookup per two start positions (PFAC_tableGetFilter) */
    int hfilt_bits_set;    /* of 32 * hfilt_words */
    int code_shift;        /* b = 2 with an arithmetic symbol code: code = (byte >> code_shift) & 3; else -1.  With
                              hashed_filter == 2 and code_bits == 2 the first stage hashes ten symbols (20 bits):
                              word ((x * 0x9E3779B1) >> 19) & 8191, bits 31 - ((x * 0x85EBCA6B) >> 27) and
                              31 - ((x * 0xC2B2AE35) >> 27), products mod 2^32 */
    size_t device_bytes;   /* total bytes uploaded */
    unsigned hfilt_words;  /* words of the hashed filter: 0, 8192, or 16384 (hashed_filter == 2, byte alphabet, budget >= 64 KB) */
} PFAC_tableInfo_t;

PFAC_status_t PFAC_tableCompile(const char *image, size_t size, size_t hot_budget_bytes,
                                PFAC_table_t *table);
PFAC_status_t PFAC_tableCompileFile(const char *filename, size_t hot_budget_bytes,
                                    PFAC_table_t *table);
PFAC_status_t PFAC_tableCompileArrays(const char *const *patterns, const size_t *lengths,
                                      size_t num_patterns, size_t hot_budget_bytes,
                                      PFAC_table_t *table);
PFAC_status_t PFAC_tableDestroy(PFAC_table_t table);
PFAC_status_t PFAC_tableDump(PFAC_table_t table, FILE *fp);
PFAC_status_t PFAC_tableDumpToFile(PFAC_table_t table, const char *filename);
PFAC_status_t PFAC_tableGetInfo(PFAC_table_t table, PFAC_tableInfo_t *info);
/* read-only views of the layout arrays (valid until PFAC_tableDestroy):
 * root: 256 int; pre2: 2048 unsigned (bit idx = sum code(c_i)<<(b*i) at word idx>>5, bit 31-(idx&31));
 * rank2: 2048 unsigned short prefix popcounts; next2: pre2_bits_set unsigned (state, chain
 * reference or 0xFFFFFFFF); hot/cold: 4 unsigned per bucket {key0,val0,key1,val1} (val bit 31
 * set = chain index); chains: 4 unsigned per record {tail offset, len, end state | leaf bit
 * 31, first 4 tail bytes}; tails: tail_bytes bytes */
PFAC_status_t PFAC_tableGetLayout(PFAC_table_t table, const int **root, const unsigned **pre2,
                                  const unsigned short **rank2, const unsigned **next2,
                                  const unsigned **hot, const unsigned **cold,
                                  const unsigned **chains, const unsigned char **tails);
/* lut: 256 bytes (symbol code | 0x80 = byte in no pattern); best2: parallel to next2, NULL when
 * has_best2 == 0; chk2: parallel to next2 (16-bit set of next byte & 15), NULL when has_chk2 == 0 */
PFAC_status_t PFAC_tableGetLayout2(PFAC_table_t table, const unsigned char **lut,
                                   const unsigned **best2, const unsigned short **chk2);

/* hashed 4-gram first stage, used instead of pre2 as the per-position test for byte alphabets
 * (b = 8) whenever the shared-memory budget holds its 32 KB: hfilt_words unsigned, NULL when
 * hashed_filter == 0.  x = c0|c1<<8|c2<<16|c3<<24.  hashed_filter == 1 (sparse tables): word
 * ((x * 0x9E3779B1) >> 2) & 8191, bit 31 - (umulhi(x, 0x85EBCA6B) & 31).  hashed_filter == 2 (dense
 * tables): word x & (hfilt_words - 1), i.e. picked by c0 and the low bits of c1 themselves (a 1-byte
 * pattern then fills the words of its own first byte only), that bit and bit
 * 31 - (umulhi(x, 0xC2B2AE35) & 31).  hashed_filter == 3 (sparse tables, no pattern shorter than three
 * bytes): one lookup for the start positions q and q+1, keyed by the bytes they share,
 * y = text[q+1] | text[q+2]<<8 | text[q+3]<<16, h = (y * 0x9E3779B1) mod 2^24, word (h >> 2) & 8191, bit
 * 31 - (h >> 19); the table holds bytes 0..2 and bytes 1..3 of every pattern.  Survivors are re-checked
 * exactly against pre2 / chk2 by the walker.
 * PFAC_B200_FILTER=exact keeps the exact 2-gram stage, =nopair the per-position hashed filter (read at
 * table compile time). */
PFAC_status_t PFAC_tableGetFilter(PFAC_table_t table, const unsigned **hfilt);

/* ---- compiled-table files: parse, sort, number and lay out a large dictionary once, later
 * processes read the result back (20,000 patterns: 0.13 s to load, 0.15 s to compile).  Versioned and checksummed; a file that
 * is not recognised gives PFAC_STATUS_INVALID_PARAMETER (compile from the pattern file instead), a
 * missing one PFAC_STATUS_FILE_OPEN_ERROR.  PFAC_loadCompiledPatterns leaves the handle exactly as
 * PFAC_readPatternFromFile on the original pattern file would (same IDs, same dump, same results);
 * layouts stored for another shared-memory budget / filter policy are recompiled from the stored
 * automaton. */
PFAC_status_t PFAC_saveCompiledPatterns(PFAC_handle_t handle, const char *filename);
PFAC_status_t PFAC_loadCompiledPatterns(PFAC_handle_t handle, const char *filename);
PFAC_status_t PFAC_tableSave(PFAC_table_t table, const char *filename);      /* host only */
PFAC_status_t PFAC_tableLoad(const char *filename, PFAC_table_t *table);     /* host only */

/* info / dump-to-path for a live handle */
PFAC_status_t PFAC_getTableInfo(PFAC_handle_t handle, PFAC_tableInfo_t *info);        /* layout of the dense kernel */
PFAC_status_t PFAC_getTableInfoReduce(PFAC_handle_t handle, PFAC_tableInfo_t *info);  /* layout of the reduce kernel: its
                                                                                         own shared-memory budget, and the pair
                                                                                         filter where the dictionary allows it */
PFAC_status_t PFAC_dumpTransitionTableToFile(PFAC_handle_t handle, const char *filename);

/* ---- table-size report to stdout, counterpart of the reference's PFAC_memoryUsage
 * (src/PFAC.cpp:1250-1306; not in the reference's public header either) */
PFAC_status_t PFAC_memoryUsage(PFAC_handle_t handle);

/* ---- multi-GPU driver, one process: what reference test/omp_PFAC.cpp:257-394 builds by hand.
 * One handle per listed device, contiguous shards with a (maxPatternLen-1)-byte tail halo, one host
 * thread per GPU running the chunked host pipeline; the reduced form places each GPU's run at the
 * exclusive scan of the per-GPU counts.  h_matched_result / h_pos of the reduced call must hold `size`
 * entries (the reference asks the same of its reduce buffers).  For one process per GPU use the shard
 * entry points above with the in-kernel count scan (PFAC_comm, above). */
typedef struct PFAC_mgpu *PFAC_mgpu_t;
PFAC_status_t PFAC_mgpuCreate(PFAC_mgpu_t *mg, const int *devices, int num_devices);
PFAC_status_t PFAC_mgpuDestroy(PFAC_mgpu_t mg);
PFAC_status_t PFAC_mgpuReadPatternFromFile(PFAC_mgpu_t mg, char *filename);
PFAC_status_t PFAC_mgpuMatchFromHost(PFAC_mgpu_t mg, char *h_inputString, size_t size, int *h_matched_result);
PFAC_status_t PFAC_mgpuMatchFromHostReduce64(PFAC_mgpu_t mg, char *h_inputString, size_t size,
                                             int *h_matched_result, long long *h_pos,
                                             unsigned long long *h_num_matched);

/*
 * Pageable (malloc'ed) host buffers -- what the reference's callers pass (reference
 * test/simple_example.cpp) -- are moved through library-owned pinned staging buffers by a small
 * pool of host threads, overlapped with the DMA and the kernel of the neighbouring chunks.
 * Pinned (cudaHostAlloc / cudaHostRegister) buffers are DMA'd directly.  Environment:
 *   PFAC_B200_COPY_THREADS   host threads per process that copy, caller included (default min(8, cores/2))
 *   PFAC_B200_STAGE_CHUNK_MB bytes of input per staged chunk (default 8)
 *   PFAC_B200_STAGE=0        no staging: pageable pointers go straight to cudaMemcpyAsync
 * PFAC_hostCopy is that pool's memcpy (host only, no GPU needed; exported for tests and for callers
 * that fill their own pinned buffers).
 */
/*
 * PFAC_matchFromHost returns 4 bytes per input byte, almost all of them zero.  By default only the
 * (id, position) pairs of each chunk cross PCIe (the fused match + compaction kernel produces
 * them) and the host writes the dense array: zero fill by the copy pool while the chunk is on the
 * GPU, then a scatter of the pairs.  Chunks with more than one match per 16 positions use the dense
 * kernel and a plain D2H.  The result is the same array either way.
 *   PFAC_B200_HOST_RESULT=dense   always return the dense array over PCIe
 * PFAC_lastHostTransfer: bytes the handle's last PFAC_matchFromHost* call moved over PCIe.
 */
PFAC_status_t PFAC_lastHostTransfer(PFAC_handle_t handle, size_t *h2d_bytes, size_t *d2h_bytes);
/* The host pipelines keep their buffers in the handle between calls: per handle 2 x (chunk + halo) of
 * input, 2 x 4 x chunk of ids / dense results, 2 x 4 x chunk of positions (8 x chunk once a 64-bit
 * host call was made) on the device -- about 770 MiB at the default 32 MiB chunk
 * (PFAC_B200_HOST_CHUNK_MB) -- plus 2 x chunk/2 of pinned (id, position) lists and, for pageable
 * buffers, the pinned staging.  PFAC_destroy frees them; this call frees them earlier. */
PFAC_status_t PFAC_releaseHostBuffers(PFAC_handle_t handle);

PFAC_status_t PFAC_hostCopy(void *dst, const void *src, size_t bytes);
/* the same pool's zero fill (streaming stores); PFAC_matchFromHost uses it for the sparse result path */
PFAC_status_t PFAC_hostZero(void *dst, size_t bytes);

/* kernels launched by this library since it was loaded (bench.py's gpu_launches) */
unsigned long long PFAC_kernelLaunchCount(void);

/* library build tag, e.g. "pfac-b200 sm_100a" */
const char *PFAC_versionString(void);

#ifdef __cplusplus
}
#endif

#endif /* PFAC_EXT_H_ */
