/*
 * PFAC.h -- C ABI of the B200-native PFAC matching library (libpfac.so).
 *
 * Drop-in boundary: every enum value, type name and prototype below is ABI-identical to the
 * reference's public header (reference PFAC/include/PFAC.h), so a program compiled against
 * the reference header relinks against this library unchanged.  Each declaration cites the
 * reference line it replaces.  Implementation: pfac_b200/csrc/ (sm_100a only; there is no CPU
 * matching path in this library -- PFAC_create fails without a CUDA device).
 *
 * Additive entry points (64-bit sizes/positions, shards, streams, table compiler) are in
 * PFAC_ext.h; nothing here depends on them.
 */
#ifndef PFAC_H_
#define PFAC_H_

#include <stdio.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference PFAC.h:27-31.  B200 build: matchFromDevice* always run on the GPU; matchFromHost*
 * run on the GPU for every platform value (CPU / CPU_OMP are accepted and stored, but this
 * library has no CPU matcher -- the reference's CPU path lives on only as the test oracle). */
typedef enum {
    PFAC_PLATFORM_GPU = 0,
    PFAC_PLATFORM_CPU = 1,
    PFAC_PLATFORM_CPU_OMP = 2
} PFAC_platform_t;

/* reference PFAC.h:33-37.  Accepted and stored; texture references do not exist on sm_100. */
typedef enum {
    PFAC_AUTOMATIC = 0,
    PFAC_TEXTURE_ON = 1,
    PFAC_TEXTURE_OFF = 2
} PFAC_textureMode_t;

/* reference PFAC.h:39-42.  Both modes give identical results.  SPACE_DRIVEN asks the table
 * compiler for the most compact shared-memory footprint (no hot hash rows in smem). */
typedef enum {
    PFAC_TIME_DRIVEN = 0,
    PFAC_SPACE_DRIVEN = 1
} PFAC_perfMode_t;

/* reference PFAC.h:57-70.  Values below PFAC_STATUS_BASE are raw cudaError_t codes. */
typedef enum {
    PFAC_STATUS_SUCCESS = 0,
    PFAC_STATUS_BASE = 10000,
    PFAC_STATUS_ALLOC_FAILED,
    PFAC_STATUS_CUDA_ALLOC_FAILED,
    PFAC_STATUS_INVALID_HANDLE,
    PFAC_STATUS_INVALID_PARAMETER,
    PFAC_STATUS_PATTERNS_NOT_READY,
    PFAC_STATUS_FILE_OPEN_ERROR,
    PFAC_STATUS_LIB_NOT_EXIST,
    PFAC_STATUS_ARCH_MISMATCH,
    PFAC_STATUS_MUTEX_ERROR,
    PFAC_STATUS_INTERNAL_ERROR
} PFAC_status_t;

/* reference PFAC.h:72-74: opaque handle */
struct PFAC_context;
typedef struct PFAC_context *PFAC_handle_t;

/* reference PFAC.h:87 (src/PFAC.cpp:133-204).  Binds to the current CUDA device.
 * Returns the raw CUDA error when no device is usable (as the reference does, :148-151),
 * PFAC_STATUS_ARCH_MISMATCH when the device is not compute capability 10.x. */
PFAC_status_t PFAC_create(PFAC_handle_t *handle);

/* reference PFAC.h:96 (src/PFAC.cpp:207-218) */
PFAC_status_t PFAC_destroy(PFAC_handle_t handle);

/* reference PFAC.h:106 (src/PFAC.cpp:741-757) */
PFAC_status_t PFAC_setPlatform(PFAC_handle_t handle, PFAC_platform_t platform);

/* reference PFAC.h:119 (src/PFAC.cpp:764-780) */
PFAC_status_t PFAC_setTextureMode(PFAC_handle_t handle, PFAC_textureMode_t textureModeSel);

/* reference PFAC.h:130 (src/PFAC.cpp:782-817).  Changing the mode after patterns are loaded
 * re-emits the device layout, as the reference rebuilds its table. */
PFAC_status_t PFAC_setPerfMode(PFAC_handle_t handle, PFAC_perfMode_t perfModeSel);

/* reference PFAC.h:139 (src/PFAC.cpp:1131-1185): static strings, same text */
const char *PFAC_getErrorString(PFAC_status_t status);

/* reference PFAC.h:148 (src/PFAC.cpp:1188-1246): byte-identical text dump; fp==NULL -> stdout */
PFAC_status_t PFAC_dumpTransitionTable(PFAC_handle_t handle, FILE *fp);

/* reference PFAC.h:166 (src/PFAC.cpp:653-735).  Same file grammar: patterns split on 0x0A
 * only, a final line without newline is dropped, IDs are 1-based file order.  A blank line
 * followed by a pattern makes the reference abort on an assert; here it returns
 * PFAC_STATUS_INVALID_PARAMETER and leaves the handle without patterns. */
PFAC_status_t PFAC_readPatternFromFile(PFAC_handle_t handle, char *filename);

/* reference PFAC.h:179 (src/PFAC.cpp:843-876).  d_matched_result[i] = ID of the longest
 * pattern that is a prefix of input[i..size), else 0; every element i<size is written.
 * Asynchronous on the handle's stream (legacy default stream unless PFAC_setStream).
 * Fast path needs both pointers 16-byte aligned (cudaMalloc gives 256); any alignment works.
 * Unlike the reference, nothing beyond input[size) is read. */
PFAC_status_t PFAC_matchFromDevice(PFAC_handle_t handle, char *d_inputString, size_t size,
                                   int *d_matched_result);

/* reference PFAC.h:198 (src/PFAC.cpp:879-961).  Host buffers; chunked, double-buffered
 * H2D / kernel / D2H pipeline instead of the reference's malloc-copy-run-copy-free. */
PFAC_status_t PFAC_matchFromHost(PFAC_handle_t handle, char *h_inputString, size_t size,
                                 int *h_matched_result);

/* reference PFAC.h:206 (src/PFAC.cpp:964-1008).  *h_num_matched = M, d_matched_result[0..M)
 * and d_pos[0..M) = the non-zero (ID, position) pairs in ascending position.  M==0 leaves
 * the buffers untouched.  Synchronous.  size must be < 2^31 (int positions); use
 * PFAC_matchFromDeviceReduce64 (PFAC_ext.h) beyond that.  Same NULL checks, in the same
 * order, as the reference (:967-981). */
PFAC_status_t PFAC_matchFromDeviceReduce(PFAC_handle_t handle, char *d_inputString, size_t size,
                                         int *d_matched_result, int *d_pos, int *h_num_matched);

/* reference PFAC.h:214 (src/PFAC.cpp:1010-1128) */
PFAC_status_t PFAC_matchFromHostReduce(PFAC_handle_t handle, char *h_inputString, size_t size,
                                       int *h_matched_result, int *h_pos, int *h_num_matched);

#ifdef __cplusplus
}
#endif

#endif /* PFAC_H_ */
