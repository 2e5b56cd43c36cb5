/*
 * ref_driver.cpp -- thin driver over the UNMODIFIED reference CPU path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/pfac_oracle.c header for who may load it).
 *
 * oracle/Makefile compiles this file together with the reference's own sources, read in
 * place from $(REF)/PFAC/src (PFAC_reorder_Table.cpp, PFAC_CPU.cpp, PFAC_CPU_OMP.cpp,
 * PFAC.cpp) into oracle/_ref/libpfac_ref.so.  No reference source is copied into this repo.
 *
 * The reference's public PFAC_create() needs a GPU plus a per-arch kernel module
 * (src/PFAC.cpp:133-204), so the driver assembles a PFAC_context by hand from the
 * reference's internal functions, in the order PFAC_readPatternFromFile does
 * (src/PFAC.cpp:653-735), and fills the host 2-D table the way PFAC_create2DTable does
 * (src/PFAC.cpp:364-382) minus the cudaMalloc/cudaMemcpy.  Matching then goes through the
 * reference's PFAC_CPU / PFAC_CPU_OMP (src/PFAC_CPU.cpp:43, src/PFAC_CPU_OMP.cpp) and the
 * dump through the reference's PFAC_dumpTransitionTable (src/PFAC.cpp:1188).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include <omp.h>

#include <PFAC.h>
#include <PFAC_P.h>

using namespace std;

/* internal reference entry points (external C++ linkage in the reference sources) */
PFAC_status_t PFAC_CPU(PFAC_handle_t handle, char *h_input_string, const int input_size,
                       int *h_matched_result);
PFAC_status_t PFAC_CPU_OMP(PFAC_handle_t handle, char *input_string, const int input_size,
                           int *h_matched_result);
PFAC_status_t create_PFACTable_spaceDriven(const char **rowPtr, const int *patternLen_table,
                                           const int *patternID_table, const int max_state_num,
                                           const int pattern_num, const int initial_state,
                                           const int baseOfUsableStateID, int *state_num_ptr,
                                           vector<vector<TableEle> > &PFAC_table);
void PFAC_freeResource(PFAC_handle_t handle);

extern "C" {

int ref_create(const char *pattern_file, void **out)
{
    PFAC_handle_t h = (PFAC_handle_t)calloc(1, sizeof(PFAC_context));
    if (!h) return PFAC_STATUS_ALLOC_FAILED;
    PFAC_status_t st = parsePatternFile((char *)pattern_file, &h->rowPtr, &h->valPtr,
                                        &h->patternID_table, &h->patternLen_table,
                                        &h->max_numOfStates, &h->numOfPatterns);
    if (st != PFAC_STATUS_SUCCESS) { free(h); return st; }
    h->maxPatternLen = 0;
    for (int i = 1; i <= h->numOfPatterns; i++)
        if (h->maxPatternLen < h->patternLen_table[i]) h->maxPatternLen = h->patternLen_table[i];
    h->initial_state = h->numOfPatterns + 1;
    h->numOfFinalStates = h->numOfPatterns;
    h->table_compact = new vector<vector<TableEle> >;
    st = create_PFACTable_spaceDriven((const char **)h->rowPtr, h->patternLen_table,
                                      h->patternID_table, h->max_numOfStates, h->numOfPatterns,
                                      h->initial_state, h->initial_state + 1, &h->numOfStates,
                                      *h->table_compact);
    if (st != PFAC_STATUS_SUCCESS) { PFAC_freeResource(h); free(h); return st; }
    size_t cells = (size_t)h->numOfStates * CHAR_SET;
    h->numOfTableEntry = cells;
    h->sizeOfTableEntry = sizeof(int);
    h->sizeOfTableInBytes = cells * sizeof(int);
    h->h_PFAC_table = (int *)malloc(cells * sizeof(int));
    if (!h->h_PFAC_table) { PFAC_freeResource(h); free(h); return PFAC_STATUS_ALLOC_FAILED; }
    for (size_t i = 0; i < cells; i++) h->h_PFAC_table[i] = TRAP_STATE;
    for (int s = 0; s < h->numOfStates; s++) {
        const vector<TableEle> &row = (*h->table_compact)[s];
        for (size_t j = 0; j < row.size(); j++)
            h->h_PFAC_table[(size_t)s * CHAR_SET + row[j].ch] = row[j].nextState;
    }
    h->perfMode = PFAC_TIME_DRIVEN;
    h->platform = PFAC_PLATFORM_CPU;
    h->isPatternsReady = true;
    *out = h;
    return 0;
}

void ref_destroy(void *p)
{
    PFAC_handle_t h = (PFAC_handle_t)p;
    if (!h) return;
    PFAC_freeResource(h); /* frees host tables; device pointers are NULL */
    free(h);
}

/* use_omp: 0 -> PFAC_CPU, 1 -> PFAC_CPU_OMP (thread count from OMP_NUM_THREADS / runtime) */
int ref_match(void *p, const char *in, int n, int *out, int use_omp)
{
    PFAC_handle_t h = (PFAC_handle_t)p;
    return use_omp ? PFAC_CPU_OMP(h, (char *)in, n, out) : PFAC_CPU(h, (char *)in, n, out);
}

int ref_dump(void *p, const char *path)
{
    FILE *fp = fopen(path, "w");
    if (!fp) return PFAC_STATUS_FILE_OPEN_ERROR;
    int st = PFAC_dumpTransitionTable((PFAC_handle_t)p, fp);
    fclose(fp);
    return st;
}

/* torchrun exports OMP_NUM_THREADS=1; the bench sets the count it reports explicitly */
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ref_get_threads(void) { return omp_get_max_threads(); }

int ref_num_patterns(void *p) { return ((PFAC_handle_t)p)->numOfPatterns; }
int ref_num_states(void *p) { return ((PFAC_handle_t)p)->numOfStates; }
int ref_initial_state(void *p) { return ((PFAC_handle_t)p)->initial_state; }
int ref_max_pattern_len(void *p) { return ((PFAC_handle_t)p)->maxPatternLen; }
const int *ref_dense_table(void *p) { return ((PFAC_handle_t)p)->h_PFAC_table; }

} /* extern "C" */
