/*
 * pfac_oracle.c -- CPU restatement of the reference PFAC matching path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (pfac_b200/lib/libpfac.so) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   - the doc goldens G1..G4 (README.md:114-120 of the reference,
 *     user guide r1.2 p.21/27/29, PFAC_hash_draft.pdf p.1), and
 *   - fixtures produced by the reference's own CPU path compiled in oracle/_ref
 *     (tests/golden/make_golden.py, committed outputs under tests/golden/), and
 *   - oracle/_ref itself on randomized planted cases whenever oracle/_ref is built.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference/PFAC).  Plain C99, no dependencies; OpenMP optional (-fopenmp).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define ORC_CHAR_SET 256
#define ORC_TRAP (-1) /* TRAP_STATE 0xFFFFFFFF compares equal to int -1, include/PFAC_P.h:181-182 */

typedef struct {
    int ch;
    int next;
} orc_edge_t;

typedef struct {
    orc_edge_t *e;
    int n, cap;
} orc_row_t;

typedef struct orc_machine {
    /* pattern store, restating parsePatternFile (src/PFAC_reorder_Table.cpp:121-231) */
    char *buf;          /* file image */
    long file_size;
    int num_patterns;   /* k */
    char **sorted_ptr;  /* rowPtr[0..k) : sorted pattern pointers */
    int *sorted_id;     /* patternID_table[0..k) : original 1-based IDs in sorted order */
    int *len_by_id;     /* patternLen_table[0..k] : [0] = 0 */
    int max_pattern_len;
    /* automaton, restating create_PFACTable_spaceDriven (src/PFAC_reorder_Table.cpp:256-329) */
    int initial_state;  /* k+1, src/PFAC.cpp:693 */
    int num_final;      /* k */
    int num_states;     /* next unused id; counts the unused state 0 */
    orc_row_t *rows;    /* per-state edges in insertion order */
    int *dense;         /* num_states x 256, src/PFAC.cpp:364-382 */
} orc_machine_t;

/* ---- pattern order: pattern_cmp_functor, src/PFAC_reorder_Table.cpp:37-72 ----------------
 * Plain (signed on x86-64) char comparison, '\n'-terminated, proper prefix first.
 * The reference returns true for equal strings (not a strict weak order, duplicate order
 * unspecified); here equal strings tie-break on file order (ascending ID) -- same choice as
 * the product's table compiler; parity for duplicate patterns is otherwise unpinned
 * (SURVEY.md section 8(c)). */
typedef struct {
    char *s;
    int id;
} orc_pat_t;

static int orc_pat_cmp(const void *a, const void *b)
{
    const orc_pat_t *pa = (const orc_pat_t *)a, *pb = (const orc_pat_t *)b;
    const char *s = pa->s, *t = pb->s;
    for (;;) {
        char sc = *s++, tc = *t++;
        int se = (sc == '\n'), te = (tc == '\n');
        if (se || te) {
            if (se && te) return (pa->id > pb->id) - (pa->id < pb->id);
            return se ? -1 : 1;
        }
        if (sc < tc) return -1; /* char is signed here, as in the reference build */
        if (sc > tc) return 1;
    }
}

static void orc_row_push(orc_row_t *r, int ch, int next)
{
    if (r->n == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 2;
        r->e = (orc_edge_t *)realloc(r->e, sizeof(orc_edge_t) * (size_t)r->cap);
    }
    r->e[r->n].ch = ch;
    r->e[r->n].next = next;
    r->n++;
}

/* lookup(): first edge with matching ch, src/PFAC_reorder_Table.cpp:234-244 */
static int orc_lookup(const orc_machine_t *m, int state, int ch)
{
    const orc_row_t *r = &m->rows[state];
    for (int j = 0; j < r->n; j++)
        if (r->e[j].ch == ch) return r->e[j].next;
    return ORC_TRAP;
}

void orc_free(orc_machine_t *m)
{
    if (!m) return;
    if (m->rows) {
        for (int i = 0; i < m->num_states; i++) free(m->rows[i].e);
        free(m->rows);
    }
    free(m->dense);
    free(m->sorted_ptr);
    free(m->sorted_id);
    free(m->len_by_id);
    free(m->buf);
    free(m);
}

/* Build from an in-memory file image.  Return codes mirror PFAC_status_t values
 * (include/PFAC.h:57-70): 0 ok, 10001 alloc, 10004 invalid parameter (blank line: the
 * reference aborts on assert at src/PFAC_reorder_Table.cpp:291; we return an error). */
int orc_build_from_memory(const char *image, long size, orc_machine_t **out)
{
    orc_machine_t *m = (orc_machine_t *)calloc(1, sizeof(*m));
    if (!m) return 10001;
    m->buf = (char *)malloc((size_t)size + 1);
    if (!m->buf) { free(m); return 10001; }
    memcpy(m->buf, image, (size_t)size);
    m->buf[size] = '\n';
    m->file_size = size;

    /* split on '\n' only; a line is a pattern iff its '\n' at i has i>0 && buf[i-1] != '\n'
     * (src/PFAC_reorder_Table.cpp:176-193).  Last line without '\n' is dropped. */
    int cap = 16, k = 0;
    orc_pat_t *pats = (orc_pat_t *)malloc(sizeof(orc_pat_t) * (size_t)cap);
    int *lens = (int *)malloc(sizeof(int) * (size_t)cap);
    long line_start = 0;
    int blank = 0, pending_blank = 0;
    for (long i = 0; i < size; i++) {
        if (m->buf[i] == '\n') {
            if (i > 0 && m->buf[i - 1] != '\n') {
                /* the reference leaves this pattern's pointer on the first blank line's '\n'
                 * when blank lines precede it, then asserts in the trie build (:291) */
                if (pending_blank) blank = 1;
                if (k == cap) {
                    cap *= 2;
                    pats = (orc_pat_t *)realloc(pats, sizeof(orc_pat_t) * (size_t)cap);
                    lens = (int *)realloc(lens, sizeof(int) * (size_t)cap);
                }
                pats[k].s = m->buf + line_start;
                pats[k].id = k + 1;
                lens[k] = (int)(i - line_start);
                k++;
            } else {
                pending_blank = 1; /* harmless if no pattern follows (trailing blank lines) */
            }
            line_start = i + 1;
        }
    }
    if (blank) { free(pats); free(lens); orc_free(m); return 10004; }

    m->num_patterns = k;
    m->len_by_id = (int *)calloc((size_t)k + 1, sizeof(int));
    for (int i = 0; i < k; i++) {
        m->len_by_id[i + 1] = lens[i];
        if (lens[i] > m->max_pattern_len) m->max_pattern_len = lens[i];
    }
    free(lens);
    qsort(pats, (size_t)k, sizeof(orc_pat_t), orc_pat_cmp);
    m->sorted_ptr = (char **)malloc(sizeof(char *) * (size_t)(k + 1));
    m->sorted_id = (int *)malloc(sizeof(int) * (size_t)(k + 1));
    for (int i = 0; i < k; i++) {
        m->sorted_ptr[i] = pats[i].s;
        m->sorted_id[i] = pats[i].id;
    }
    free(pats);

    /* trie: finals 1..k are pattern IDs, initial k+1, internal from k+2 in first-visit order
     * (src/PFAC.cpp:693,703; src/PFAC_reorder_Table.cpp:279-321) */
    m->num_final = k;
    m->initial_state = k + 1;
    long max_states = size + 2; /* reference bound is file_size+1; +1 keeps k==0 in range */
    m->rows = (orc_row_t *)calloc((size_t)max_states, sizeof(orc_row_t));
    int state_num = m->initial_state + 1;
    for (int p = 0; p < k; p++) {
        const char *pos = m->sorted_ptr[p];
        int id = m->sorted_id[p];
        int len = m->len_by_id[id];
        int state = m->initial_state;
        for (int off = 0; off < len; off++) {
            int ch = (unsigned char)pos[off];
            if (off == len - 1) {
                orc_row_push(&m->rows[state], ch, id); /* unconditional push, :293-298 */
            } else {
                int nx = orc_lookup(m, state, ch);
                if (nx == ORC_TRAP) {
                    orc_row_push(&m->rows[state], ch, state_num);
                    state = state_num++;
                } else {
                    state = nx;
                }
            }
        }
    }
    m->num_states = state_num;
    {   /* shrink rows to num_states so orc_free walks the right range */
        orc_row_t *r = (orc_row_t *)realloc(m->rows, sizeof(orc_row_t) * (size_t)state_num);
        if (r) m->rows = r;
    }

    /* dense table: all TRAP, then edges in state order / insertion order, last wins
     * (src/PFAC.cpp:371-381) */
    size_t cells = (size_t)m->num_states * ORC_CHAR_SET;
    m->dense = (int *)malloc(sizeof(int) * cells);
    if (!m->dense) { orc_free(m); return 10001; }
    for (size_t i = 0; i < cells; i++) m->dense[i] = ORC_TRAP;
    for (int s = 0; s < m->num_states; s++)
        for (int j = 0; j < m->rows[s].n; j++)
            m->dense[(size_t)s * ORC_CHAR_SET + m->rows[s].e[j].ch] = m->rows[s].e[j].next;

    *out = m;
    return 0;
}

/* PFAC_readPatternFromFile front half, src/PFAC.cpp:653-735 */
int orc_build_from_file(const char *path, orc_machine_t **out)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return 10006; /* PFAC_STATUS_FILE_OPEN_ERROR */
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    rewind(fp);
    char *img = (char *)malloc((size_t)sz + 1);
    if (!img) { fclose(fp); return 10001; }
    sz = (long)fread(img, 1, (size_t)sz, fp);
    fclose(fp);
    int rc = orc_build_from_memory(img, sz, out);
    free(img);
    return rc;
}

int orc_num_patterns(const orc_machine_t *m) { return m->num_patterns; }
int orc_num_states(const orc_machine_t *m) { return m->num_states; }
int orc_initial_state(const orc_machine_t *m) { return m->initial_state; }
int orc_max_pattern_len(const orc_machine_t *m) { return m->max_pattern_len; }
const int *orc_dense_table(const orc_machine_t *m) { return m->dense; }

/* PFAC_CPU_timeDriven, src/PFAC_CPU.cpp:60-100.  n is 64-bit here so that chunk drivers
 * are not needed for inputs >= 2^31; semantics are unchanged (the walk stops at n). */
void orc_match(const orc_machine_t *m, const unsigned char *in, int64_t n, int *out)
{
    const int *T = m->dense;
    const int nf = m->num_final, init = m->initial_state;
    for (int64_t i = 0; i < n; i++) out[i] = 0;
    for (int64_t start = 0; start < n; start++) {
        int state = init;
        int64_t pos = start;
        while (pos < n) {
            state = T[(size_t)state * ORC_CHAR_SET + in[pos]];
            if (state == ORC_TRAP) break;
            if (state <= nf) out[start] = state;
            pos++;
        }
    }
}

/* PFAC_CPU_OMP_timeDriven, src/PFAC_CPU_OMP.cpp:81-120: same loop, parallel over start */
void orc_match_omp(const orc_machine_t *m, const unsigned char *in, int64_t n, int *out)
{
    const int *T = m->dense;
    const int nf = m->num_final, init = m->initial_state;
    for (int64_t i = 0; i < n; i++) out[i] = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t start = 0; start < n; start++) {
        int state = init;
        int64_t pos = start;
        while (pos < n) {
            state = T[(size_t)state * ORC_CHAR_SET + in[pos]];
            if (state == ORC_TRAP) break;
            if (state <= nf) out[start] = state;
            pos++;
        }
    }
}

/* Shard form used by the multi-GPU parity tests: positions [0,n_owned) are reported, the
 * walk may read up to n_total (owned bytes + tail halo), SURVEY.md section 8(e). */
void orc_match_shard(const orc_machine_t *m, const unsigned char *in, int64_t n_owned,
                     int64_t n_total, int *out)
{
    const int *T = m->dense;
    const int nf = m->num_final, init = m->initial_state;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t start = 0; start < n_owned; start++) {
        int state = init, res = 0;
        int64_t pos = start;
        while (pos < n_total) {
            state = T[(size_t)state * ORC_CHAR_SET + in[pos]];
            if (state == ORC_TRAP) break;
            if (state <= nf) res = state;
            pos++;
        }
        out[start] = res;
    }
}

/* the reduce oracle: serial zip of the dense result, src/PFAC.cpp:1058-1068.
 * out_id/out_pos may alias `dense` exactly as the reference's in-place zip does. */
int64_t orc_reduce(const int *dense, int64_t n, int *out_id, int64_t *out_pos)
{
    int64_t z = 0;
    for (int64_t i = 0; i < n; i++) {
        int r = dense[i];
        if (0 < r) {
            out_id[z] = r;
            out_pos[z] = i;
            z++;
        }
    }
    return z;
}

/* printString, src/PFAC_reorder_Table.cpp:93-105 */
static void orc_print_string(const char *s, int n, FILE *fp)
{
    fprintf(fp, "%c", '\"');
    for (int i = 0; i < n; i++) {
        int ch = (unsigned char)s[i];
        if (32 <= ch && ch <= 126) fprintf(fp, "%c", ch);
        else fprintf(fp, "%2.2x", ch);
    }
    fprintf(fp, "%c", '\"');
}

/* PFAC_dumpTransitionTable, src/PFAC.cpp:1188-1246 */
int orc_dump(const orc_machine_t *m, const char *path)
{
    FILE *fp = fopen(path, "w");
    if (!fp) return 10006;
    fprintf(fp, "# Transition table: number of states = %d, initial state = %d\n", m->num_states,
            m->initial_state);
    fprintf(fp, "# (current state, input character) -> next state \n");
    for (int s = 0; s < m->num_states; s++) {
        for (int j = 0; j < m->rows[s].n; j++) {
            int ch = m->rows[s].e[j].ch, nx = m->rows[s].e[j].next;
            if (32 <= ch && ch <= 126) fprintf(fp, "(%4d,%4c) -> %d \n", s, ch, nx);
            else fprintf(fp, "(%4d,%4.2x) -> %d \n", s, ch, nx);
        }
    }
    char **by_id = (char **)calloc((size_t)m->num_final + 1, sizeof(char *));
    for (int i = 0; i < m->num_final; i++) by_id[m->sorted_id[i]] = m->sorted_ptr[i];
    fprintf(fp, "# Output table: number of final states = %d\n", m->num_final);
    fprintf(fp, "# [final state] [matched pattern ID] [pattern length] [pattern(string literal)] \n");
    for (int s = 1; s <= m->num_final; s++) {
        fprintf(fp, "%5d %5d %5d    ", s, s, m->len_by_id[s]);
        orc_print_string(by_id[s], m->len_by_id[s], fp);
        fprintf(fp, "\n");
    }
    free(by_id);
    fclose(fp);
    return 0;
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_omp_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
