/*
 * oracle/psa_oracle.c -- TEST INFRASTRUCTURE ONLY (see psa_oracle.h).
 *
 * CPU restatement of GuyKabiri/Parallel-Sequence-Alignment's mutant-offset search.
 * Written from the reference's behaviour, not from its text; every routine names the
 * reference lines it restates.  Parity status: PINNED (header of psa_oracle.h).
 *
 * Deliberate differences from the reference (all outside its defined behaviour):
 *   - symbols outside [A-Z-] are rejected (-1) instead of reading uninitialised
 *     doubles (cuda_funcs.cu:322-339);
 *   - the sign table is a full 27x27 matrix built once, single-threaded (the reference
 *     fills a lower triangle under a racy `omp parallel for`, cpu_funcs.c:304-318);
 *   - sequences are length-explicit (no 10000/5000 capacity, def.h:35-36).
 */
#include "psa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NSYM 27          /* 'A'..'Z' then '-' */
#define GAP  26

/* cpu_funcs.c:19-20 -- the ClustalW-style conservative / semi-conservative groups */
static const char* const k_conservative[] = { "NDEQ", "NEQK", "STA", "MILV", "QHRK", "NHQK", "FYW", "HY", "MILF" };
static const char* const k_semi[] = { "SAG", "ATV", "CSA", "SGND", "STPA", "STNK", "NEQHRK", "NDEQHK", "SNDEQK", "HFY", "FVLIM" };

static char g_sign[NSYM][NSYM];
static int  g_sign_ready = 0;

static int sym_index(char c)
{
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c == '-') return GAP;
    return -1;
}

static int share_group(const char* const* groups, int ngroups, char a, char b)
{
    for (int g = 0; g < ngroups; g++)
        if (strchr(groups[g], a) && strchr(groups[g], b)) return 1;
    return 0;
}

/* cuda_funcs.cu:495-502 (letters) + cuda_funcs.cu:426-427 (gap rules) */
static void build_sign_table(void)
{
    for (int a = 0; a < 26; a++)
        for (int b = 0; b < 26; b++) {
            char ca = (char)('A' + a), cb = (char)('A' + b), s;
            if (a == b) s = '*';
            else if (share_group(k_conservative, 9, ca, cb)) s = ':';
            else if (share_group(k_semi, 11, ca, cb)) s = '.';
            else s = '_';
            g_sign[a][b] = s;
        }
    for (int a = 0; a < NSYM; a++) { g_sign[a][GAP] = '_'; g_sign[GAP][a] = '_'; }
    g_sign[GAP][GAP] = '*';
    g_sign_ready = 1;
}

static inline void ensure_sign_table(void)
{
    if (!g_sign_ready) {
#ifdef _OPENMP
#pragma omp critical(psa_oracle_sign_init)
#endif
        if (!g_sign_ready) build_sign_table();
    }
}

char psa_oracle_sign(char c1, char c2)
{
    ensure_sign_table();
    int a = sym_index(c1), b = sym_index(c2);
    if (a < 0 || b < 0) return '\0';
    return g_sign[a][b];
}

double psa_oracle_weight(char sign, const double* w)
{
    switch (sign) {
    case '*': return w[0];
    case ':': return -w[1];
    case '.': return -w[2];
    case '_': return -w[3];
    }
    return 0;
}

/* cuda_funcs.cu:412-421: first letter (A..Z order) that makes `want` against `by`
   and is not conservative (':') with the letter being replaced */
static char first_letter_with(char by, char want, char replaced)
{
    for (char ch = 'A'; ch <= 'Z'; ch++)
        if (psa_oracle_sign(by, ch) == want && psa_oracle_sign(replaced, ch) != ':')
            return ch;
    return '\0';
}

/* cuda_funcs.cu:396-409 */
static char pick_of_two(int is_max, double d1, char s1, double d2, char s2)
{
    int first_not_worse = is_max ? (d1 >= d2) : (d1 <= d2);
    if (first_not_worse && s1 != '\0') return s1;
    if (s2 != '\0') return s2;
    return s1;
}

char psa_oracle_substitute(char c1, char c2, const double* w, int is_max)
{
    char sign = psa_oracle_sign(c1, c2);
    if (sign == '\0') return '\0';
    char to_colon = first_letter_with(c1, ':', c2);
    char to_dot   = first_letter_with(c1, '.', c2);
    char to_space = first_letter_with(c1, '_', c2);

    if (is_max) {
        /* cuda_funcs.cu:320-345 */
        if (sign == '.' || sign == '_') return c1;
        double d_dot, d_space;
        if (sign == '*') { d_dot = -w[0] - w[2]; d_space = -w[0] - w[3]; }
        else             { d_dot =  w[1] - w[2]; d_space =  w[1] - w[3]; }
        return pick_of_two(1, d_dot, to_dot, d_space, to_space);
    }

    /* cuda_funcs.cu:348-393 */
    double d1, d2; char s1, s2;
    switch (sign) {
    case '*': d1 = -w[0] - w[2]; s1 = to_dot;   d2 = -w[0] - w[3]; s2 = to_space; break;
    case ':': d1 =  w[1] - w[2]; s1 = to_dot;   d2 =  w[1] - w[3]; s2 = to_space; break;
    case '.': d1 =  w[2] - w[1]; s1 = to_colon; d2 =  w[2] - w[3]; s2 = to_space; break;
    default:  d1 =  w[3] - w[1]; s1 = to_colon; d2 =  w[3] - w[2]; s2 = to_dot;   break;
    }
    char sub = pick_of_two(0, d1, s1, d2, s2);
    if ((sign == '.' || sign == '_') && sub == '\0') return c1;
    return sub;
}

int psa_oracle_is_swapable(int off1, int coff1, int off2, int coff2, double s1, double s2, int is_max)
{
    if ((is_max && s2 > s1) || (!is_max && s2 < s1)) return 1;
    if (s2 == s1) {
        if (off2 < off1) return 1;
        if (off2 == off1 && coff2 < coff1) return 1;
    }
    return 0;
}

static int class_of_sign(char s) { return s == '*' ? 0 : s == ':' ? 1 : s == '.' ? 2 : 3; }

int psa_oracle_offset_naive(const double* w, int is_max, const char* seq1, const char* seq2, long len2,
                            long offset, psa_oracle_result* out)
{
    /* cpu_funcs.c:257-300, evaluated pair by pair */
    double total = 0, best = is_max ? -INFINITY : INFINITY;
    out->offset = -1; out->char_offset = -1; out->ch = '\0';
    memset(out->counts, 0, sizeof(out->counts));
    for (long i = 0; i < len2; i++) {
        char c1 = seq1[offset + i], c2 = seq2[i];
        char sg = psa_oracle_sign(c1, c2);
        if (sg == '\0') return -1;
        double ps = psa_oracle_weight(sg, w);
        total += ps;
        out->counts[class_of_sign(sg)]++;
        char sub = psa_oracle_substitute(c1, c2, w, is_max);
        if (sub == '\0') continue;
        double d = psa_oracle_weight(psa_oracle_sign(c1, sub), w) - ps;
        if ((is_max && d > best) || (!is_max && d < best)) {
            best = d; out->ch = sub; out->char_offset = (int)i; out->offset = (int)offset;
        }
    }
    out->score = (out->ch == '\0') ? best : total + best;
    return 0;
}

/* per-problem fold of the pure per-pair functions: for (c1,c2): pair weight, substitute, diff */
typedef struct pair_entry { double ps; double diff; char sub; unsigned char cls; } pair_entry;

static void build_pair_table(const double* w, int is_max, pair_entry t[NSYM][NSYM])
{
    ensure_sign_table();
    for (int a = 0; a < NSYM; a++)
        for (int b = 0; b < NSYM; b++) {
            char c1 = a < 26 ? (char)('A' + a) : '-', c2 = b < 26 ? (char)('A' + b) : '-';
            char sg = g_sign[a][b];
            pair_entry* e = &t[a][b];
            e->ps = psa_oracle_weight(sg, w);
            e->cls = (unsigned char)class_of_sign(sg);
            e->sub = psa_oracle_substitute(c1, c2, w, is_max);
            e->diff = e->sub ? psa_oracle_weight(psa_oracle_sign(c1, e->sub), w) - e->ps : 0.0;
        }
}

/* cpu_funcs.c:257-300 with the table; same accumulation order, same strict compares */
static void offset_with_table(pair_entry t[NSYM][NSYM], int is_max, const unsigned char* s1, const unsigned char* s2,
                              long len2, long offset, psa_oracle_result* out, int want_counts)
{
    double total = 0, best = is_max ? -INFINITY : INFINITY;
    int coff = -1; char ch = '\0';
    const unsigned char* p1 = s1 + offset;
    if (is_max) {
        for (long i = 0; i < len2; i++) {
            const pair_entry* e = &t[p1[i]][s2[i]];
            total += e->ps;
            if (e->sub && e->diff > best) { best = e->diff; ch = e->sub; coff = (int)i; }
        }
    } else {
        for (long i = 0; i < len2; i++) {
            const pair_entry* e = &t[p1[i]][s2[i]];
            total += e->ps;
            if (e->sub && e->diff < best) { best = e->diff; ch = e->sub; coff = (int)i; }
        }
    }
    out->ch = ch; out->char_offset = coff; out->offset = ch ? (int)offset : -1;
    out->score = ch ? total + best : best;
    if (want_counts) {
        memset(out->counts, 0, sizeof(out->counts));
        for (long i = 0; i < len2; i++) out->counts[t[p1[i]][s2[i]].cls]++;
    }
}

static int to_indices(const char* s, long n, unsigned char* out)
{
    for (long i = 0; i < n; i++) {
        int k = sym_index(s[i]);
        if (k < 0) return -1;
        out[i] = (unsigned char)k;
    }
    return 0;
}

/* cpu_funcs.c:222-243: ascending offsets, replace incumbent iff is_swapable */
static void range_best(pair_entry t[NSYM][NSYM], int is_max, const unsigned char* s1, const unsigned char* s2,
                       long len2, long first, long last, psa_oracle_result* best)
{
    best->offset = -1; best->char_offset = -1; best->ch = '\0';
    best->score = is_max ? -INFINITY : INFINITY;
    int have = 0;
    for (long n = first; n < last; n++) {
        psa_oracle_result cur;
        offset_with_table(t, is_max, s1, s2, len2, n, &cur, 0);
        if (!have || psa_oracle_is_swapable(best->offset, best->char_offset, cur.offset, cur.char_offset,
                                            best->score, cur.score, is_max)) {
            /* (!have) stands in for the reference's first compare against +-inf: any finite score wins it */
            if (!have && cur.ch == '\0') continue;
            *best = cur; have = 1;
        }
    }
}

static int search_indexed(pair_entry t[NSYM][NSYM], int is_max, const unsigned char* s1, const unsigned char* s2,
                          long len2, long first, long last, int nthreads, psa_oracle_result* out)
{
    long n = last - first;
    if (nthreads < 1) nthreads = 1;
    if ((long)nthreads > n) nthreads = (int)n;
    psa_oracle_result* part = (psa_oracle_result*)malloc(sizeof(psa_oracle_result) * (size_t)nthreads);
    if (!part) return -3;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
#endif
    for (int k = 0; k < nthreads; k++) {
        long a = first + n * k / nthreads, b = first + n * (k + 1) / nthreads;
        range_best(t, is_max, s1, s2, len2, a, b, &part[k]);
    }
    /* cpu_funcs.c:245-251: merge under the same order */
    psa_oracle_result best = part[0];
    for (int k = 1; k < nthreads; k++) {
        if (part[k].ch == '\0') continue;
        if (best.ch == '\0' || psa_oracle_is_swapable(best.offset, best.char_offset, part[k].offset, part[k].char_offset,
                                                      best.score, part[k].score, is_max))
            best = part[k];
    }
    free(part);
    if (best.ch != '\0') {
        psa_oracle_result again;
        offset_with_table(t, is_max, s1, s2, len2, best.offset, &again, 1);
        memcpy(best.counts, again.counts, sizeof(best.counts));
    } else {
        memset(best.counts, 0, sizeof(best.counts));
    }
    *out = best;
    return 0;
}

int psa_oracle_search(const double* w, int is_max, const char* seq1, long len1, const char* seq2, long len2,
                      long first, long last, int nthreads, psa_oracle_result* out)
{
    if (len2 < 1 || len1 < len2 || first < 0 || last > len1 - len2 + 1 || first >= last) return -2;
    unsigned char* s1 = (unsigned char*)malloc((size_t)len1);
    unsigned char* s2 = (unsigned char*)malloc((size_t)len2);
    if (!s1 || !s2) { free(s1); free(s2); return -3; }
    int rc = 0;
    if (to_indices(seq1, len1, s1) || to_indices(seq2, len2, s2)) rc = -1;
    else {
        pair_entry t[NSYM][NSYM];
        build_pair_table(w, is_max, t);
        rc = search_indexed(t, is_max, s1, s2, len2, first, last, nthreads, out);
    }
    free(s1); free(s2);
    return rc;
}

int psa_oracle_search_batch(const double* w, int is_max, const char* seq1, long len1,
                            const char* seq2s, const long long* q_off, int nq, int nthreads,
                            psa_oracle_result* out)
{
    if (nq < 0 || len1 < 1) return -2;
    unsigned char* s1 = (unsigned char*)malloc((size_t)len1);
    if (!s1) return -3;
    if (to_indices(seq1, len1, s1)) { free(s1); return -1; }
    pair_entry t[NSYM][NSYM];
    build_pair_table(w, is_max, t);
    int rc = 0;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 4)
#endif
    for (int q = 0; q < nq; q++) {
        long len2 = (long)(q_off[q + 1] - q_off[q]);
        int r = 0;
        if (len2 < 1 || len2 > len1) r = -2;
        else {
            unsigned char* s2 = (unsigned char*)malloc((size_t)len2);
            if (!s2) r = -3;
            else if (to_indices(seq2s + q_off[q], len2, s2)) r = -1;
            else r = search_indexed(t, is_max, s1, s2, len2, 0, len1 - len2 + 1, 1, &out[q]);
            free(s2);
        }
        if (r) {
#ifdef _OPENMP
#pragma omp critical(psa_oracle_batch_rc)
#endif
            rc = r;
        }
    }
    free(s1);
    return rc;
}

int psa_oracle_scores(const double* w, int is_max, const char* seq1, long len1, const char* seq2, long len2,
                      long first, long last, double* scores)
{
    if (len2 < 1 || len1 < len2 || first < 0 || last > len1 - len2 + 1 || first > last) return -2;
    unsigned char* s1 = (unsigned char*)malloc((size_t)len1);
    unsigned char* s2 = (unsigned char*)malloc((size_t)len2);
    if (!s1 || !s2) { free(s1); free(s2); return -3; }
    int rc = 0;
    if (to_indices(seq1, len1, s1) || to_indices(seq2, len2, s2)) rc = -1;
    else {
        pair_entry t[NSYM][NSYM];
        build_pair_table(w, is_max, t);
        for (long n = first; n < last; n++) {
            psa_oracle_result cur;
            offset_with_table(t, is_max, s1, s2, len2, n, &cur, 0);
            scores[n - first] = cur.score;
        }
    }
    free(s1); free(s2);
    return rc;
}
