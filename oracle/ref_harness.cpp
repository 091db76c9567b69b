/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin extern "C" shim around the UNMODIFIED reference sources, which are compiled
 * where they lie under /root/reference by oracle/Makefile (target `ref`) into
 * oracle/_ref/libpsa_ref.so.  Nothing from the reference is copied into this repo;
 * this file only *calls* it:
 *
 *   fill_hash                  cpu_funcs.c:304-318
 *   find_best_mutant_cpu       cpu_funcs.c:222-253
 *   find_best_mutant_offset    cpu_funcs.c:257-300
 *   divide_execute_tasks       cpu_funcs.c:123-218
 *   get_hashtable_sign / get_weight / get_substitute / get_pair_sign / is_swapable
 *                              cuda_funcs.cu:290-502 (host side of __host__ __device__)
 *
 * Used by tests/ (to pin the C restatement in oracle/psa_oracle.c) and by
 * bench.py's cpu_baseline / --impl reference legs.  Capacity is the reference's own:
 * Seq1 <= 10000, Seq2 <= 5000 (def.h:35-36).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <omp.h>
#include <mpi.h>
#include <thread>
#include <cuda_runtime_api.h>

#include "def.h"
#include "program_data.h"
#include "mutant.h"
#include "cpu_funcs.h"

/* Host-side prototypes of the reference's scoring primitives (cuda_funcs.h:44-56).
   cuda_funcs.h itself cannot be included here: it DEFINES the __constant__/__device__
   arrays (cuda_funcs.h:21-23), so a third TU including it would clash at link time. */
char get_substitute(char c1, char c2, double* w, int is_max);
char get_hashtable_sign(char c1, char c2);
double get_weight(char sign, double* w);
char get_pair_sign(char a, char b);
int is_swapable(Mutant* m1, Mutant* m2, double score1, double score2, int is_max);

/* globals the reference expects some other TU to define (main.h:6, mpi_funcs.c:7-8) */
int cuda_percentage = 0;
MPI_Datatype mutant_type = 0;
MPI_Datatype program_data_type = 0;

static int fill_problem(ProgramData* d, const double* w, int is_max, const char* seq1, const char* seq2)
{
    size_t l1 = strlen(seq1), l2 = strlen(seq2);
    if (l1 >= sizeof(d->seq1) || l2 >= sizeof(d->seq2)) return -1;
    memset(d, 0, sizeof(*d));
    d->is_max = is_max ? MAXIMUM_FUNC : MINIMUM_FUNC;
    for (int i = 0; i < WEIGHTS_COUNT; i++) d->weights[i] = w[i];
    memcpy(d->seq1, seq1, l1 + 1);
    memcpy(d->seq2, seq2, l2 + 1);
    return 0;
}

static void fill_hash_single_thread(double* w)
{
    /* the reference's fill_hash races on c1/c2 when run by >1 thread (shared
       loop temporaries); one thread is the deterministic behaviour */
    int saved = omp_get_max_threads();
    omp_set_num_threads(1);
    fill_hash(w, 0);
    omp_set_num_threads(saved);
}

extern "C" {

int ref_seq1_capacity(void) { return (int)sizeof(((ProgramData*)0)->seq1) - 1; }
int ref_seq2_capacity(void) { return (int)sizeof(((ProgramData*)0)->seq2) - 1; }
int ref_sizeof_program_data(void) { return (int)sizeof(ProgramData); }
int ref_sizeof_mutant(void) { return (int)sizeof(Mutant); }

/* Correctness oracle: the reference's 1-thread path ("-100" argument, main.c:33-37):
   fill_hash + find_best_mutant_cpu over [first,last) on one thread. */
double ref_search_seq(const double* w, int is_max, const char* seq1, const char* seq2,
                      int first, int last, int* out_offset, int* out_char_offset, char* out_ch)
{
    static ProgramData d;
    if (fill_problem(&d, w, is_max, seq1, seq2)) return NAN;
    fill_hash_single_thread(d.weights);
    Mutant m = { -1, -1, NOT_FOUND_CHAR };
    double score = d.is_max ? -INFINITY : INFINITY;
    if (last > first)
        find_best_mutant_cpu(0, &d, &m, first, last, &score);
    *out_offset = m.offset; *out_char_offset = m.char_offset; *out_ch = m.ch;
    return score;
}

/* The reference's own thread split of find_best_mutant_cpu (cpu_funcs.c:192-197) for
   any thread count, with the table filled race-free first.  Deterministic. */
double ref_search_omp(const double* w, int is_max, const char* seq1, const char* seq2,
                      int nthreads, int* out_offset, int* out_char_offset, char* out_ch)
{
    static ProgramData d;
    if (fill_problem(&d, w, is_max, seq1, seq2)) return NAN;
    fill_hash_single_thread(d.weights);
    int chars = (int)strlen(d.seq2);
    int offsets = (int)strlen(d.seq1) - chars + 1;
    Mutant m = { -1, -1, NOT_FOUND_CHAR };
    double score = d.is_max ? -INFINITY : INFINITY;
    if (offsets <= 0) { *out_offset = -1; *out_char_offset = -1; *out_ch = 0; return score; }
    if (nthreads > offsets) nthreads = offsets;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        int nt = omp_get_num_threads();
        int tid = omp_get_thread_num();
        int per = offsets / nt;
        int start = per * tid;
        int end = start + per;
        if (tid == nt - 1) end += offsets % nt;
        if (end > start)
            find_best_mutant_cpu(0, &d, &m, start, end, &score);
    }
    *out_offset = m.offset; *out_char_offset = m.char_offset; *out_ch = m.ch;
    return score;
}

/* As shipped: divide_execute_tasks with cuda_percentage forced (0 = OpenMP only).
   With >1 thread this inherits the reference's fill_hash race: timing baseline only. */
double ref_divide_execute_tasks(const double* w, int is_max, const char* seq1, const char* seq2,
                                int num_processes, int pid, int pct, int nthreads,
                                int* out_offset, int* out_char_offset, char* out_ch)
{
    static ProgramData d;
    if (fill_problem(&d, w, is_max, seq1, seq2)) return NAN;
    cuda_percentage = pct;
    omp_set_num_threads(nthreads);
    Mutant m = { -1, -1, NOT_FOUND_CHAR };
    double s = divide_execute_tasks(&d, num_processes, pid, &m);
    *out_offset = m.offset; *out_char_offset = m.char_offset; *out_ch = m.ch;
    return s;
}

/* EMULATION of the reference's "mpiexec -np 2" CUDA+OpenMP run (MPI is not installed in this image): two host threads
   stand in for the two ranks.  Each selects its own GPU (rank % ndev, as two ranks on two one-GPU machines would each
   have one) and calls the reference's own divide_execute_tasks(&data, 2, pid, &mutant) (cpu_funcs.c:123-218) with the
   CUDA percentage at `pct`; the two (score, rank) pairs are then reduced the way MPI_MAXLOC / MPI_MINLOC do it
   (cpu_funcs.c:64-77: better score wins, ties go to the lower rank) and the winner's Mutant is "sent" to rank 0
   (cpu_funcs.c:82-94).  rank_scores[2] receives the per-rank scores.  Timing baseline only: the reference's GPU path
   races across blocks. */
double ref_np2(const double* w, int is_max, const char* seq1, const char* seq2, int pct, int ndev, int nthreads,
               int* out_offset, int* out_char_offset, char* out_ch, double* rank_scores)
{
    ProgramData* d = new ProgramData[2];
    Mutant m[2] = { { -1, -1, NOT_FOUND_CHAR }, { -1, -1, NOT_FOUND_CHAR } };
    double s[2] = { NAN, NAN };
    if (fill_problem(&d[0], w, is_max, seq1, seq2)) { delete[] d; return NAN; }
    d[1] = d[0];
    cuda_percentage = pct;
    omp_set_num_threads(nthreads);
    auto rank = [&](int pid) {
        if (ndev > 0) cudaSetDevice(pid % ndev);
        omp_set_num_threads(nthreads);
        s[pid] = divide_execute_tasks(&d[pid], 2, pid, &m[pid]);
    };
    std::thread t1(rank, 1);
    rank(0);
    t1.join();
    const int win = (is_max ? s[1] > s[0] : s[1] < s[0]) ? 1 : 0;     /* MAXLOC / MINLOC: ties -> lowest rank */
    *out_offset = m[win].offset; *out_char_offset = m[win].char_offset; *out_ch = m[win].ch;
    if (rank_scores) { rank_scores[0] = s[0]; rank_scores[1] = s[1]; }
    const double best = s[win];
    delete[] d;
    return best;
}

/* One offset (cpu_funcs.c:257-300). Requires ref_fill_hash() first. */
double ref_offset_score(const double* w, int is_max, const char* seq1, const char* seq2, int offset,
                        int* out_char_offset, char* out_ch)
{
    static ProgramData d;
    if (fill_problem(&d, w, is_max, seq1, seq2)) return NAN;
    Mutant m;
    double s = find_best_mutant_offset(&d, offset, &m);
    *out_char_offset = m.char_offset; *out_ch = m.ch;
    return s;
}

void ref_fill_hash(void) { double w[4] = { 0, 0, 0, 0 }; fill_hash_single_thread(w); }
char ref_sign(char c1, char c2) { return get_hashtable_sign(c1, c2); }
char ref_pair_sign(char a, char b) { return get_pair_sign(a, b); }
double ref_weight(char sign, const double* w) { double ww[4] = { w[0], w[1], w[2], w[3] }; return get_weight(sign, ww); }
char ref_substitute(char c1, char c2, const double* w, int is_max)
{ double ww[4] = { w[0], w[1], w[2], w[3] }; return get_substitute(c1, c2, ww, is_max); }
int ref_is_swapable(int off1, int coff1, int off2, int coff2, double s1, double s2, int is_max)
{ Mutant a = { off1, coff1, 'A' }, b = { off2, coff2, 'A' }; return is_swapable(&a, &b, s1, s2, is_max); }

/* File I/O of the reference (cpu_funcs.c:353-378) for the format tests. */
int ref_read_input(const char* path, double* w, int* is_max, char* seq1, char* seq2)
{
    static ProgramData d;
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    ProgramData* r = read_seq_and_weights_from_file(f, &d);
    fclose(f);
    if (!r) return -2;
    for (int i = 0; i < 4; i++) w[i] = d.weights[i];
    *is_max = d.is_max;
    strcpy(seq1, d.seq1); strcpy(seq2, d.seq2);
    return 0;
}
int ref_write_output(const char* path, const char* mutant, int offset, double score)
{
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    int ok = write_results_to_file(f, (char*)mutant, offset, score);
    fclose(f);
    return ok ? 0 : -2;
}

} /* extern "C" */
