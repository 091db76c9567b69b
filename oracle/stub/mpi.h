/*
 * oracle/stub/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Single-process stand-in for <mpi.h>, just enough to let the reference's
 * cpu_funcs.c / mpi_funcs.h compile without an MPI installation (the image has
 * none).  Every call behaves as "world of one rank".  The oracle harness never
 * enters the MPI branches of the reference (they are all guarded by
 * num_processes > 1), so the bodies only need to link.
 */
#ifndef PSA_ORACLE_STUB_MPI_H
#define PSA_ORACLE_STUB_MPI_H

#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int  MPI_Datatype;
typedef int  MPI_Comm;
typedef int  MPI_Op;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD          0
#define MPI_SUCCESS             0
#define MPI_INT                 1
#define MPI_DOUBLE              2
#define MPI_CHAR                3
#define MPI_2DOUBLE_PRECISION   4
#define MPI_MAXLOC              1
#define MPI_MINLOC              2

static inline int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }
static inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void* in, void* out, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)op; (void)c; if (t == MPI_2DOUBLE_PRECISION) memcpy(out, in, (size_t)n * 2 * sizeof(double)); return MPI_SUCCESS; }
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; return MPI_SUCCESS; }
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s; return MPI_SUCCESS; }
static inline int MPI_Type_create_struct(int n, const int* bl, const MPI_Aint* d, const MPI_Datatype* t, MPI_Datatype* out)
{ (void)n; (void)bl; (void)d; (void)t; *out = 100; return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype* t) { (void)t; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype* t) { (void)t; return MPI_SUCCESS; }
static inline double MPI_Wtime(void)
{ struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

#endif /* PSA_ORACLE_STUB_MPI_H */
