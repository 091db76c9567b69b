/*
 * cuda_funcs.h -- header-level replacement for the reference's cuda_funcs.h, to be dropped into a
 * reference checkout next to its def.h / program_data.h / mutant.h when cuda_funcs.cu is REMOVED
 * from the build and libpsa_b200.so is linked instead.
 *
 * It declares exactly the host-visible functions that cpu_funcs.c uses out of the reference's
 * cuda_funcs.cu (cuda_funcs.h:33 and :44-61); libpsa_b200.so exports every one of them under the
 * same C++-mangled name (csrc/psa_capi.cpp), answering from the host-resolved pair table of
 * csrc/psa_table.cpp:
 *
 *     gpu_run_program      cuda_funcs.h:33   _Z15gpu_run_programP5_dataP7_mutantii
 *     get_substitute       cuda_funcs.h:44   _Z14get_substituteccPdi
 *     get_hashtable_sign   cuda_funcs.h:50   _Z18get_hashtable_signcc
 *     get_weight           cuda_funcs.h:51   _Z10get_weightcPd
 *     get_pair_sign        cuda_funcs.h:55   _Z13get_pair_signcc
 *     is_swapable          cuda_funcs.h:56   _Z11is_swapableP7_mutantS0_ddi
 *     strlen_gpu           cuda_funcs.h:61   _Z10strlen_gpuPc
 *
 * What the reference's header has and this one deliberately does not:
 *   - the absolute-path CUDA includes (cuda_funcs.h:4-5): nothing here needs the CUDA headers;
 *   - the __constant__ / __device__ array DEFINITIONS (cuda_funcs.h:21-23), which give every
 *     including translation unit its own copy: the device tables live inside the library;
 *   - the kernel prototypes and device-only helpers (cuda_funcs.h:35-42): replaced wholesale;
 *   - helpers cpu_funcs.c never calls (get_max_substitute ... is_power2).
 * The three extern tables of cuda_funcs.h:27-29 are still defined by cpu_funcs.c:18-20; the
 * library does not read them (its sign matrix is built from its own copy of the group lists), so
 * the fill_hash race of cpu_funcs.c:306-307 no longer reaches any score.
 *
 * Compiled as C++ (the reference builds every .c with mpicxx, Makefile:9-11): C++ linkage on purpose.
 */
#ifndef __CUDA_FUNCS_H__
#define __CUDA_FUNCS_H__

#include "def.h"
#include "program_data.h"
#include "mutant.h"

#define MAX_BLOCK_SIZE 1024

extern char conservatives_cpu[CONSERVATIVE_COUNT][CONSERVATIVE_MAX_LEN];
extern char semi_conservatives_cpu[SEMI_CONSERVATIVE_COUNT][SEMI_CONSERVATIVE_MAX_LEN];
extern char hashtable_cpu[NUM_CHARS][NUM_CHARS];

/* best score over absolute offsets [first_offset, last_offset) on the current CUDA device */
double gpu_run_program(ProgramData* cpu_data, Mutant* returned_mutant, int first_offset, int last_offset);

/* host-side scoring primitives (pure functions of their arguments) */
char   get_substitute(char c1, char c2, double* w, int is_max);
char   get_hashtable_sign(char c1, char c2);
double get_weight(char sign, double* w);
char   get_pair_sign(char a, char b);
int    is_swapable(Mutant* m1, Mutant* m2, double score1, double score2, int is_max);
int    strlen_gpu(char* str);

#endif /* __CUDA_FUNCS_H__ */
