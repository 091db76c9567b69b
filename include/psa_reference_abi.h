/*
 * psa_reference_abi.h -- the two records of the reference that cross the drop-in boundary,
 * restated with the same tags and layout so that the C++-mangled entry point
 *     double gpu_run_program(ProgramData*, Mutant*, int, int)      (cuda_funcs.h:33)
 *     -> _Z15gpu_run_programP5_dataP7_mutantii
 * exported by libpsa_b200.so links against an unmodified cpu_funcs.o of the reference.
 *
 *   struct _data   == ProgramData  program_data.h:6-11  (capacities def.h:35-36)
 *   struct _mutant == Mutant       mutant.h:6-10
 *
 * Capacities may be raised at build time (-DPSA_SEQ1_CAPACITY=...), which the reference's own
 * unguarded macros do not allow; the defaults are the reference's.
 */
#ifndef PSA_REFERENCE_ABI_H
#define PSA_REFERENCE_ABI_H

#ifndef PSA_SEQ1_CAPACITY
#define PSA_SEQ1_CAPACITY 10000
#endif
#ifndef PSA_SEQ2_CAPACITY
#define PSA_SEQ2_CAPACITY 5000
#endif

typedef struct _data {
    int    is_max;                          /* 1 = "maximum", 0 = anything else (cpu_funcs.c:365) */
    double weights[4];                      /* W1 '*', W2 ':', W3 '.', W4 '_'                     */
    char   seq1[PSA_SEQ1_CAPACITY + 1];     /* NUL-terminated                                     */
    char   seq2[PSA_SEQ2_CAPACITY + 1];
} ProgramData;

typedef struct _mutant {
    int  offset;
    int  char_offset;
    char ch;
} Mutant;

#ifdef __cplusplus
/* the symbol the reference's cpu_funcs.c (compiled as C++ by mpicxx, Makefile:9-11) imports */
double gpu_run_program(ProgramData* cpu_data, Mutant* returned_mutant, int first_offset, int last_offset);
/* ... and the six host primitives it imports from cuda_funcs.cu (cuda_funcs.h:44-61); see include/cuda_funcs.h */
char   get_substitute(char c1, char c2, double* w, int is_max);
char   get_hashtable_sign(char c1, char c2);
double get_weight(char sign, double* w);
char   get_pair_sign(char a, char b);
int    is_swapable(Mutant* m1, Mutant* m2, double score1, double score2, int is_max);
int    strlen_gpu(char* str);
#endif

#endif /* PSA_REFERENCE_ABI_H */
