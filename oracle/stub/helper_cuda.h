/*
 * oracle/stub/helper_cuda.h -- TEST INFRASTRUCTURE ONLY.
 * The reference's cuda_funcs.cu includes <helper_cuda.h> (a CUDA-samples header
 * that is not in this image) but uses nothing from it; it only relies on the
 * libc headers that file happens to pull in.
 */
#ifndef PSA_ORACLE_STUB_HELPER_CUDA_H
#define PSA_ORACLE_STUB_HELPER_CUDA_H
#include <stdio.h>
#include <stdlib.h>
#include <limits.h>
#endif
