// psa_main.cpp -- command-line drop-in for the reference's `mpiCudaOpenMP` binary (main.c:13-56):
// reads ./input.txt, runs the mutant-offset search on the GPU(s), writes ./output.txt and prints
// the same three stdout lines.  argv[1] keeps the reference's meaning (CUDA percentage, -100 =
// sequential CPU) but is only echoed: there is no CPU path here, every offset runs on the GPU.
//
//   psa_b200_cli [cuda_percentage] [--gpus N] [--input PATH] [--output PATH] [--all-blocks] [--queries FILE]
// --all-blocks: consume every problem block stacked in the input file (the reference reads only the first)
// --queries FILE: weights, Seq1 and goal from the input file, the queries from FILE (FASTA, or one query per token);
//                 one output stanza per query
#include "psa_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char** argv)
{
    const char* in = "./input.txt";     // def.h:20
    const char* out = "./output.txt";   // def.h:21
    int gpus = 1;
    int percentage = 100;
    bool all_blocks = false;
    const char* queries = nullptr;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--input") && i + 1 < argc) in = argv[++i];
        else if (!std::strcmp(argv[i], "--output") && i + 1 < argc) out = argv[++i];
        else if (!std::strcmp(argv[i], "--all-blocks")) all_blocks = true;
        else if (!std::strcmp(argv[i], "--queries") && i + 1 < argc) queries = argv[++i];
        else percentage = std::atoi(argv[i]);
    }
    (void)percentage;
    psa_context* ctx = nullptr;
    int rc = psa_create(&ctx, nullptr, gpus);
    if (rc) {
        std::fprintf(stderr, "Failed to initialise %d GPU(s) - %s\n", gpus, psa_strerror(rc));
        return EXIT_FAILURE;
    }
    std::printf("threads=%2d, processes=%2d\n", 1, gpus);         // cpu_funcs.c:55 (GPUs stand in for ranks)
    std::printf("CUDA percentage set to %d\n", 100);              // cpu_funcs.c:153
    auto t0 = std::chrono::steady_clock::now();
    psa_result r;
    int blocks = 1;
    int32_t nq = 0;
    rc = queries ? psa_run_query_file(ctx, in, queries, out, &nq) : all_blocks ? psa_run_files_all(ctx, in, out, &blocks) : psa_run_files(ctx, in, out, &r);
    auto t1 = std::chrono::steady_clock::now();
    if (rc) {
        if (rc == PSA_ERR_IO) std::printf("Error reading input file `%s` or writing `%s`\n", in, out);   // cpu_funcs.c:37,43,103
        else std::fprintf(stderr, "search failed - %s (%s)\n", psa_strerror(rc), psa_last_error(ctx));
        psa_destroy(ctx);
        return 2;                                                  // MPI_Abort(..., 2)
    }
    std::printf("total time: %g\n", std::chrono::duration<double>(t1 - t0).count());   // main.c:47
    psa_destroy(ctx);
    return 0;
}
