/*
 * psa_b200.h -- C ABI of the B200-native mutant-offset search.
 *
 * This is the drop-in boundary for the one hot path of
 * GuyKabiri/Parallel-Sequence-Alignment: everything behind
 *     double gpu_run_program(ProgramData*, Mutant*, int first, int last)   (cuda_funcs.h:33)
 * i.e. cuda_funcs.cu:6-278 (driver + 3 kernels) and the per-pair primitives it runs on the
 * device (cuda_funcs.cu:290-502), plus the orchestration around it that north_star replaces:
 * the rank/CPU/GPU offset split of divide_execute_tasks (cpu_funcs.c:123-218) and the
 * MPI MAXLOC/MINLOC reduce of initiate_program (cpu_funcs.c:64-94).
 *
 * Plain C: pointers and sizes only, no torch / CUDA types in any signature
 * (a CUDA stream, where one is accepted, is passed as void*).
 * All entry points are thread-compatible (one call at a time per context).
 * There is NO CPU fallback: every search entry point fails with PSA_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef PSA_B200_H
#define PSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSA_ABI_VERSION 1

/* ---- status codes (the reference has none: it exit()s, cuda_funcs.cu:44-48) ------------- */
enum {
    PSA_OK            = 0,
    PSA_ERR_ARG       = -1,  /* NULL pointer, nq < 0, len2 < 1, len2 > len1, empty offset range */
    PSA_ERR_ALPHABET  = -2,  /* a symbol outside [A-Z-] (undefined behaviour in the reference)  */
    PSA_ERR_WEIGHTS   = -3,  /* NaN / Inf weight                                               */
    PSA_ERR_CUDA      = -4,  /* no device / CUDA runtime failure (see psa_last_error)          */
    PSA_ERR_NOMEM     = -5,
    PSA_ERR_STATE     = -6,  /* psa_batch_run / psa_batch_fetch without a prepared batch       */
    PSA_ERR_IO        = -7   /* input.txt / output.txt could not be read / written             */
};

/* ---- records ---------------------------------------------------------------------------- */

/* Same layout as the reference's Mutant {int offset; int char_offset; char ch;} (mutant.h:6-10). */
typedef struct psa_mutant {
    int32_t offset;        /* absolute offset n of Seq2 under Seq1; -1 if no mutation exists */
    int32_t char_offset;   /* index i in Seq2 of the replaced character; -1 if none          */
    char    ch;            /* replacement letter; '\0' if none                               */
} psa_mutant;

/* One result per query. score is the reference's double (cpu_funcs.c:297-299):
   sum_i w(sign(Seq1[n+i],Seq2[i])) + best single-substitution difference; +-INFINITY if none. */
typedef struct psa_result {
    psa_mutant mutant;
    int32_t    rank;        /* internal: dense rank of the winning substitution difference     */
    double     score;
    int64_t    counts[4];   /* N('*'), N(':'), N('.'), N('_') at the winning offset, unmutated */
} psa_result;

/* Host-resolved 27x27 pair table (replaces fill_hash cpu_funcs.c:304-318 and the per-pair
   get_substitute search cuda_funcs.cu:310-421).  Row = Seq2 symbol, column = Seq1 symbol;
   symbol index 0..25 = 'A'..'Z', 26 = '-'. */
typedef struct psa_pair_table {
    char    sign[27][27];      /* '*' ':' '.' '_'                                        */
    char    substitute[27][27];/* best replacement for the Seq2 symbol, '\0' if none      */
    double  diff[27][27];      /* w(sign(c1,substitute)) - w(sign(c1,c2)) as the reference computes it */
    uint8_t rank[27][27];      /* 0 = no substitute; 1..nranks, higher = better for the goal */
    int32_t nranks;
    int32_t exact;             /* 1: every partial sum is exact in double and int64 fixed point */
    int32_t frac_bits;         /* fixed-point scale 2^frac_bits used for device score keys      */
    int64_t key_slack;         /* candidate window (key units) re-scored in reference double order; 0 if exact */
} psa_pair_table;

typedef struct psa_context psa_context;

/* ---- library / context ------------------------------------------------------------------ */

int         psa_abi_version(void);
const char* psa_strerror(int status);

/* Number of visible CUDA devices of compute capability 10.x (0 on a CPU-only host). */
int psa_device_count(void);

/* Create a context over `ndevices` GPUs (device ordinals in devices[], or 0..ndevices-1 when
   devices == NULL).  Replaces MPI_Init + cudaMalloc-per-call (cuda_funcs.cu:43-69): buffers
   and streams persist across calls.  Fails with PSA_ERR_CUDA on a host without a usable GPU. */
int  psa_create(psa_context** ctx, const int* devices, int ndevices);
void psa_destroy(psa_context* ctx);
const char* psa_last_error(const psa_context* ctx);

/* Tuning knobs (for tests / benchmarks); every setting gives the same answers.  Unknown name -> PSA_ERR_ARG.
     "engine"        0 auto | 1 exact scalar kernel only | 2 bit-sliced scan
     "rank_planes"   -1 auto | 0,1,2,4 rank bit planes tracked by the scan (the rest is settled in-kernel)
     "scan_warps"    0 auto | 1..4 warps (x1024 offsets) per scan block in long mode
     "batch_mode"    -1 auto | 0 never | 1 whenever every query fits one window (len2 <= 1023)
     "stripe_mode"   -1 auto | 0 never | 1 whenever the batch qualifies: equal-length queries (len2 <= 1023), exact integer
                     order with small keys, all offsets, window within shared memory.  Then ONE kernel does everything:
                     each persistent block builds a striped bit-plane window of Seq1 in shared memory and its warp teams
                     scan and finish their queries against it (no k_profile / k_finish launch)
     "single_launch" 1 a single query in exact integer order runs as ONE cooperative launch (k_single: per-block striped windows,
                     grid barrier, combine, finish) | 0 the k_profile / k_scan (slices) / k_combine chain
     "slices"        0 auto | 1 never | n>=2 cut a single query into n ranges of alignment steps
     "sliced_keys"   1 bit-sliced epilogue when the weights allow it | 0 transpose + scalar keys
     "fused_finish"  1 the finish step runs inside the kernel before it where that saves a launch: in the scan block when
                     a query is a single tile (exact order), in k_combine's last block for a sliced query with few
                     combine blocks | 0 always its own kernel
     "derive_rank"   1 take the top-rank bit from the class planes when the table allows it | 0 always use a rank plane
     "pack_queries"  1 auto: equal-length queries that fit one window share blocks lane by lane when whole warps per
                     query would idle | 0 never | 2..8 force that many queries per block
     "zero_copy_results" 1 result sets up to 128 KB -- of any size in stripe mode, which stores a record as one 56-byte write --
                     are stored by the kernels straight into page-locked host memory (the caller's array if it is page-locked,
                     else the context's staging buffer) | 0 always copy back
     "stream_queries" 1 a one-shot stripe-mode call (psa_search_batch) copies its queries on a second stream in up to 8 pieces
                     while the kernel is already building its window; the kernel waits per task for the piece that holds the
                     task's queries | 0 one copy in front of the kernel
     "kernel_events" 1 psa_batch_run also brackets the dominant kernel with CUDA events ("main_kernel_ns")
     "gather_small"  1 a one-shot problem whose Seq1 and queries fit 32 KB (the gpu_run_program case) gathers both in a page-locked
                     buffer and sends them up as ONE copy | 0 one copy per sequence
     "min_split_work" a call is spread over at most (its pair evaluations / this value) of the context's GPUs, default 2.5e9
                     (~80 us of kernel per GPU: a shard pays its own copies, launch and wake-up whatever its size) | 0 always all
     "gate_timed_runs" 1 psa_batch_run enqueues its events and launches behind a gate (a stream wait on a page-locked word) that
                     the host opens once everything is enqueued, so the reported device time holds no host launch latency;
                     the first run of a prepared batch is never gated (lazy kernel loading) | 0 (default) no gate */
int psa_set_option(psa_context* ctx, const char* name, long long value);
/* Facts about the last run: "kernel_launches", "tiles", "candidate_tiles" (32-offset words re-scored in reference
   order), "main_kernel_ns", "engine", "rank_planes", "scan_warps", "batch_mode", "slices", "packed_queries", "packed_warps", "exact",
   "single_launch", "stripe_mode", "stripe_queries_per_task", "stripe_team_warps", "stripe_teams", "stripe_lanes", "devices_used" (GPUs the last call was spread over), "streamed_chunks" (pieces the
   last one-shot call streamed its queries in; 0 = one plain copy), and the host-side split of the
   last psa_search_batch in ns: "host_plan_ns", "host_prepare_ns", "host_enqueue_ns", "host_wait_ns", "host_total_ns".
   Unknown -> -1. */
long long psa_get_stat(const psa_context* ctx, const char* name);

/* ---- host-side table resolution (no GPU needed) ------------------------------------------ */

/* max_len2 sizes the exactness analysis (longest query in the batch). */
int psa_build_pair_table(const double weights[4], int is_max, long long max_len2, psa_pair_table* out);

/* ---- host-side partitioning (no GPU needed) ------------------------------------------------- */

/* One shard of a batch: queries [q_begin,q_end) over all their offsets (first = last = -1), or -- when
   the batch is a single query -- that query over absolute offsets [first,last). */
typedef struct psa_shard {
    int32_t q_begin, q_end;
    int64_t first, last;
} psa_shard;

/* The partition psa_search_batch applies over a context's GPUs, exposed for multi-process callers
   (one rank per GPU): contiguous query blocks balanced by pair evaluations, or for nq == 1 contiguous
   offset ranges in whole multiples of `granule` offsets (remainder to the last shard, like the
   reference's rank split cpu_funcs.c:128-133).  [first,last) restricts a single query (pass -1,-1 for
   all offsets).  Writes nshards entries; empty shards have q_begin == q_end. */
int psa_plan_shards(int64_t len1, const int64_t* q_off, int32_t nq, int nshards, int64_t granule,
                    int64_t first, int64_t last, psa_shard* out);

/* The launch shape of packed mode (equal-length queries that each fit one window share thread blocks lane by lane,
   DESIGN.md section 4): for `nq` queries of `len2` symbols against `len1`, how many queries one block takes
   (*queries_per_block, 0 = packed mode does not apply or would not pay) and how many warps it has (*warps).
   `force` = 0 picks the shape with the fewest idle lanes and requires >= 5 % over whole warps per query;
   2..8 forces that many queries per block when it fits.  Pure host arithmetic. */
int psa_plan_packing(int64_t len1, int64_t len2, int32_t nq, int force, int* queries_per_block, int* warps);

/* The launch shape of stripe mode (one launch per batch) for `nq` queries of `len2` symbols against
   `len1` on a GPU with `sm_count` SMs; rank_pass = number of rank bit planes the window carries (0 when the top rank follows
   from the sign classes, else 1 or 2).  shape[] = { applies (0/1), lanes per query S = ceil(offsets / 32), queries per task, warp
   passes per task, warps per team, teams per block, blocks, dynamic shared memory in bytes }.  Pure host arithmetic. */
int psa_plan_stripes(int64_t len1, int64_t len2, int32_t nq, int rank_pass, int sm_count, int shape[8]);

/* How a one-shot stripe-mode call streams `bytes` of queries to the device while its kernel is already running (option
   "stream_queries"): *pieces copies of *piece_bytes each (a multiple of 128; the last one shorter), every one followed by a
   flag the kernel waits for; *pieces == 0: the batch is too small to bother (one plain copy).  Pure host arithmetic. */
int psa_plan_stream_pieces(int64_t bytes, int* pieces, int64_t* piece_bytes);

/* Merge per-shard answers of ONE query given in ascending offset-range order, under the reference
   order: strictly better score wins, ties keep the earlier shard = lower offsets
   (MPI_MAXLOC/MINLOC on (score, rank), cpu_funcs.c:73-76; is_swapable cuda_funcs.cu:290-307). */
int psa_merge_results(int is_max, const psa_result* parts, int nparts, psa_result* out);

/* ---- the search ---------------------------------------------------------------------------- */

/* Batched search: nq queries against one Seq1 under one (weights, goal).
   Query q is seq2s[q_off[q] .. q_off[q+1]) (ASCII, not NUL-terminated); every query is searched
   over all offsets 0 .. len1-len2(q).  Work is partitioned over the context's GPUs by contiguous
   query blocks (or by offset range when nq == 1) and merged on the host under the reference
   order: best score, then lowest offset, then lowest char_offset (cuda_funcs.cu:290-307).
   Host pointers may be pageable or pinned. */
int psa_search_batch(psa_context* ctx, const double weights[4], int is_max,
                     const char* seq1, int64_t len1,
                     const char* seq2s, const int64_t* q_off, int32_t nq,
                     psa_result* out);

/* Many independent problems in one call, pipelined: the stacked blocks of an input file, the per-Seq1 batches of a query
   stream.  Every device slot of the context runs `lanes_per_device` lanes (1..8; 0 = the default, 2), each lane its own
   stream, buffers and host thread; the lanes take problems off a shared counter, so the copies, launch and wake-up of one
   problem overlap the kernel of another and the GPUs stay busy between problems (one psa_search_batch at a time leaves a
   GPU idle for the ~30 us either side of a 40 us kernel).  A problem is never split: with N GPUs, N x lanes problems are
   in flight.  problems[k].out receives problem k's records exactly as psa_search_batch would write them;
   problems[k].status its status.  Returns PSA_OK when every problem succeeded, else the status of the first that failed.
   The context's tuning knobs (psa_set_option) apply to every lane. */
typedef struct psa_problem {
    const double*  weights;    /* 4 weights */
    int32_t        is_max;
    int32_t        nq;
    const char*    seq1;
    int64_t        len1;
    const char*    seq2s;      /* queries, concatenated */
    const int64_t* q_off;      /* nq + 1 byte offsets into seq2s */
    psa_result*    out;        /* nq records */
    int32_t        status;     /* out */
    int32_t        reserved;
} psa_problem;
int psa_search_many(psa_context* ctx, psa_problem* problems, int32_t nproblems, int lanes_per_device);

/* One query restricted to absolute offsets [first,last) -- the gpu_run_program contract. */
int psa_search_range(psa_context* ctx, const double weights[4], int is_max,
                     const char* seq1, int64_t len1, const char* seq2, int64_t len2,
                     int64_t first, int64_t last, psa_result* out);

/* Reporting (the reference only has this as a debug printer, cpu_funcs.c:382-425): the whole score
   profile of ONE query over absolute offsets [first,last): scores[k] is the reference's score of offset first+k
   (find_best_mutant_offset, cpu_funcs.c:257-300), char_offsets[k] / letters[k] the single substitution it includes
   (-1 / '\0' where none exists).  char_offsets and letters may be NULL.  Runs on the context's first GPU. */
int psa_offset_scores(psa_context* ctx, const double weights[4], int is_max,
                      const char* seq1, int64_t len1, const char* seq2, int64_t len2,
                      int64_t first, int64_t last, double* scores, int32_t* char_offsets, char* letters);

/* Split-phase form of psa_search_batch, for callers that keep a batch resident in HBM:
   prepare = validate + H2D + table/profile resolution; run = the kernels only (returns the
   device time in ms measured with CUDA events on the library's streams, max over the
   context's GPUs); fetch = D2H of the per-query records + host scoring. */
int psa_batch_prepare(psa_context* ctx, const double weights[4], int is_max,
                      const char* seq1, int64_t len1,
                      const char* seq2s, const int64_t* q_off, int32_t nq);
int psa_batch_run(psa_context* ctx, float* device_ms);
int psa_batch_fetch(psa_context* ctx, psa_result* out);

/* Pinned host memory for callers that want true async H2D (bench e2e leg). */
void* psa_alloc_pinned(size_t bytes);
void  psa_free_pinned(void* p);

/* ---- drop-in for the reference entry point --------------------------------------------------
 * psa_gpu_run_program is the extern "C" spelling of
 *     double gpu_run_program(ProgramData* cpu_data, Mutant* returned_mutant,
 *                            int first_offset, int last_offset);           (cuda_funcs.h:33)
 * program_data points at the reference's ProgramData {int is_max; double weights[4];
 * char seq1[10001]; char seq2[5001];} (program_data.h:6-11), returned_mutant at its Mutant.
 * Returns the best score over absolute offsets [first,last), or -INFINITY/+INFINITY (goal
 * max/min) when no mutation exists (cuda_funcs.cu:143-145).  Like the reference it prints to
 * stderr and exit(EXIT_FAILURE)s on any CUDA error.  The library also exports the C++-mangled
 * symbol _Z15gpu_run_programP5_dataP7_mutantii that the reference's cpu_funcs.c links against.
 */
double psa_gpu_run_program(void* program_data, void* returned_mutant, int first_offset, int last_offset);

/* The six host primitives the reference's cpu_funcs.c imports from cuda_funcs.cu (cuda_funcs.h:44-61), C spellings.
   The library also exports them under the reference's C++-mangled names (include/cuda_funcs.h lists them), so the
   reference program links with no cuda_funcs.o at all.  Host only, no GPU needed; answers come from the pair-table
   resolver (psa_build_pair_table), never from the reference's hashtable_cpu.
     psa_get_hashtable_sign  cuda_funcs.cu:424-439  '*' ':' '.' '_' over [A-Z-], '\0' for anything else
     psa_get_pair_sign       cuda_funcs.cu:495-502  group membership only ('-' is in no group)
     psa_get_weight          cuda_funcs.cu:442-452  '*' -> +w[0], ':' -> -w[1], '.' -> -w[2], '_' -> -w[3], else 0
     psa_get_substitute      cuda_funcs.cu:310-421  best replacement for c2 facing c1 ('\0' outside [A-Z-])
     psa_is_swapable         cuda_funcs.cu:290-307  1 if (m2, score2) beats (m1, score1) under the reference order
     psa_strlen              cuda_funcs.cu:534-545  */
char   psa_get_hashtable_sign(char c1, char c2);
char   psa_get_pair_sign(char a, char b);
double psa_get_weight(char sign, const double* w);
char   psa_get_substitute(char c1, char c2, const double* w, int is_max);
int    psa_is_swapable(const psa_mutant* m1, const psa_mutant* m2, double score1, double score2, int is_max);
int    psa_strlen(const char* str);

/* input.txt / output.txt in the reference's format (cpu_funcs.c:353-378): four weights, Seq1,
   Seq2, "maximum"|"minimum" (anything else = minimum), whitespace separated; output
   "<mutant>\n<offset> <score %g>" without trailing newline.  seq buffers are malloc()ed. */
int psa_read_input_file(const char* path, double weights[4], int* is_max, char** seq1, char** seq2);
int psa_write_output_file(const char* path, const char* mutant, int offset, double score);

/* The whole reference program for one input file: read, search on the context's GPUs, write. */
int psa_run_files(psa_context* ctx, const char* input_path, const char* output_path, psa_result* out);

/* Extension: the reference's own input.txt stacks 15 problem blocks but reads only the first
   (cpu_funcs.c:353-368 parses one group of 4 weights + Seq1 + Seq2 + goal and ignores the rest).  This entry
   point consumes EVERY complete block of the file and writes one output stanza per block, stanzas separated
   by a newline, each exactly what the reference writes for that block alone.  *nblocks receives the count. */
int psa_run_files_all(psa_context* ctx, const char* input_path, const char* output_path, int* nblocks);

/* ---- widening ------------------------------------------------------------------------------------ */

/* Query lists (8f-2).  The reference's input.txt carries exactly one Seq2 (cpu_funcs.c:360); the batched API wants many.
   Two on-disk forms are read: FASTA (any line starting with '>' is a header; the lines up to the next header, white space
   removed, are one query) and plain lists (no '>' anywhere: every white-space separated token is one query).  *seq2s
   (concatenated queries, not NUL separated) and *q_off (nq + 1 byte offsets) are malloc()ed: release with psa_free. */
int  psa_read_query_file(const char* path, char** seq2s, int64_t** q_off, int32_t* nq);
void psa_free(void* p);

/* input.txt supplies weights, Seq1 and the goal (its own Seq2 is ignored), `queries_path` the queries; one output stanza
   per query, exactly what the reference writes for that query alone ("<mutant>\n<offset> <score %g>"), stanzas separated by a
   newline.  The mutant strings are produced on the device (psa_search_batch_mutants). */
int psa_run_query_file(psa_context* ctx, const char* input_path, const char* queries_path, const char* output_path, int32_t* nq);

/* Reporting on the device (8f-3; in the reference only the winner's string is built, on the host: cpu_funcs.c:96-98).
   psa_search_batch_mutants = psa_search_batch + the mutated copy of every query (Seq2 with its one substitution applied; a
   query with no possible mutation is copied unchanged), written by a kernel in the layout of seq2s (out_mutants has
   q_off[nq] bytes, query q at q_off[q]). */
int psa_search_batch_mutants(psa_context* ctx, const double weights[4], int is_max,
                             const char* seq1, int64_t len1, const char* seq2s, const int64_t* q_off, int32_t nq,
                             psa_result* out, char* out_mutants);

/* The k best offsets of ONE query over absolute offsets [first,last) under the reference order -- best score first, equal
   scores by ascending offset (is_swapable, cuda_funcs.cu:290-307) -- selected on the device from the per-offset scores of
   find_best_mutant_offset (cpu_funcs.c:257-300).  offsets[r], scores[r] and (if not NULL) char_offsets[r], letters[r]
   describe rank r; *found = entries written (< k when fewer offsets have a mutation).  offsets[0] is the psa_search_range
   answer.  Runs on the context's first GPU. */
int psa_topk_offsets(psa_context* ctx, const double weights[4], int is_max,
                     const char* seq1, int64_t len1, const char* seq2, int64_t len2, int64_t first, int64_t last,
                     int32_t k, int32_t* offsets, double* scores, int32_t* char_offsets, char* letters, int32_t* found);

#ifdef __cplusplus
}
#endif
#endif /* PSA_B200_H */
