// psa_kernels.cuh -- device-side contract between the engine and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include "psa_common.h"

namespace psa {

// Device pointers of one batch on one GPU.
struct BatchPtrs {
    const uint8_t* seq1;        // ASCII, len1 bytes (+ padding readable up to seq1_alloc)
    const uint8_t* seq2s;       // ASCII, all queries concatenated
    const int64_t* qoff;        // nq+1 byte offsets into seq2s
    const int32_t* tile_start;  // nq+1 prefix sum of tiles per query
    TileRec*       tiles;       // total_tiles records
    QueryRec*      out;         // nq records
    int64_t*       lane_keys;   // scan engine, re-score mode: per (tile, 32-offset word) upper estimate of its keys
    uint8_t*       code_table;  // [27][32] copy of DeviceTable::code in global memory (uploaded when the table changes)
    uint2*         partial;     // slice mode: [slice][offset] partial counts {N(b0) | N(b1) << 16, N(b0&b1) | rank bits << 16}
    int64_t        partial_stride;   // offsets per slice row of `partial`
    int32_t*       cand_count;  // [0] = number of 32-offset words re-scored in reference order (statistic; zeroed by the first kernel)
    int32_t*       err_flag;    // mapped host word: set to run_tag by any thread that meets a symbol outside [A-Z-]
    int32_t        run_tag;     //   (a fresh tag per run: the word never needs clearing and is read without a copy)
    // bit-plane profile of Seq1 (scan engine): [row][word] of 64-bit (class planes) and
    // [row][word][rank_planes] of 32-bit words; rows = 28 (27 symbols + zero row)
    uint2*         cls_planes;
    uint32_t*      rank_planes;
    int64_t        plane_words; // words per row
    const DeviceTable* table;   // the whole resolved table in global memory (uploaded when it changes): kernels that copy it to shared
                                //   memory take this pointer instead of a 3.7 KB by-value parameter
    int32_t*       sync;        // k_single: [0] grid barrier, [1] "last block" ticket; zero between launches (the last block resets them)
    // Streamed batches (one-shot stripe mode): the queries arrive on a second stream in `ready_chunks` pieces of
    // `ready_chunk_bytes` while the kernel already runs; ready[c] == ready_tag once every byte of piece c (and of the pieces
    // before it) has landed.  ready == nullptr: the batch is resident.
    const int32_t* ready;
    int32_t        ready_tag;
    int32_t        ready_chunks;
    int64_t        ready_chunk_bytes;   // multiple of 128
};

#if defined(__CUDACC__)
__device__ __forceinline__ void report_bad_symbol(const BatchPtrs& P)
{
    *reinterpret_cast<volatile int32_t*>(P.err_flag) = P.run_tag;
}

// Streamed batches: wait (whole warp) until the query bytes [.., byte_end) have landed.  Lane 0 polls the piece's flag with an
// acquire load at system scope (the writer is the copy stream: a DMA followed by a stream memory operation); the warp
// barrier passes the ordering on to the other lanes.  The piece waited for is the one that holds the END of the last 128-byte
// line touched, so no line can enter L1 before all of its bytes are final.  Bounded: should the flag not come within
// kStreamWaitNs the warp reports it (err_flag[1]) and returns false -- the caller leaves the kernel, the host fails the call.
constexpr unsigned long long kStreamWaitNs = 4000000000ull;
__device__ __forceinline__ bool stream_wait(const BatchPtrs& P, int64_t byte_end)
{
    if (P.ready == nullptr) return true;
    const int64_t line_last = ((byte_end + 127) & ~int64_t(127)) - 1;
    int64_t c = line_last / P.ready_chunk_bytes;
    if (c >= P.ready_chunks) c = P.ready_chunks - 1;
    int ok = 1;
    if ((threadIdx.x & 31) == 0) {
        const int32_t* flag = P.ready + c;
        unsigned long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            int32_t v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v == P.ready_tag) break;
            __nanosleep(64);
            if ((spins & 1023u) == 1023u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > kStreamWaitNs) { ok = 0; break; }
            }
        }
        if (!ok) reinterpret_cast<volatile int32_t*>(P.err_flag)[1] = P.run_tag;
    }
    ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
    return ok != 0;
}

// A 56-byte result record held by lane 0 leaves the warp as ONE store instruction (lanes 0..6, eight bytes each): one
// contiguous write instead of seven -- what matters when `dst` is page-locked host memory and every store is a bus
// transaction of its own.  All 32 lanes call this.
__device__ __forceinline__ void warp_store_record(QueryRec* dst, const QueryRec& rec)
{
    static_assert(sizeof(QueryRec) == 56, "seven 8-byte words");
    const int lane = threadIdx.x & 31;
    unsigned long long w[7];
    w[0] = (unsigned long long)(uint32_t)rec.offset | ((unsigned long long)(uint32_t)rec.char_offset << 32);
    w[1] = (unsigned long long)(uint32_t)rec.ch | ((unsigned long long)(uint32_t)rec.rank << 32);
    w[2] = (unsigned long long)__double_as_longlong(rec.score);
    w[3] = (unsigned long long)rec.counts[0];
    w[4] = (unsigned long long)rec.counts[1];
    w[5] = (unsigned long long)rec.counts[2];
    w[6] = (unsigned long long)rec.counts[3];
    unsigned long long mine = 0ull;
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const unsigned long long x = __shfl_sync(0xFFFFFFFFu, w[k], 0);
        if (lane == k) mine = x;
    }
    if (lane < 7) reinterpret_cast<unsigned long long*>(dst)[lane] = mine;
}
#endif

// Offsets of a query are tiled from `base` = first rounded down to a multiple of 128 so that bit-plane
// rows line up on 16 bytes (TMA bulk copies); a tile covers [base + t*tile, base + (t+1)*tile) intersected
// with [first, last).
__host__ __device__ inline int64_t tile_base(int64_t first) { return first & ~int64_t(127); }

// Slice mode (one query that cannot fill the GPU): the alignment steps are cut into `slices` ranges of
// `slice_len` steps, each scanned by its own blocks; k_combine adds the partial counts per offset.
struct SliceGeom {
    int slices = 1;        // 1 = off
    int slice_len = 0;     // steps per slice (multiple of 128)
    int scan_tile = 0;     // offsets per scan block (warps x 1024)
    int scan_tiles = 0;    // scan blocks per slice
    bool fused_finish = false;  // long mode, exact order, one tile per query: the scan block also finishes its query
    bool allow_derive = true;   // option "derive_rank": take the top-rank bit from the class planes when the table allows it
    bool fused_combine = false; // slice mode on a small grid: k_combine's last block runs the finish step (no k_finish launch)
    int pack_q = 0;             // packed mode (k_scan_packed): queries per block, 0 = off
    int pack_warps = 0;         //   and warps per block = ceil(pack_q * lanes per query / 32)
};
constexpr int kCombineTile = 256;   // offsets per tile record in slice mode
constexpr size_t kZeroCopyMaxBytes = 128 * 1024;   // result sets up to this size are written straight into host memory
constexpr int kPackMaxQ = 8;        // packed mode: queries per block
constexpr int kPackMaxWarps = 8;    //   and warps per block
// packed mode: does one window hold every offset of a (len1, len2) query, inside the plane buffer?
bool scan_packed_fits(int64_t len1, int64_t len2);

// Stripe mode (psa_stripe.cu): equal-length queries against a striped window that one persistent block per SM builds
// in shared memory -- the whole batch in one launch.
struct StripeGeom {
    int ok = 0;          // 0: stripe mode does not apply to this batch
    int S = 0;           // lanes per query = ceil(offsets / 32); bit t of lane l is offset l + t * S
    int steps = 0;       // len2 rounded up to a multiple of 32
    int Wn = 0;          // words per window row = S + steps
    int Q = 0;           // queries per task
    int passes = 0;      // warp passes per task = ceil(Q * S / 32)
    int T = 0;           // warps per team (a team owns a task)
    int teams = 0;       // teams per block
    int ntasks = 0;
    int ro_stride = 0;   // words between the row-offset vectors of two queries of a task
    int blocks = 0;
    size_t smem = 0;
    int threads = 0;     // block size
    int rank_planes_read = 0;   // rank bit planes in the window: 0 (none tracked, or the top rank is derived), 1 (uint32), 2 (uint2)
    int split = 0;       // passes of a task that are cut in two along their steps (then T == passes + split: one unit per warp), so that
                         //   the four schedulers of an SM carry equal shares of a task whose pass count is not a multiple of four
};
// threads of the one block per SM: 20 warps at <= 96 registers (28 warps at 72 registers measured no faster on short queries)
__host__ __device__ constexpr int stripe_threads(int nb) { return nb <= 7 ? 640 : 640; }
inline int stripe_threads_for_len2(int64_t len2) { return stripe_threads(len2 <= 127 ? 7 : 10); }
// two rank planes (uint2 entries) sit at this FIXED distance behind the class window, so that one address register serves both
// loads of a step (LDS [addr] and LDS [addr + kStripeRankBase]); the class window must fit below it: 28 x Wn x 8 <= 96 KB
constexpr int kStripeRankBase = 96 * 1024;
constexpr int kStripeMaxQ = 32;                     // queries per task
constexpr int kStripeMaxPasses = 64;                // passes per task
constexpr size_t kStripeSmemMax = 224 * 1024;        // of the 227 KB a block may have (a little static shared memory on top)
StripeGeom stripe_plan(int64_t len1, int64_t len2, int32_t nq, int rank_planes_read, int sm_count, bool force = false);
// does the scan take the top-rank bit from the class counts (no rank plane read) for this table / plane count?
bool stripe_derives_rank(const DeviceTable& T, int rank_planes, bool allow_derive);
bool stripe_keys_ok(const DeviceTable& T, int64_t len2);     // the bit-sliced key epilogue applies (required by stripe mode)
void launch_stripe(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int rank_planes, bool allow_derive,
                   const StripeGeom& SG, cudaStream_t stream);
// bit-sliced key epilogue: plane count for this table and query length (0 = keys too wide), and the bias that makes keys non-negative
int sliced_key_planes(const DeviceTable& T, int64_t max_len2, int nb, int64_t* bias_out);

// One query in one launch (psa_single.cu): (1024-offset tile) x (step slice) units build their own striped windows, a grid
// barrier, then combine + finish in the same kernel.
struct SingleGeom {
    int ok = 0;
    int slice_steps = 0;   // alignment steps per slice (multiple of 32, <= 992)
    int slices = 0;
    int tiles = 0;         // 1024-offset tiles from tile_base(first)
    int units = 0;         // tiles x slices
    int Wn = 0;            // window words per plane row = 32 + slice_steps
    int span = 0;          // Seq1 symbols a unit can touch
    int blocks = 0;
    size_t smem = 0;
};
SingleGeom single_plan(int64_t len1, int64_t len2, int64_t first, int64_t last, int sm_count);
void launch_single(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int rank_planes, const SingleGeom& SG, cudaStream_t stream);

// ---- launchers (all asynchronous on `stream`) -------------------------------------------------
// exact scalar kernel over every tile (engine 1)
void launch_exact_tiles(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, cudaStream_t stream);
// per-query finish: winner over the tile records (re-scored in reference order when the weights are not
// exactly summable and the records come from the scan), then char_offset + counts + substitute letter
void launch_finish(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, bool scan_records, cudaStream_t stream);
// per-offset scores / mutated position / letter of one query over [G.first, G.last)
void launch_offset_profile(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, double* scores, int32_t* char_offsets,
                           uint8_t* letters, cudaStream_t stream);
// top-k offsets of one query from its per-offset profile (one block, k selection passes), and the mutated copies of a batch's queries
struct TopkRec {
    double  score;
    int32_t offset;
    int32_t char_offset;
    int32_t letter;
    int32_t pad;
};
void launch_topk(int is_max, int64_t first, int64_t n, int k, const double* scores, const int32_t* char_offsets, const uint8_t* letters,
                 TopkRec* out, int32_t* found, cudaStream_t stream);
void launch_emit_mutants(const BatchGeom& G, const BatchPtrs& P, const QueryRec* recs, int64_t nbytes, uint8_t* out, int sm_count,
                         cudaStream_t stream);
// bit-plane profile of Seq1 for the scan engine
void launch_profile(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int rank_planes, int sm_count,
                    cudaStream_t stream);
// bit-sliced scan of every tile
void launch_scan(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int rank_planes, int64_t max_len2,
                 bool batch, bool sliced_keys_ok, int sm_count, const SliceGeom& slices, cudaStream_t stream);
// batch mode (window shared by many queries, tile = 1024) applies when every query fits one window
bool scan_batch_mode(const BatchGeom& G, int64_t max_len2, int sm_count);

// Launch `kernel` so that it may overlap the tail of the previous kernel in `stream` (programmatic dependent
// launch); the kernel itself orders its reads with pdl_wait().
template <class... KArgs, class... Args>
inline void launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// scan engine limits
constexpr int kDefaultEngine = 2;                   // engine picked by "auto" (1 scalar, 2 bit-sliced scan)
constexpr int kScanWarps = 4;                       // max warps per block; a warp owns 1024 offsets
constexpr int kScanTile = kScanWarps * 1024;        // largest tile (the engine picks 1..4 warps per batch)
constexpr int kScanMaxLen2 = 32767;                 // 15 counter planes
constexpr int kExactTile = 256;                     // tile of the scalar engine when used alone
constexpr int64_t kExactMaxLen2 = (1 << 20) - 1;    // 20-bit count fields

int scan_chunk_steps(int rank_planes, int64_t max_len2);   // i-steps staged per shared-memory window (multiple of 128)
int64_t scan_plane_words(int64_t len1);                    // words per bit-plane row incl. zero padding (multiple of 4)
size_t scan_smem_bytes(int rank_planes, int chunk, int warps);

} // namespace psa
