// psa_engine.cu -- context, persistent buffers, (query | offset-range) partitioning over GPUs,
// and the C ABI declared in include/psa_b200.h.
//
// Replaces the host side of the reference's GPU path and the orchestration around it:
//   gpu_run_program            cuda_funcs.cu:6-146   (cudaMalloc x3 + copies + cudaFree x3 per call)
//   divide_execute_tasks       cpu_funcs.c:123-218   (rank x CPU/GPU x thread split of offsets)
//   initiate_program's reduce  cpu_funcs.c:64-94     (MPI_Allreduce MAXLOC/MINLOC + Send/Recv)
// with one process driving 1..8 GPUs: buffers and streams persist in a context, every GPU gets a
// contiguous block of queries (or, for a single query, a contiguous range of offsets -- the same
// split as cpu_funcs.c:128-133 with GPUs in place of ranks), all GPUs are enqueued before any is
// waited for, and the <= 8 candidates per query are merged on the host under the reference order
// (best score, then lowest offset: MAXLOC/MINLOC ties go to the lowest rank = lowest offsets).
#include "psa_host.h"
#include "psa_kernels.cuh"

#include <cuda_runtime.h>
#include <cuda.h>            // types and the prototype of cuStreamWriteValue32 only: the entry point is looked up at run time

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

using namespace psa;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct DeviceState {
    int dev = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // streamed batches: the queries' host-to-device copies (the kernel is already running)
    DevBuf ready;                            //   flags the copy stream writes behind each piece (BatchPtrs::ready)
    int32_t* h_tags = nullptr;               //   page-locked source of those flag values when stream memory operations are not available
    int32_t ready_tag = 0;
    int32_t gate_tag = 0;                    // timed runs: value of h_err[2] that opens the current run's gate
    int runs_of_batch = 0;                   //   runs since the batch was prepared / an option changed (the first one is never gated)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;   // whole run / dominant kernel
    DevBuf seq1, seq2s, qoff, tile_start, tiles, out, lane_keys, cls_planes, rank_planes, partial, code_table, mutants, sync, table;
    SliceGeom SG{};
    StripeGeom stripe{};       // ok != 0: this shard runs in stripe mode (one launch: window + scan + finish)
    SingleGeom single{};       // ok != 0: this shard (one query, exact order) runs as one cooperative launch (k_single)
    int64_t plan_key[5] = { -1, -1, -1, -1, -1 };   // (len1, len2, nq, rank planes asked, forced) of the cached stripe plan:
    StripeGeom plan_cached{};                        //   the planner's search over (Q, T) is not repeated for a repeated batch shape
    PinBuf h_qoff, h_tile_start, h_out;
    DevBuf inbuf;              // small one-shot problems: Seq1 and the queries in ONE device buffer, filled by ONE copy from h_in
    PinBuf h_in;
    // slice of the current batch owned by this GPU
    int q_begin = 0, q_end = 0;
    BatchGeom G{};
    BatchPtrs P{};
    bool active = false;
    long long st_launches = 0, st_tiles = 0, st_main_ns = 0;   // counters of the last run on this GPU
    long long st_chunks = 0;                                   // pieces the last one-shot call streamed its queries in (0: one plain copy)
    long long st_prepare_ns = 0, st_enqueue_ns = 0, st_wait_ns = 0;   // host time of the last one-shot search on this GPU's thread
    int32_t* h_err = nullptr;  // mapped page-locked word the kernels write the run's tag into on a bad symbol (no copy, no memset)
    int32_t run_tag = 0;
    long long table_epoch = -1;   // ctx->table_epoch whose pair table is in code_table
    bool zc_out = false;       // this run's records are written over the bus by the kernels themselves (no device-to-host copy)
    bool zc_direct = false;    //   ... into the caller's page-locked array (else into h_out)
    bool streamed = false;     // this run's queries arrive on copy_stream while the kernel runs
    float run_ms = 0.f;
};

// One host thread per additional GPU: enqueueing copies and launches for 8 devices one after the other from a
// single thread costs more than the work itself on small problems (every device would wait for its turn).
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool pending = false, quit = false;
    int rc = 0;
};

} // namespace

struct psa_context {
    std::vector<DeviceState> devs;
    std::vector<std::unique_ptr<Worker>> workers;     // workers[g-1] serves devs[g]
    std::vector<psa_shard> plan;                      // shard of the current batch per GPU
    std::vector<psa_context*> lanes;                  // psa_search_many: one single-device child context per (device slot, lane)
    std::vector<std::unique_ptr<Worker>> lane_workers; //   and a host thread for every lane but the first (which runs on the caller's)
    std::mutex err_mu;
    std::string err;
    // options
    int opt_engine = 0;        // 0 auto, 1 exact scalar, 2 bit-sliced scan
    int opt_rank_planes = -1;  // -1 auto
    int opt_scan_warps = 0;    // 0 auto, 1..4
    int opt_fused_finish = 1;  // 0: always run k_finish as its own kernel
    int opt_derive_rank = 1;   // 0: always read the top-rank bit from a rank plane
    int opt_pack_queries = 1;  // 0 never pack | 1 auto | 2..8 force that many queries per block (tests)
    int opt_zero_copy = 1;     // 1: result sets are written by the kernels straight into page-locked host memory (small ones; any size in stripe mode)
    long long opt_min_split_work = 2500000000ll;   // a call is spread over at most work / this many GPUs (pair evaluations; 0: always over all)
    int opt_gate_timed_runs = 0; // 1: psa_batch_run enqueues its events and launches behind a host-released gate (device time only in the bracket)
    int opt_gather_small = 1;   // 1: a small one-shot problem's Seq1 and queries are gathered on the host and go up as one copy
    int opt_stream_queries = 1; // 1: one-shot stripe-mode batches copy their queries on a second stream while the kernel builds its window
    long long table_epoch = 0; // bumped whenever `table` is rebuilt
    bool one_shot = false;     // the batch being prepared belongs to a prepare + run + fetch call (psa_search_batch / _range)
    int opt_kernel_events = 0; // 1: psa_batch_run also brackets the dominant kernel with events (stat main_kernel_ns)
    int opt_slices = 0;        // 0 auto, 1 never cut a query along its alignment steps, n >= 2: ask for n slices
    int opt_sliced_keys = 1;   // 1: bit-sliced epilogue when the keys allow it, 0: always transpose + scalar keys
    int opt_batch_mode = -1;   // -1 auto, 0 never, 1 whenever the queries fit one window
    int opt_stripe_mode = -1;  // -1 auto, 0 never, 1 whenever the batch qualifies (equal lengths, exact order, window fits)
    int opt_single_launch = 1; // 1: a single query in exact order runs as one cooperative launch (k_single), 0: the k_profile / k_scan / k_combine chain
    // current batch
    bool prepared = false, ran = false;
    bool range_split = false;  // single query split by offset range over the GPUs
    DeviceTable table{};
    bool table_valid = false;
    int64_t table_len2 = -1;
    double weights[4] = { 0, 0, 0, 0 };
    int is_max = 0;
    int nq = 0;
    int engine = 1;
    int rank_planes = 0;
    int scan_tile = kScanTile;   // offsets per scan tile = 1024 x warps per block
    bool batch_mode = false;     // scan engine: queries share staged windows (k_scan_batch)
    int64_t max_len2 = 0;
    int64_t uniform_len2 = 0;    // > 0: all queries of the batch have this length
    long long st_plan_ns = 0, st_total_ns = 0;   // host time of the last one-shot search: planning / the whole call
    const char* in_seq1 = nullptr;       // the caller's buffers of the batch being prepared (read by the per-GPU copies)
    const char* in_seq2s = nullptr;
    const int64_t* in_qoff = nullptr;
    int64_t in_len1 = 0;
};

namespace {

int fail(psa_context* ctx, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) {
        std::lock_guard<std::mutex> lk(ctx->err_mu);
        ctx->err = buf;
    }
    return code;
}

#define PSA_CUDA(ctx, call)                                                                              \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, PSA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                       \
    } while (0)

int ensure_dev(psa_context* ctx, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return PSA_OK;
    if (b.p) PSA_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    PSA_CUDA(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return PSA_OK;
}

int ensure_pin(psa_context* ctx, PinBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return PSA_OK;
    if (b.p) PSA_CUDA(ctx, cudaFreeHost(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    PSA_CUDA(ctx, cudaMallocHost(&b.p, want));
    b.cap = want;
    return PSA_OK;
}

void release(DeviceState& d)
{
    cudaSetDevice(d.dev);
    for (DevBuf* b : { &d.seq1, &d.seq2s, &d.qoff, &d.tile_start, &d.tiles, &d.out, &d.lane_keys, &d.partial, &d.code_table,
                       &d.cls_planes, &d.rank_planes, &d.mutants, &d.sync, &d.table, &d.ready, &d.inbuf })
        if (b->p) cudaFree(b->p);
    for (PinBuf* b : { &d.h_qoff, &d.h_tile_start, &d.h_out, &d.h_in })
        if (b->p) cudaFreeHost(b->p);
    if (d.h_err) cudaFreeHost(d.h_err);
    if (d.h_tags) cudaFreeHost(d.h_tags);
    if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
    if (d.ev0) cudaEventDestroy(d.ev0);
    if (d.ev1) cudaEventDestroy(d.ev1);
    if (d.evk0) cudaEventDestroy(d.evk0);
    if (d.evk1) cudaEventDestroy(d.evk1);
    if (d.stream) cudaStreamDestroy(d.stream);
}

inline int64_t offsets_of(int64_t len1, int64_t len2) { return len1 - len2 + 1; }

int pick_rank_planes(const psa_context* ctx)
{
    int avail = ctx->table.nranks - (ctx->table.has_none ? 0 : 1);   // planes needed to resolve everything
    if (avail < 0) avail = 0;
    // offsets the planes leave unresolved are settled inside the scan kernel, so the plane count is a pure
    // speed knob: long queries saturate the top rank within a few dozen steps, short ones benefit from two
    // measured (tools/stats.py): one plane is best at every length once unresolved offsets are settled in-kernel
    int want = ctx->opt_rank_planes >= 0 ? ctx->opt_rank_planes : 1;
    int k = std::min(want, std::max(avail, 0));
    // the kernels are instantiated for 0, 1, 2 and 4 planes and the plane buffer is sized from this number: 3 (e.g. four
    // planes asked for, three ranks to resolve) becomes 4 -- a plane for a rank that does not exist is simply all zero
    if (k == 3 || k > 4) k = 4;
    return k;
}

// Run fn(device) for every GPU of the context: GPU 0 on the calling thread, the others on their worker threads.
// Devices whose shard of the current plan is empty are not woken (waking a thread to find nothing costs a few microseconds).
template <class F>
int for_each_device(psa_context* ctx, F fn)
{
    const int ndev = (int)ctx->devs.size();
    auto idle = [ctx](int g) {
        return g > 0 && (size_t)g < ctx->plan.size() && ctx->plan[g].q_begin == ctx->plan[g].q_end;
    };
    for (int g = 1; g < ndev; g++) {
        if (idle(g)) { ctx->devs[g].active = false; continue; }
        Worker& w = *ctx->workers[g - 1];
        std::lock_guard<std::mutex> lk(w.mu);
        w.job = [ctx, g, &fn]() { return fn(ctx->devs[g]); };
        w.pending = true;
        w.cv.notify_one();
    }
    int rc = fn(ctx->devs[0]);
    for (int g = 1; g < ndev; g++) {
        if (idle(g)) continue;
        Worker& w = *ctx->workers[g - 1];
        std::unique_lock<std::mutex> lk(w.mu);
        w.cv.wait(lk, [&w] { return !w.pending; });
        if (!rc) rc = w.rc;
    }
    return rc;
}

void worker_loop(Worker* w)
{
    std::unique_lock<std::mutex> lk(w->mu);
    for (;;) {
        w->cv.wait(lk, [w] { return w->pending || w->quit; });
        if (w->quit) return;
        lk.unlock();
        const int rc = w->job();
        lk.lock();
        w->rc = rc;
        w->pending = false;
        w->cv.notify_all();
    }
}

// ---- streamed batches -------------------------------------------------------------------------------------------
// A one-shot stripe-mode call does not wait for its queries before it launches: Seq1 (a few KB) goes ahead on the main
// stream, the queries follow on `copy_stream` in up to kStreamMaxChunks pieces, each with a flag written behind it by a
// stream memory operation (cuStreamWriteValue32; a 4-byte copy from page-locked memory where that is not available), and
// k_stripe -- launched right after the copies are enqueued -- builds its window while they are in flight and waits per task
// for the piece that holds the task's queries (stream_wait, psa_kernels.cuh).  Everything the kernel waits for is enqueued
// BEFORE the kernel is, so a failing copy can never leave a launched kernel waiting.
constexpr int64_t kGatherMaxBytes = 32 * 1024;      // Seq1 + padding + queries up to this size go up as one copy (one-shot calls)
constexpr int kMaxLanes = 8;                       // psa_search_many: lanes per device slot
constexpr int kStreamMaxChunks = 8;
constexpr int64_t kStreamChunkBytes = 512 * 1024;
constexpr int64_t kStreamMinBytes = 64 * 1024;

using WriteValue32Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static WriteValue32Fn stream_memop(const char* symbol)
{
    if (std::getenv("PSA_NO_STREAM_MEMOPS")) return nullptr;                // tests: force the copy-based flags / ungated runs
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st{};
    if (cudaGetDriverEntryPoint(symbol, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return reinterpret_cast<WriteValue32Fn>(p);
}
WriteValue32Fn stream_write_value32()
{
    static const WriteValue32Fn fn = stream_memop("cuStreamWriteValue32");
    return fn;
}
WriteValue32Fn stream_wait_value32()                                         // same signature
{
    static const WriteValue32Fn fn = stream_memop("cuStreamWaitValue32");
    return fn;
}

// Timed runs (psa_batch_run) are enqueued behind a gate: the stream first waits for a page-locked word that the host writes
// AFTER it has enqueued event + launches + event, so the event bracket holds device time only.  Without it the GPU is idle
// when the first event is recorded and the host's launch latency (~3 us) sits inside the bracket
// (tools/probes/launch_probe.cu: 8.0 -> 5.0 us around an empty launch).  The gate opens when this object goes out of scope,
// on every path.  A kernel's FIRST launch in a process must not sit behind a closed gate: with lazy module loading (the
// default) the driver loads the function at that launch and may wait for the context to drain -- which the gated stream
// never does.  So only runs after the first run of a prepared batch are gated (same data, same kernels), any option change
// makes the next run a first run again, and the option is off unless the caller (bench.py) turns it on.
struct TimedGate {
    volatile int32_t* word = nullptr;
    int32_t tag = 0;
    void close(DeviceState& d, bool enabled)
    {
        // Never under a tool that is injected into the CUDA calls (Nsight Compute profiles a kernel INSIDE its launch call,
        // i.e. before the host gets to open the gate: the profiled kernel would wait behind it for ever).
        static bool ok = std::getenv("CUDA_INJECTION64_PATH") == nullptr && std::getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") == nullptr;
        const WriteValue32Fn wait = enabled && ok ? stream_wait_value32() : nullptr;   // (ok is also cleared the first time the driver refuses)
        if (!wait) return;
        if (d.gate_tag >= 0x7FFFFFF0) d.gate_tag = 0;
        const int32_t t = ++d.gate_tag;
        void* dv = nullptr;
        if (cudaHostGetDevicePointer(&dv, d.h_err + 2, 0) != cudaSuccess || !dv) { cudaGetLastError(); ok = false; return; }
        if (wait((CUstream)d.stream, (CUdeviceptr)dv, (cuuint32_t)t, CU_STREAM_WAIT_VALUE_EQ) != CUDA_SUCCESS) { ok = false; return; }
        word = d.h_err + 2;
        tag = t;
    }
    ~TimedGate() { if (word) *word = tag; }
};

// copies of the queries in pieces on the copy stream + their flags; fills the ready fields of d.P
int enqueue_streamed_queries(psa_context* ctx, DeviceState& d, const char* src, int64_t bytes)
{
    int rc;
    if (!d.copy_stream) PSA_CUDA(ctx, cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    if (!d.ready.p) {
        if ((rc = ensure_dev(ctx, d.ready, sizeof(int32_t) * kStreamMaxChunks))) return rc;
        PSA_CUDA(ctx, cudaMemset(d.ready.p, 0, sizeof(int32_t) * kStreamMaxChunks));
        PSA_CUDA(ctx, cudaHostAlloc((void**)&d.h_tags, sizeof(int32_t) * kStreamMaxChunks, cudaHostAllocDefault));
    }
    if (d.ready_tag >= 0x7FFFFFF0) {                                        // no run is in flight here
        PSA_CUDA(ctx, cudaMemset(d.ready.p, 0, sizeof(int32_t) * kStreamMaxChunks));
        d.ready_tag = 0;
    }
    const int32_t tag = ++d.ready_tag;
    int pieces = 0;
    int64_t cb = 0;
    psa_plan_stream_pieces(bytes, &pieces, &cb);
    const int64_t chunks = pieces;
    static bool memops_ok = true;                                            // cleared for good the first time the driver refuses one
    const WriteValue32Fn wv = memops_ok ? stream_write_value32() : nullptr;
    for (int64_t c = 0; c < chunks; c++) {
        const int64_t b0 = c * cb, n = std::min(cb, bytes - b0);
        PSA_CUDA(ctx, cudaMemcpyAsync((char*)d.seq2s.p + b0, src + b0, (size_t)n, cudaMemcpyHostToDevice, d.copy_stream));
        bool flagged = false;
        if (wv && memops_ok) {
            flagged = wv((CUstream)d.copy_stream, (CUdeviceptr)((int32_t*)d.ready.p + c), (cuuint32_t)tag, 0u) == CUDA_SUCCESS;
            if (!flagged) memops_ok = false;
        }
        if (!flagged) {
            d.h_tags[c] = tag;
            PSA_CUDA(ctx, cudaMemcpyAsync((int32_t*)d.ready.p + c, d.h_tags + c, sizeof(int32_t), cudaMemcpyHostToDevice, d.copy_stream));
        }
    }
    d.P.ready = (const int32_t*)d.ready.p;
    d.P.ready_tag = tag;
    d.P.ready_chunks = (int32_t)chunks;
    d.P.ready_chunk_bytes = cb;
    d.streamed = true;
    d.st_chunks = chunks;
    return PSA_OK;
}

// Enqueue H2D copies and geometry for one GPU's slice. first/last >= 0 selects range mode (nq == 1).
int prepare_device(psa_context* ctx, DeviceState& d, const char* seq1, int64_t len1, const char* seq2s,
                   const int64_t* q_off, int q_begin, int q_end, int64_t first, int64_t last)
{
    PSA_CUDA(ctx, cudaSetDevice(d.dev));
    const int nq = q_end - q_begin;
    d.q_begin = q_begin; d.q_end = q_end;
    d.active = nq > 0;
    d.runs_of_batch = 0;
    if (!d.active) return PSA_OK;

    const bool scan = ctx->engine == 2;
    const int tile = scan ? ctx->scan_tile : kExactTile;
    int rc;
    if ((rc = ensure_pin(ctx, d.h_out, sizeof(QueryRec) * nq + 16))) return rc;
    const int64_t byte0 = q_off[q_begin];
    const int64_t seq2_bytes = q_off[q_end] - byte0;
    const bool uniform_len = ctx->uniform_len2 > 0;       // every query of the batch has the same length
    int64_t tiles = 0, uniform = -1;
    if (uniform_len) {
        // nothing per query to build or upload: offsets and tile ranges are implicit on the device
        const int64_t f = last >= 0 ? first : 0, l = last >= 0 ? last : offsets_of(len1, ctx->uniform_len2);
        uniform = (l - tile_base(f) + tile - 1) / tile;
        tiles = uniform * nq;
        if (tiles > 0x7FFFFF00) return fail(ctx, PSA_ERR_ARG, "batch too large: more than 2^31 tiles on one GPU");
    } else {
        if ((rc = ensure_pin(ctx, d.h_qoff, sizeof(int64_t) * (nq + 1)))) return rc;
        if ((rc = ensure_pin(ctx, d.h_tile_start, sizeof(int32_t) * (nq + 1)))) return rc;
        int64_t* hq = (int64_t*)d.h_qoff.p;
        int32_t* ht = (int32_t*)d.h_tile_start.p;
        for (int k = 0; k < nq; k++) {
            hq[k] = q_off[q_begin + k] - byte0;
            const int64_t len2 = q_off[q_begin + k + 1] - q_off[q_begin + k];
            const int64_t f = last >= 0 ? first : 0, l = last >= 0 ? last : offsets_of(len1, len2);
            ht[k] = (int32_t)tiles;
            const int64_t mine = (l - tile_base(f) + tile - 1) / tile;
            uniform = k == 0 ? mine : (uniform == mine ? uniform : 0);
            tiles += mine;
            if (tiles > 0x7FFFFF00) return fail(ctx, PSA_ERR_ARG, "batch too large: more than 2^31 tiles on one GPU");
        }
        hq[nq] = seq2_bytes;
        ht[nq] = (int32_t)tiles;
    }

    // One query, exact order, few warp-tiles: the whole search as one cooperative launch (k_single)
    d.single = SingleGeom{};
    if (scan && nq == 1 && last >= 0 && ctx->table.exact && ctx->rank_planes <= 1 && ctx->opt_single_launch != 0 && ctx->opt_slices == 0 &&
        ctx->opt_scan_warps == 0 && ctx->max_len2 >= 64)
        d.single = single_plan(len1, ctx->max_len2, first, last, d.sm_count);

    // Slice mode: one query whose warp-tiles cannot fill the GPU is also cut along the alignment steps
    d.SG = SliceGeom{};
    d.SG.allow_derive = ctx->opt_derive_rank != 0;
    int fin_tile = tile;
    if (d.single.ok) {
        fin_tile = kCombineTile;
        tiles = (last - tile_base(first) + fin_tile - 1) / fin_tile;    // tile records come from the combine phase
        uniform = tiles;
        if ((rc = ensure_dev(ctx, d.partial, sizeof(uint2) * (size_t)d.single.slices * d.single.tiles * 1024))) return rc;
        if (!d.sync.p) {
            if ((rc = ensure_dev(ctx, d.sync, 64))) return rc;
            PSA_CUDA(ctx, cudaMemsetAsync(d.sync.p, 0, 64, d.stream));
        }
    } else
    if (scan && !ctx->batch_mode && nq == 1 && last >= 0 && ctx->opt_slices != 1 && ctx->max_len2 >= 256) {
        const int64_t span = last - tile_base(first);
        const int64_t warp_tiles = (span + 1023) / 1024;
        const int64_t steps_all = (ctx->max_len2 + 31) & ~int64_t(31);
        int64_t want = ctx->opt_slices > 1 ? ctx->opt_slices : (int64_t(d.sm_count) * 16) / warp_tiles;
        want = std::min<int64_t>(want, steps_all / 128);
        if (want >= 2) {
            const int64_t slice_len = ((steps_all + want - 1) / want + 127) & ~int64_t(127);
            d.SG.slice_len = (int)slice_len;
            d.SG.slices = (int)((steps_all + slice_len - 1) / slice_len);
            d.SG.scan_tile = tile;
            d.SG.scan_tiles = (int)tiles;
            if (d.SG.slices >= 2) {
                fin_tile = kCombineTile;
                tiles = (span + fin_tile - 1) / fin_tile;       // tile records now come from k_combine
                uniform = tiles;
                if ((rc = ensure_dev(ctx, d.partial, sizeof(uint2) * (size_t)d.SG.slices * d.SG.scan_tiles * d.SG.scan_tile))) return rc;
            } else {
                d.SG = SliceGeom{};
                d.SG.allow_derive = ctx->opt_derive_rank != 0;
            }
        }
    }

    d.SG.fused_finish = scan && !ctx->batch_mode && d.SG.slices <= 1 && ctx->table.exact && uniform == 1 &&
                        ctx->opt_fused_finish != 0;
    // slice mode with few combine blocks: the last of them finishes the query
    d.SG.fused_combine = scan && d.SG.slices > 1 && tiles <= 2 * (int64_t)d.sm_count && ctx->opt_fused_finish != 0;

    // Packed mode: equal-length queries that each fit one window share blocks lane by lane (k_scan_packed) when that
    // leaves fewer idle lanes than whole warps per query do.
    if (scan && !ctx->batch_mode && d.SG.slices <= 1 && uniform_len && uniform == 1 && last < 0 && ctx->opt_pack_queries != 0)
        psa_plan_packing(len1, ctx->uniform_len2, nq, ctx->opt_pack_queries >= 2 ? ctx->opt_pack_queries : 0, &d.SG.pack_q,
                         &d.SG.pack_warps);

    // Stripe mode: equal-length queries, exact integer order, all offsets, window fits shared memory -> one launch
    d.stripe = StripeGeom{};
    if (scan && uniform_len && last < 0 && ctx->table.exact && ctx->rank_planes <= 1 && ctx->opt_stripe_mode != 0 &&
        ctx->opt_sliced_keys != 0 && stripe_keys_ok(ctx->table, ctx->uniform_len2) &&
        (ctx->opt_stripe_mode == 1 || nq >= 2)) {
        // rank planes the window carries: none when the top rank follows from the class counts; else two (top and second rank,
        // interleaved at the class pitch: as cheap to read as one, and a short query's winner often has only the second) when
        // the table has a second rank to track and the window still fits, else one
        const bool rank_pass = ctx->rank_planes == 1 && !stripe_derives_rank(ctx->table, ctx->rank_planes, ctx->opt_derive_rank != 0);
        const int avail = ctx->table.nranks - (ctx->table.has_none ? 0 : 1);
        const bool force = ctx->opt_stripe_mode == 1;
        const int64_t key[5] = { len1, ctx->uniform_len2, nq, rank_pass ? (avail >= 2 ? 2 : 1) : 0, force ? 1 : 0 };
        if (std::memcmp(key, d.plan_key, sizeof(key)) == 0) d.stripe = d.plan_cached;
        else {
            if (rank_pass && avail >= 2) d.stripe = stripe_plan(len1, ctx->uniform_len2, nq, 2, d.sm_count, force);
            if (!d.stripe.ok) d.stripe = stripe_plan(len1, ctx->uniform_len2, nq, rank_pass ? 1 : 0, d.sm_count, force);
            std::memcpy(d.plan_key, key, sizeof(key));
            d.plan_cached = d.stripe;
        }
        // auto: short queries only in the one-warp-per-task shape (running best in bit planes: config 5 2.02 -> 1.39 ms); with
        // several warps per task their per-pass epilogue makes stripe mode no faster than batch mode's shared windows
        if (d.stripe.ok && !force && ctx->uniform_len2 < 128 && d.stripe.T != 1) d.stripe = StripeGeom{};
        if (d.stripe.ok) { d.SG.fused_finish = false; d.SG.fused_combine = false; d.SG.pack_q = d.SG.pack_warps = 0; }
    }

    const int64_t plane_words = scan_plane_words(len1);
    if ((rc = ensure_dev(ctx, d.code_table, kSymbols * kRowPad))) return rc;
    if ((rc = ensure_dev(ctx, d.table, sizeof(DeviceTable)))) return rc;
    if (d.table_epoch != ctx->table_epoch) {
        PSA_CUDA(ctx, cudaMemcpyAsync(d.table.p, &ctx->table, sizeof(DeviceTable), cudaMemcpyHostToDevice, d.stream));
        // the pair table in global memory for the kernels that index it per thread (divergent reads of a kernel
        // parameter are slow); uploaded when the table changes, not per batch
        PSA_CUDA(ctx, cudaMemcpyAsync(d.code_table.p, ctx->table.code, kSymbols * kRowPad, cudaMemcpyHostToDevice, d.stream));
        d.table_epoch = ctx->table_epoch;
    }
    if ((rc = ensure_dev(ctx, d.seq1, (size_t)len1 + 64))) return rc;
    if ((rc = ensure_dev(ctx, d.seq2s, (size_t)seq2_bytes + 64))) return rc;
    if (!uniform_len) {
        if ((rc = ensure_dev(ctx, d.qoff, sizeof(int64_t) * (nq + 1)))) return rc;
        if ((rc = ensure_dev(ctx, d.tile_start, sizeof(int32_t) * (nq + 1)))) return rc;
    }
    if ((rc = ensure_dev(ctx, d.tiles, sizeof(TileRec) * (size_t)tiles))) return rc;
    if ((rc = ensure_dev(ctx, d.out, sizeof(QueryRec) * nq + 16))) return rc;
    if (scan && !ctx->table.exact)
        if ((rc = ensure_dev(ctx, d.lane_keys, sizeof(int64_t) * (size_t)tiles * (fin_tile / 32)))) return rc;
    if (scan) {
        if ((rc = ensure_dev(ctx, d.cls_planes, sizeof(uint2) * (size_t)plane_words * kPlaneRows))) return rc;
        if ((rc = ensure_dev(ctx, d.rank_planes,
                             sizeof(uint32_t) * (size_t)plane_words * kPlaneRows * std::max(ctx->rank_planes, 1))))
            return rc;
    }

    // A small one-shot problem (the gpu_run_program case: a few KB of Seq1 and one query) goes up as ONE copy: both
    // sequences are gathered in a page-locked buffer first (a host memcpy of a few KB costs less than a second copy's
    // enqueue and its turn on the copy engine, ~2.5 us), and land in one device buffer.
    const void* dev_seq1 = d.seq1.p;
    const void* dev_seq2s = d.seq2s.p;
    const int64_t gather_at = (len1 + 64 + 255) & ~int64_t(255);            // the queries' place behind Seq1 and its padding
    const bool gathered = ctx->one_shot && ctx->opt_gather_small != 0 && gather_at + seq2_bytes <= kGatherMaxBytes;
    if (gathered) {
        if ((rc = ensure_dev(ctx, d.inbuf, (size_t)kGatherMaxBytes + 256))) return rc;
        if ((rc = ensure_pin(ctx, d.h_in, (size_t)kGatherMaxBytes + 256))) return rc;
        std::memcpy(d.h_in.p, seq1, (size_t)len1);
        std::memcpy((char*)d.h_in.p + gather_at, seq2s + byte0, (size_t)seq2_bytes);
        PSA_CUDA(ctx, cudaMemcpyAsync(d.inbuf.p, d.h_in.p, (size_t)(gather_at + seq2_bytes), cudaMemcpyHostToDevice, d.stream));
        dev_seq1 = d.inbuf.p;
        dev_seq2s = (const char*)d.inbuf.p + gather_at;
    } else
        PSA_CUDA(ctx, cudaMemcpyAsync(d.seq1.p, seq1, (size_t)len1, cudaMemcpyHostToDevice, d.stream));
    d.streamed = false;
    d.st_chunks = 0;
    d.P.ready = nullptr; d.P.ready_tag = 0; d.P.ready_chunks = 0; d.P.ready_chunk_bytes = 128;
    if (gathered) {
        // (both sequences are on their way)
    } else
    if (ctx->one_shot && d.stripe.ok && ctx->opt_stream_queries != 0 && seq2_bytes >= kStreamMinBytes) {
        if ((rc = enqueue_streamed_queries(ctx, d, seq2s + byte0, seq2_bytes))) return rc;
    } else
        PSA_CUDA(ctx, cudaMemcpyAsync(d.seq2s.p, seq2s + byte0, (size_t)seq2_bytes, cudaMemcpyHostToDevice, d.stream));
    if (!uniform_len) {
        PSA_CUDA(ctx, cudaMemcpyAsync(d.qoff.p, d.h_qoff.p, sizeof(int64_t) * (nq + 1), cudaMemcpyHostToDevice, d.stream));
        PSA_CUDA(ctx, cudaMemcpyAsync(d.tile_start.p, d.h_tile_start.p, sizeof(int32_t) * (nq + 1), cudaMemcpyHostToDevice, d.stream));
    }

    d.G.len1 = len1;
    d.G.first = last >= 0 ? first : 0;
    d.G.last = last >= 0 ? last : -1;
    d.G.nq = nq;
    d.G.tile = fin_tile;
    d.G.total_tiles = (int32_t)tiles;
    d.G.tiles_per_query = uniform > 0 ? (int32_t)uniform : 0;
    d.G.uniform_len2 = uniform_len ? (int32_t)ctx->uniform_len2 : 0;
    d.P.seq1 = (const uint8_t*)dev_seq1;
    d.P.seq2s = (const uint8_t*)dev_seq2s;
    d.P.qoff = (const int64_t*)d.qoff.p;
    d.P.tile_start = (const int32_t*)d.tile_start.p;
    d.P.tiles = (TileRec*)d.tiles.p;
    // Small result sets: the finishing threads store their 56-byte records straight into page-locked host memory, which
    // takes the device-to-host copy (an extra stream operation, ~4 us) off the end of the chain.  Large ones (config 5:
    // 3.7 MB) stay on the copy engine -- one DMA beats half a million 8-byte posted writes.
    // Only the one-shot calls do this: the split-phase form keeps its results resident until psa_batch_fetch.
    // Stripe mode stores a record as one 56-byte write (warp_store_record) spread over the whole run, so there any size goes.
    d.zc_out = ctx->one_shot && ctx->opt_zero_copy != 0 && (sizeof(QueryRec) * (size_t)nq <= kZeroCopyMaxBytes || d.stripe.ok);
    d.zc_direct = false;
    d.P.out = d.zc_out ? (QueryRec*)d.h_out.p : (QueryRec*)d.out.p;
    d.P.lane_keys = (int64_t*)d.lane_keys.p;
    d.P.code_table = (uint8_t*)d.code_table.p;
    d.P.partial = (uint2*)d.partial.p;
    d.P.partial_stride = d.single.ok ? int64_t(d.single.tiles) * 1024 : int64_t(d.SG.scan_tiles) * d.SG.scan_tile;
    d.P.sync = (int32_t*)d.sync.p;
    d.P.table = (const DeviceTable*)d.table.p;
    d.P.cand_count = (int32_t*)((char*)d.out.p + sizeof(QueryRec) * nq);     // the counter sits behind the records
    d.P.err_flag = d.h_err;                                                  // unified addressing: the host pointer is the device pointer
    d.P.cls_planes = (uint2*)d.cls_planes.p;
    d.P.rank_planes = (uint32_t*)d.rank_planes.p;
    d.P.plane_words = plane_words;
    return PSA_OK;
}

// A fresh tag per run: a kernel that meets a bad symbol stores the tag in the mapped word, the host compares after the
// stream has drained.  Nothing to clear between runs, nothing to copy back.
static int32_t next_run_tag(DeviceState& d)
{
    if (d.run_tag >= 0x7FFFFFF0) { d.run_tag = 0; ((volatile int32_t*)d.h_err)[0] = 0; ((volatile int32_t*)d.h_err)[1] = 0; }   // no run is in flight here
    return ++d.run_tag;
}
static bool bad_symbol_seen(const DeviceState& d) { return *(volatile const int32_t*)d.h_err == d.run_tag; }
static bool stream_wait_timed_out(const DeviceState& d) { return ((volatile const int32_t*)d.h_err)[1] == d.run_tag; }

int run_device(psa_context* ctx, DeviceState& d, bool timed)
{
    d.st_launches = d.st_tiles = d.st_main_ns = 0;
    if (!d.active) return PSA_OK;
    PSA_CUDA(ctx, cudaSetDevice(d.dev));
    TimedGate gate;
    gate.close(d, timed && ctx->opt_gate_timed_runs != 0 && d.runs_of_batch > 0);
    d.runs_of_batch++;
    if (timed) PSA_CUDA(ctx, cudaEventRecord(d.ev0, d.stream));
    d.P.run_tag = next_run_tag(d);
    if (ctx->engine == 2 && d.single.ok) {
        if (timed && ctx->opt_kernel_events) PSA_CUDA(ctx, cudaEventRecord(d.evk0, d.stream));
        launch_single(ctx->table, d.G, d.P, ctx->rank_planes, d.single, d.stream);
        if (timed && ctx->opt_kernel_events) PSA_CUDA(ctx, cudaEventRecord(d.evk1, d.stream));
        d.st_launches += 1;
        PSA_CUDA(ctx, cudaGetLastError());
        if (timed) PSA_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
        d.st_tiles += d.G.total_tiles;
        return PSA_OK;
    }
    if (ctx->engine == 2 && d.stripe.ok) {
        if (timed && ctx->opt_kernel_events) PSA_CUDA(ctx, cudaEventRecord(d.evk0, d.stream));
        launch_stripe(ctx->table, d.G, d.P, ctx->rank_planes, ctx->opt_derive_rank != 0, d.stripe, d.stream);
        if (timed && ctx->opt_kernel_events) PSA_CUDA(ctx, cudaEventRecord(d.evk1, d.stream));
        d.st_launches += 1;
        PSA_CUDA(ctx, cudaGetLastError());
        if (timed) PSA_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
        d.st_tiles += d.G.total_tiles;
        return PSA_OK;
    }
    if (ctx->engine == 2) {
        launch_profile(ctx->table, d.G, d.P, ctx->rank_planes, d.sm_count, d.stream);
        if (timed && ctx->opt_kernel_events) PSA_CUDA(ctx, cudaEventRecord(d.evk0, d.stream));
        launch_scan(ctx->table, d.G, d.P, ctx->rank_planes, ctx->max_len2, ctx->batch_mode, ctx->opt_sliced_keys != 0, d.sm_count, d.SG, d.stream);
        if (d.SG.slices > 1) d.st_launches += 1;
        if (timed && ctx->opt_kernel_events) PSA_CUDA(ctx, cudaEventRecord(d.evk1, d.stream));
        d.st_launches += 2;
    } else {
        if (timed && ctx->opt_kernel_events) PSA_CUDA(ctx, cudaEventRecord(d.evk0, d.stream));
        launch_exact_tiles(ctx->table, d.G, d.P, d.stream);
        if (timed && ctx->opt_kernel_events) PSA_CUDA(ctx, cudaEventRecord(d.evk1, d.stream));
        d.st_launches += 1;
    }
    if (!d.SG.fused_finish && !d.SG.fused_combine) {
        launch_finish(ctx->table, d.G, d.P, ctx->engine == 2, d.stream);
        d.st_launches += 1;
    }
    PSA_CUDA(ctx, cudaGetLastError());
    if (timed) PSA_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
    d.st_tiles += d.G.total_tiles;
    return PSA_OK;
}

// reference order on (score, offset): strictly better score, or equal score and lower offset
inline bool swapable(int is_max, double s_old, int off_old, double s_new, int off_new)
{
    if (is_max ? s_new > s_old : s_new < s_old) return true;
    return s_new == s_old && off_new < off_old;
}

static_assert(sizeof(psa_problem) == 64 && offsetof(psa_problem, seq1) == 16 && offsetof(psa_problem, out) == 48 && offsetof(psa_problem, status) == 56,
              "psa_problem is what the ctypes mirror (and any FFI binding) lays out");
static_assert(sizeof(QueryRec) == sizeof(psa_result), "device record == public result");
static_assert(offsetof(QueryRec, ch) == offsetof(psa_result, mutant) + offsetof(psa_mutant, ch), "ch");
static_assert(offsetof(QueryRec, rank) == offsetof(psa_result, rank) && offsetof(QueryRec, score) == offsetof(psa_result, score) &&
              offsetof(QueryRec, counts) == offsetof(psa_result, counts), "device record == public result");

} // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int psa_abi_version(void) { return PSA_ABI_VERSION; }

const char* psa_strerror(int status)
{
    switch (status) {
    case PSA_OK: return "ok";
    case PSA_ERR_ARG: return "invalid argument";
    case PSA_ERR_ALPHABET: return "sequence symbol outside [A-Z-]";
    case PSA_ERR_WEIGHTS: return "weights must be finite";
    case PSA_ERR_CUDA: return "CUDA error or no usable sm_100 GPU";
    case PSA_ERR_NOMEM: return "out of memory";
    case PSA_ERR_STATE: return "no batch prepared";
    case PSA_ERR_IO: return "file I/O error";
    }
    return "unknown status";
}

int psa_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

int psa_create(psa_context** out, const int* devices, int ndevices)
{
    if (!out || ndevices < 1) return PSA_ERR_ARG;
    *out = nullptr;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) { cudaGetLastError(); return PSA_ERR_CUDA; }
    psa_context* ctx = new (std::nothrow) psa_context();
    if (!ctx) return PSA_ERR_NOMEM;
    ctx->devs.resize(ndevices);
    for (int i = 0; i < ndevices; i++) {
        DeviceState& d = ctx->devs[i];
        d.dev = devices ? devices[i] : i;
        int major = 0;
        if (d.dev < 0 || d.dev >= visible || cudaSetDevice(d.dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d.dev) != cudaSuccess || major != 10 ||
            cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.dev) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&d.ev0) != cudaSuccess || cudaEventCreate(&d.ev1) != cudaSuccess ||
            cudaEventCreate(&d.evk0) != cudaSuccess || cudaEventCreate(&d.evk1) != cudaSuccess ||
            cudaHostAlloc((void**)&d.h_err, 64, cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            for (DeviceState& x : ctx->devs) release(x);
            delete ctx;
            return PSA_ERR_CUDA;     // the kernels are sm_100a only: no other device can run them
        }
    }
    for (DeviceState& d : ctx->devs) std::memset(d.h_err, 0, 64);
    for (int i = 1; i < ndevices; i++) {
        ctx->workers.emplace_back(new Worker());
        Worker* w = ctx->workers.back().get();
        w->th = std::thread(worker_loop, w);
    }
    *out = ctx;
    return PSA_OK;
}

void psa_destroy(psa_context* ctx)
{
    if (!ctx) return;
    for (auto& w : ctx->lane_workers) {
        if (!w) continue;
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
    }
    ctx->lane_workers.clear();
    for (psa_context* lane : ctx->lanes) psa_destroy(lane);
    ctx->lanes.clear();
    for (auto& w : ctx->workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
    }
    for (DeviceState& d : ctx->devs) {
        cudaSetDevice(d.dev);
        if (d.stream) cudaStreamSynchronize(d.stream);
        release(d);
    }
    delete ctx;
}

const char* psa_last_error(const psa_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int psa_set_option(psa_context* ctx, const char* name, long long value)
{
    if (!ctx || !name) return PSA_ERR_ARG;
    for (DeviceState& d : ctx->devs) d.runs_of_batch = 0;          // a knob may change which kernels the next run launches
    if (!std::strcmp(name, "engine") && value >= 0 && value <= 2) { ctx->opt_engine = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "rank_planes") && value >= -1 && value <= 8) { ctx->opt_rank_planes = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "fused_finish") && value >= 0 && value <= 1) { ctx->opt_fused_finish = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "derive_rank") && value >= 0 && value <= 1) { ctx->opt_derive_rank = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "pack_queries") && value >= 0 && value <= kPackMaxQ) { ctx->opt_pack_queries = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "zero_copy_results") && value >= 0 && value <= 1) { ctx->opt_zero_copy = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "stream_queries") && value >= 0 && value <= 1) { ctx->opt_stream_queries = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "gather_small") && value >= 0 && value <= 1) { ctx->opt_gather_small = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "gate_timed_runs") && value >= 0 && value <= 1) { ctx->opt_gate_timed_runs = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "min_split_work") && value >= 0) { ctx->opt_min_split_work = value; return PSA_OK; }
    if (!std::strcmp(name, "kernel_events") && value >= 0 && value <= 1) { ctx->opt_kernel_events = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "slices") && value >= 0 && value <= 256) { ctx->opt_slices = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "sliced_keys") && value >= 0 && value <= 1) { ctx->opt_sliced_keys = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "batch_mode") && value >= -1 && value <= 1) { ctx->opt_batch_mode = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "stripe_mode") && value >= -1 && value <= 1) { ctx->opt_stripe_mode = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "single_launch") && value >= 0 && value <= 1) { ctx->opt_single_launch = (int)value; return PSA_OK; }
    if (!std::strcmp(name, "scan_warps") && value >= 0 && value <= kScanWarps) { ctx->opt_scan_warps = (int)value; return PSA_OK; }
    return PSA_ERR_ARG;
}

long long psa_get_stat(const psa_context* ctx, const char* name)
{
    if (!ctx || !name) return -1;
    long long launches = 0, tiles = 0, main_ns = 0;
    for (const DeviceState& d : ctx->devs) {
        launches += d.st_launches; tiles += d.st_tiles;
        main_ns = std::max(main_ns, d.st_main_ns);
    }
    if (!std::strcmp(name, "kernel_launches")) return launches;
    // host-side split of the last psa_search_batch (ns): planning on the calling thread, then per device slot (max over
    // slots) H2D enqueue, kernel launches, and the wait for the stream + results
    if (!std::strcmp(name, "host_plan_ns")) return ctx->st_plan_ns;
    if (!std::strcmp(name, "host_total_ns")) return ctx->st_total_ns;
    if (!std::strcmp(name, "host_prepare_ns") || !std::strcmp(name, "host_enqueue_ns") || !std::strcmp(name, "host_wait_ns")) {
        long long worst = 0;
        for (const DeviceState& d : ctx->devs)
            if (d.active) worst = std::max(worst, name[5] == 'p' ? d.st_prepare_ns : name[5] == 'e' ? d.st_enqueue_ns : d.st_wait_ns);
        return worst;
    }
    if (!std::strcmp(name, "candidate_tiles")) {
        // a statistic nobody waits for: the counter stays on the device and is read only when asked
        long long cand = 0;
        for (const DeviceState& d : ctx->devs) {
            if (!d.active || !ctx->ran || !d.P.cand_count) continue;
            int32_t v = 0;
            if (cudaSetDevice(d.dev) != cudaSuccess || cudaStreamSynchronize(d.stream) != cudaSuccess ||
                cudaMemcpy(&v, d.P.cand_count, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return -1; }
            cand += v;
        }
        return cand;
    }
    if (!std::strcmp(name, "tiles")) return tiles;
    if (!std::strcmp(name, "main_kernel_ns")) return main_ns;           // dominant kernel of the last psa_batch_run
    if (!std::strcmp(name, "engine")) return ctx->engine;
    if (!std::strcmp(name, "rank_planes")) return ctx->rank_planes;
    if (!std::strcmp(name, "scan_warps")) return ctx->scan_tile / 1024;
    auto first_active = [ctx]() -> const DeviceState* { for (const DeviceState& d : ctx->devs) if (d.active) return &d; return nullptr; };
    if (!std::strcmp(name, "stripe_mode")) { const DeviceState* d = first_active(); return d && d->stripe.ok ? 1 : 0; }
    if (!std::strcmp(name, "streamed_chunks")) { long long n = 0; for (const DeviceState& d : ctx->devs) if (d.active) n = std::max<long long>(n, d.st_chunks); return n; }
    if (!std::strcmp(name, "single_launch")) { const DeviceState* d = first_active(); return d && d->single.ok ? 1 : 0; }
    if (!std::strcmp(name, "stripe_queries_per_task")) { const DeviceState* d = first_active(); return d && d->stripe.ok ? d->stripe.Q : 0; }
    if (!std::strcmp(name, "stripe_team_warps")) { const DeviceState* d = first_active(); return d && d->stripe.ok ? d->stripe.T : 0; }
    if (!std::strcmp(name, "stripe_teams")) { const DeviceState* d = first_active(); return d && d->stripe.ok ? d->stripe.teams : 0; }
    if (!std::strcmp(name, "devices_used")) { long long n = 0; for (const DeviceState& d : ctx->devs) n += d.active ? 1 : 0; return n; }
    if (!std::strcmp(name, "stripe_split")) { const DeviceState* d = first_active(); return d && d->stripe.ok ? d->stripe.split : 0; }
    if (!std::strcmp(name, "stripe_lanes")) { const DeviceState* d = first_active(); return d && d->stripe.ok ? d->stripe.S : 0; }
    if (!std::strcmp(name, "batch_mode")) { const DeviceState* d = first_active(); return ctx->batch_mode && !(d && d->stripe.ok) ? 1 : 0; }
    if (!std::strcmp(name, "slices")) { for (const DeviceState& d : ctx->devs) if (d.active) return d.single.ok ? d.single.slices : d.SG.slices; return 1; }
    if (!std::strcmp(name, "packed_queries")) { for (const DeviceState& d : ctx->devs) if (d.active) return d.SG.pack_q; return 0; }
    if (!std::strcmp(name, "packed_warps")) { for (const DeviceState& d : ctx->devs) if (d.active) return d.SG.pack_warps; return 0; }
    if (!std::strcmp(name, "exact")) return ctx->table.exact;
    return -1;
}

int psa_plan_packing(int64_t len1, int64_t len2, int32_t nq, int force, int* queries_per_block, int* warps)
{
    if (!queries_per_block || !warps) return PSA_ERR_ARG;
    *queries_per_block = *warps = 0;
    if (len1 < 1 || len2 < 1 || len2 > len1 || nq < 0 || force < 0 || force == 1 || force > kPackMaxQ) return PSA_ERR_ARG;
    if (nq < 2 || !scan_packed_fits(len1, len2)) return PSA_OK;
    const int64_t lanes = (offsets_of(len1, len2) + 31) / 32;                        // per query
    int best_q = 0, best_w = 0;
    if (force >= 2) {
        const int64_t w = (force * lanes + 31) / 32;
        if (w <= kPackMaxWarps) { best_q = force; best_w = (int)w; }
    } else {
        const double plain = double(lanes) / double(32 * ((lanes + 31) / 32));
        double best_u = 0.0;
        for (int w = 1; w <= kPackMaxWarps; w++) {
            const int q = (int)std::min<int64_t>(std::min<int64_t>(kPackMaxQ, nq), (32 * w) / lanes);
            if (q < 2) continue;
            const int ww = (int)((q * lanes + 31) / 32);                             // warps those q queries really need
            const double u = double(q * lanes) / double(32 * ww);
            if (u > best_u + 1e-9) { best_u = u; best_q = q; best_w = ww; }
        }
        if (best_u < plain * 1.05) best_q = best_w = 0;                              // not worth leaving the plain kernel
    }
    *queries_per_block = best_q;
    *warps = best_q ? best_w : 0;
    return PSA_OK;
}

int psa_plan_stream_pieces(int64_t bytes, int* pieces, int64_t* piece_bytes)
{
    if (!pieces || !piece_bytes || bytes < 0) return PSA_ERR_ARG;
    *pieces = 0; *piece_bytes = 0;
    if (bytes < kStreamMinBytes) return PSA_OK;                             // one plain copy in front of the kernel
    int64_t chunks = std::min<int64_t>(std::max<int64_t>(bytes / kStreamChunkBytes, 1), kStreamMaxChunks);
    const int64_t cb = ((bytes + chunks - 1) / chunks + 127) & ~int64_t(127);
    *pieces = int((bytes + cb - 1) / cb);
    *piece_bytes = cb;
    return PSA_OK;
}

int psa_plan_stripes(int64_t len1, int64_t len2, int32_t nq, int rank_pass, int sm_count, int shape[8])
{
    if (!shape || len1 < 1 || len2 < 1 || len2 > len1 || nq < 0 || sm_count < 1) return PSA_ERR_ARG;
    const StripeGeom g = stripe_plan(len1, len2, nq, rank_pass < 0 ? 0 : rank_pass > 2 ? 2 : rank_pass, sm_count);
    const int v[8] = { g.ok, g.S, g.Q, g.passes, g.T, g.teams, g.blocks, int(g.smem) };
    for (int k = 0; k < 8; k++) shape[k] = v[k];
    return PSA_OK;
}

int psa_plan_shards(int64_t len1, const int64_t* q_off, int32_t nq, int nshards, int64_t granule,
                    int64_t first, int64_t last, psa_shard* out)
{
    if (!q_off || !out || nq < 0 || nshards < 1 || granule < 1 || len1 < 1) return PSA_ERR_ARG;
    for (int g = 0; g < nshards; g++) out[g] = psa_shard{ 0, 0, -1, -1 };
    if (nq == 0) return PSA_OK;
    if (nq == 1) {
        const int64_t len2 = q_off[1] - q_off[0];
        if (len2 < 1 || len2 > len1) return PSA_ERR_ARG;
        const int64_t f = last >= 0 ? first : 0, l = last >= 0 ? last : offsets_of(len1, len2);
        if (f < 0 || f >= l || l > offsets_of(len1, len2)) return PSA_ERR_ARG;
        // whole granules from the aligned base, so no shard gets a sliver and plane words stay aligned
        const int64_t base = tile_base(f);
        const int64_t ngran = (l - base + granule - 1) / granule;
        const int use = (int)std::min<int64_t>(nshards, ngran);
        int64_t t0 = 0;
        for (int g = 0; g < use; g++) {
            const int64_t t1 = ngran * (g + 1) / use;
            out[g].q_begin = 0; out[g].q_end = 1;
            out[g].first = std::max(f, base + t0 * granule);
            out[g].last = std::min(l, base + t1 * granule);
            t0 = t1;
        }
        return PSA_OK;
    }
    if (last >= 0) return PSA_ERR_ARG;               // an offset range is only defined for one query
    if (nshards == 1) {                              // nothing to balance (lengths are validated by the caller's pass)
        out[0].q_begin = 0; out[0].q_end = nq;
        return PSA_OK;
    }
    // one pass: validate, total work, and whether every query has the same length (then the balanced split is arithmetic
    // -- a batch of 65 536 equal-length queries must not pay for a prefix-sum array on every call)
    const int64_t len_first = q_off[1] - q_off[0];
    bool uniform = true;
    double total = 0.0;
    for (int q = 0; q < nq; q++) {
        const int64_t len2 = q_off[q + 1] - q_off[q];
        if (len2 < 1 || len2 > len1) return PSA_ERR_ARG;
        uniform = uniform && len2 == len_first;
        total += double(offsets_of(len1, len2)) * double(len2);
    }
    if (uniform) {
        for (int g = 0; g < nshards; g++) {
            out[g].q_begin = int32_t(int64_t(nq) * g / nshards);
            out[g].q_end = int32_t(int64_t(nq) * (g + 1) / nshards);
        }
        return PSA_OK;
    }
    // ragged: walk the running sum once more; a shard ends at the query boundary nearest to its share of the work
    int qb = 0, q = 0;
    double done = 0.0;                               // work of queries [0, q)
    for (int g = 0; g < nshards; g++) {
        int qe;
        if (g == nshards - 1) qe = nq;
        else {
            const double target = total * double(g + 1) / double(nshards);
            double next = done;
            while (q < nq) {
                const int64_t len2 = q_off[q + 1] - q_off[q];
                next = done + double(offsets_of(len1, len2)) * double(len2);
                if (next >= target) break;
                done = next;
                q++;
            }
            // boundary q (work `done` before it) or q + 1 (work `next`), whichever is nearer to the target
            qe = q;
            if (q < nq && next - target <= target - done) { qe = q + 1; done = next; q++; }
            qe = std::max(qb, std::min(qe, nq));
        }
        out[g].q_begin = qb; out[g].q_end = qe;
        qb = qe;
    }
    return PSA_OK;
}

int psa_merge_results(int is_max, const psa_result* parts, int nparts, psa_result* out)
{
    if (!parts || !out || nparts < 1) return PSA_ERR_ARG;
    bool have = false;
    psa_result best;
    std::memset(&best, 0, sizeof(best));
    for (int k = 0; k < nparts; k++) {
        const psa_result& cur = parts[k];
        if (cur.mutant.offset < 0 || cur.mutant.ch == '\0') continue;
        if (!have || swapable(is_max, best.score, best.mutant.offset, cur.score, cur.mutant.offset)) { best = cur; have = true; }
    }
    if (!have) {
        best.mutant.offset = -1; best.mutant.char_offset = -1; best.mutant.ch = '\0';
        best.score = is_max ? -INFINITY : INFINITY;
    }
    *out = best;
    return PSA_OK;
}

static int prepare_common(psa_context* ctx, const double* weights, int is_max, const char* seq1, int64_t len1,
                          const char* seq2s, const int64_t* q_off, int32_t nq, int64_t first, int64_t last)
{
    if (!ctx) return PSA_ERR_ARG;
    ctx->prepared = ctx->ran = false;
    if (!weights || !seq1 || !seq2s || !q_off || nq < 0) return fail(ctx, PSA_ERR_ARG, "null argument or nq < 0");
    if (len1 < 1 || len1 > 0x7FFF0000ll) return fail(ctx, PSA_ERR_ARG, "len1 out of range");
    int64_t max_len2 = 0, min_len2 = INT64_MAX;
    {
        // equal-length batches (the common large case: config 5 has 65 536 queries) are recognised by a branch-free pass the
        // compiler vectorises; only ragged batches pay for the min / max walk
        const int64_t len0 = nq > 0 ? q_off[1] - q_off[0] : 0;
        int64_t differs = 0;
        for (int q = 0; q < nq; q++) differs |= (q_off[q + 1] - q_off[q]) ^ len0;
        if (nq > 0 && differs == 0) {
            if (len0 < 1 || len0 > len1) return fail(ctx, PSA_ERR_ARG, "query 0: len2=%lld must be in [1, len1]", (long long)len0);
            max_len2 = min_len2 = len0;
        } else {
            for (int q = 0; q < nq; q++) {
                const int64_t len2 = q_off[q + 1] - q_off[q];
                if (len2 < 1 || len2 > len1) return fail(ctx, PSA_ERR_ARG, "query %d: len2=%lld must be in [1, len1]", q, (long long)len2);
                max_len2 = std::max(max_len2, len2);
                min_len2 = std::min(min_len2, len2);
            }
        }
    }
    ctx->uniform_len2 = nq > 0 && min_len2 == max_len2 ? max_len2 : 0;
    if (max_len2 > kExactMaxLen2) return fail(ctx, PSA_ERR_ARG, "len2 > %lld is not supported", (long long)kExactMaxLen2);
    if (last >= 0) {
        if (nq != 1 || first < 0 || first >= last || last > offsets_of(len1, max_len2))
            return fail(ctx, PSA_ERR_ARG, "offset range [%lld,%lld) invalid", (long long)first, (long long)last);
    }
    int rc;
    // the table depends on (weights, goal) and, through the exactness analysis, on the longest query
    if (!ctx->table_valid || std::memcmp(ctx->weights, weights, sizeof(ctx->weights)) != 0 ||
        ctx->is_max != (is_max ? 1 : 0) || ctx->table_len2 != max_len2) {
        ctx->table_valid = false;
        rc = build_tables(weights, is_max, std::max<int64_t>(max_len2, 1), nullptr, &ctx->table);
        if (rc) return fail(ctx, rc, "%s", psa_strerror(rc));
        std::memcpy(ctx->weights, weights, sizeof(ctx->weights));
        ctx->table_epoch++;
        ctx->is_max = is_max ? 1 : 0;
        ctx->table_len2 = max_len2;
        ctx->table_valid = true;
    }
    ctx->nq = nq;
    ctx->max_len2 = max_len2;
    ctx->engine = ctx->opt_engine ? ctx->opt_engine : kDefaultEngine;
    if (ctx->engine == 2 && max_len2 > kScanMaxLen2) ctx->engine = 1;
    ctx->rank_planes = ctx->engine == 2 ? pick_rank_planes(ctx) : 0;
    if (ctx->engine == 2) {
        // warps per block (1024 offsets each): the count that wastes the fewest idle warp-tiles, larger on ties
        double best_cost = 0;
        int best_w = kScanWarps;
        for (int w = kScanWarps; w >= 1; w--) {
            double cost = 0;
            auto tiles_of = [&](int64_t len2) {
                const int64_t n = last >= 0 ? last - tile_base(first) : offsets_of(len1, len2);
                return double((n + 1024 * w - 1) / (1024 * w)) * w;
            };
            if (ctx->uniform_len2 > 0) cost = tiles_of(ctx->uniform_len2) * nq;
            else
                for (int q = 0; q < nq; q++) cost += tiles_of(q_off[q + 1] - q_off[q]);
            if (w == kScanWarps || cost < best_cost) { best_cost = cost; best_w = w; }
        }
        ctx->scan_tile = ctx->opt_scan_warps > 0 ? 1024 * ctx->opt_scan_warps : 1024 * best_w;
        // batch mode (every query fits one window): tiles are single warps that share a staged window
        BatchGeom probe{};
        probe.last = last;
        probe.len1 = len1;
        probe.nq = (nq + (int)ctx->devs.size() - 1) / (int)ctx->devs.size();
        // (a single query always runs on an explicit offset range per GPU, which batch mode does not take)
        ctx->batch_mode = nq > 1 && ctx->opt_batch_mode != 0 && (ctx->opt_batch_mode == 1 || ctx->opt_scan_warps == 0) &&
                          scan_batch_mode(probe, max_len2, ctx->opt_batch_mode == 1 ? 0 : ctx->devs[0].sm_count);
        if (ctx->batch_mode) ctx->scan_tile = 1024;
    }

    const int ndev = (int)ctx->devs.size();
    ctx->range_split = false;
    for (DeviceState& d : ctx->devs) d.active = false;
    if (nq == 0) { ctx->prepared = true; return PSA_OK; }

    const int64_t granule = ctx->engine == 2 ? ctx->scan_tile : kExactTile;
    ctx->plan.assign(ndev, psa_shard{ 0, 0, -1, -1 });
    // A call is worth spreading over the GPUs only when each gets enough to do: a shard pays its own copies, launch and
    // wake-up (~30 us) whatever its size, and eight host threads enqueue no faster than one.  Below `min_split_work` pair
    // evaluations per GPU (2.5e9 = ~80 us of kernel) fewer GPUs take the call -- config 3 (1.3e9) and config 4 (2e9) run on
    // one, config 5 (4.2e10) on all eight.  Streams of small problems scale through psa_search_many instead.
    int use = ndev;
    if (ndev > 1 && ctx->opt_min_split_work > 0) {
        double work = 0.0;
        if (ctx->uniform_len2 > 0) work = double(nq) * double(last >= 0 ? last - first : offsets_of(len1, max_len2)) * double(max_len2);
        else for (int q = 0; q < nq; q++) work += double(offsets_of(len1, q_off[q + 1] - q_off[q])) * double(q_off[q + 1] - q_off[q]);
        use = int(std::min<double>(ndev, std::max(1.0, std::floor(work / double(ctx->opt_min_split_work)))));
    }
    if (nq > 1 && ctx->uniform_len2 > 0) {
        // lengths were validated above and are all equal: the balanced split is arithmetic (no second pass over q_off)
        for (int g = 0; g < use; g++) {
            ctx->plan[g].q_begin = int32_t(int64_t(nq) * g / use);
            ctx->plan[g].q_end = int32_t(int64_t(nq) * (g + 1) / use);
        }
    } else if ((rc = psa_plan_shards(len1, q_off, nq, use, granule, first, last, ctx->plan.data())))
        return fail(ctx, rc, "cannot partition the batch");
    int used = 0;
    for (int g = 0; g < ndev; g++) used += ctx->plan[g].q_begin != ctx->plan[g].q_end;
    ctx->range_split = nq == 1 && used > 1;
    ctx->in_seq1 = seq1; ctx->in_len1 = len1; ctx->in_seq2s = seq2s; ctx->in_qoff = q_off;
    ctx->prepared = true;
    return PSA_OK;
}

// the device-side half of prepare: copies and geometry of this GPU's shard (inputs must still be alive)
static int prepare_shard(psa_context* ctx, DeviceState& d)
{
    const psa_shard& sh = ctx->plan[&d - ctx->devs.data()];
    if (sh.q_begin == sh.q_end) { d.active = false; return PSA_OK; }
    return prepare_device(ctx, d, ctx->in_seq1, ctx->in_len1, ctx->in_seq2s, ctx->in_qoff, sh.q_begin, sh.q_end,
                          ctx->nq == 1 ? sh.first : -1, ctx->nq == 1 ? sh.last : -1);
}

// device-to-host copy of this GPU's records (+ flags) and the wait for everything enqueued before it
static int fetch_shard(psa_context* ctx, DeviceState& d, psa_result* out, bool direct)
{
    if (!d.active) return PSA_OK;
    PSA_CUDA(ctx, cudaSetDevice(d.dev));
    const size_t bytes = sizeof(QueryRec) * d.G.nq;
    if (d.zc_out) {
        PSA_CUDA(ctx, cudaStreamSynchronize(d.stream));
        if (!d.zc_direct && ctx->nq > 1) std::memcpy(out + d.q_begin, d.h_out.p, bytes);
        return PSA_OK;
    }
    PSA_CUDA(ctx, cudaMemcpyAsync(direct ? (void*)(out + d.q_begin) : d.h_out.p, d.out.p, bytes, cudaMemcpyDeviceToHost, d.stream));
    PSA_CUDA(ctx, cudaStreamSynchronize(d.stream));
    if (!direct && ctx->nq > 1) std::memcpy(out + d.q_begin, d.h_out.p, bytes);
    return PSA_OK;
}

// after every shard has been fetched: alphabet check and, for a single query, the merge of the per-GPU answers
static int finish_fetch(psa_context* ctx, psa_result* out)
{
    for (DeviceState& d : ctx->devs)
        if (d.active && stream_wait_timed_out(d)) return fail(ctx, PSA_ERR_CUDA, "streamed queries did not reach cuda:%d in time", d.dev);
    for (DeviceState& d : ctx->devs)
        if (d.active && bad_symbol_seen(d)) return fail(ctx, PSA_ERR_ALPHABET, "%s", psa_strerror(PSA_ERR_ALPHABET));
    if (ctx->nq == 1) {
        // per-GPU candidates of the single query in ascending offset order
        std::vector<psa_result> parts;
        for (DeviceState& d : ctx->devs) {
            if (!d.active) continue;
            psa_result cur;
            std::memcpy(&cur, d.h_out.p, sizeof(cur));
            parts.push_back(cur);
        }
        return psa_merge_results(ctx->is_max, parts.data(), (int)parts.size(), out);
    }
    return PSA_OK;
}

static bool result_array_is_pinned(const psa_context* ctx, const psa_result* out)
{
    // Records have the caller's layout.  A page-locked result array is filled by the copy engine directly; a pageable
    // one goes through the context's pinned buffer and one memcpy.  The single-query case always stages (merge).
    if (ctx->nq <= 1) return false;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, out) == cudaSuccess) return attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return false;
}

// the address kernels can use to write a page-locked host array in place, or null
static void* device_view_of(const void* host)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost) return attr.devicePointer;
    cudaGetLastError();
    return nullptr;
}

int psa_batch_prepare(psa_context* ctx, const double weights[4], int is_max, const char* seq1, int64_t len1,
                      const char* seq2s, const int64_t* q_off, int32_t nq)
{
    int rc = prepare_common(ctx, weights, is_max, seq1, len1, seq2s, q_off, nq, -1, -1);
    if (rc) return rc;
    if (nq == 0) return PSA_OK;
    // the split-phase form promises a resident batch: wait for the copies here
    ctx->one_shot = false;
    return for_each_device(ctx, [ctx](DeviceState& d) {
        int r = prepare_shard(ctx, d);
        if (r || !d.active) return r;
        PSA_CUDA(ctx, cudaStreamSynchronize(d.stream));
        return (int)PSA_OK;
    });
}

int psa_batch_run(psa_context* ctx, float* device_ms)
{
    if (!ctx || !ctx->prepared) return ctx ? fail(ctx, PSA_ERR_STATE, "no batch prepared") : PSA_ERR_ARG;
    int rc = for_each_device(ctx, [ctx](DeviceState& d) {
        int r = run_device(ctx, d, true);
        d.run_ms = 0.f;
        if (r || !d.active) return r;
        PSA_CUDA(ctx, cudaEventSynchronize(d.ev1));
        PSA_CUDA(ctx, cudaEventElapsedTime(&d.run_ms, d.ev0, d.ev1));
        if (ctx->opt_kernel_events) {
            float kms = 0.f;
            PSA_CUDA(ctx, cudaEventElapsedTime(&kms, d.evk0, d.evk1));
            d.st_main_ns = (long long)(kms * 1e6);
        }
        return (int)PSA_OK;
    });
    if (rc) return rc;
    ctx->ran = true;
    float worst = 0.f;
    for (const DeviceState& d : ctx->devs) worst = std::max(worst, d.run_ms);
    if (device_ms) *device_ms = worst;
    return PSA_OK;
}

int psa_batch_fetch(psa_context* ctx, psa_result* out)
{
    if (!ctx || !ctx->prepared || !ctx->ran) return ctx ? fail(ctx, PSA_ERR_STATE, "no batch has run") : PSA_ERR_ARG;
    if (ctx->nq == 0) return PSA_OK;
    if (!out) return fail(ctx, PSA_ERR_ARG, "null result buffer");
    const bool direct = result_array_is_pinned(ctx, out);
    int rc = for_each_device(ctx, [ctx, out, direct](DeviceState& d) { return fetch_shard(ctx, d, out, direct); });
    return rc ? rc : finish_fetch(ctx, out);
}

// prepare + run + fetch in one round per GPU (what psa_search_batch / psa_search_range do)
static int search_prepared(psa_context* ctx, psa_result* out)
{
    if (ctx->nq == 0) return PSA_OK;
    if (!out) return fail(ctx, PSA_ERR_ARG, "null result buffer");
    const bool direct = result_array_is_pinned(ctx, out);
    ctx->one_shot = true;
    int rc = for_each_device(ctx, [ctx, out, direct](DeviceState& d) {
        using clk = std::chrono::steady_clock;
        auto ns = [](clk::time_point a, clk::time_point b) { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
        const auto t0 = clk::now();
        int r = prepare_shard(ctx, d);
        if (!r && d.active && d.zc_out && direct) {                 // the caller's array is page-locked: write it in place
            if (void* dv = device_view_of(out + d.q_begin)) {
                d.P.out = static_cast<QueryRec*>(dv);
                d.zc_direct = true;
            }
        }
        const auto t1 = clk::now();
        if (!r) r = run_device(ctx, d, false);
        const auto t2 = clk::now();
        if (!r) r = fetch_shard(ctx, d, out, direct);
        const auto t3 = clk::now();
        d.st_prepare_ns = ns(t0, t1); d.st_enqueue_ns = ns(t1, t2); d.st_wait_ns = ns(t2, t3);
        if (d.zc_direct) {                                          // the caller's array is theirs again: a later
            d.P.out = (QueryRec*)d.h_out.p;                         // psa_batch_run on this batch writes the staging buffer
            d.zc_direct = false;
        }
        if (d.streamed) {
            // a finished kernel has seen every flag, so the copies are done; on an error path make sure of it before the
            // caller's buffers are theirs again.  The batch is resident from here on (psa_batch_run does not wait for flags).
            if (r) { cudaStreamSynchronize(d.copy_stream); cudaGetLastError(); }
            d.P.ready = nullptr;
            d.streamed = false;
        }
        return r;
    });
    ctx->ran = rc == PSA_OK;
    return rc ? rc : finish_fetch(ctx, out);
}

int psa_search_batch(psa_context* ctx, const double weights[4], int is_max, const char* seq1, int64_t len1,
                     const char* seq2s, const int64_t* q_off, int32_t nq, psa_result* out)
{
    static const bool trace = std::getenv("PSA_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = prepare_common(ctx, weights, is_max, seq1, len1, seq2s, q_off, nq, -1, -1);
    if (rc) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    rc = search_prepared(ctx, out);
    const auto t2 = std::chrono::steady_clock::now();
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    ctx->st_plan_ns = (long long)(us(t0, t1) * 1e3);
    ctx->st_total_ns = (long long)(us(t0, t2) * 1e3);
    if (trace) {
        std::fprintf(stderr, "[psa] host planning %.1f us, copies + kernels + copy back %.1f us\n", us(t0, t1), us(t1, t2));
        for (const DeviceState& d : ctx->devs)
            if (d.active)
                std::fprintf(stderr, "[psa]   slot %d (cuda:%d): prepare + H2D enqueue %.1f us, launches %.1f us, wait + results %.1f us\n",
                             int(&d - ctx->devs.data()), d.dev, d.st_prepare_ns * 1e-3, d.st_enqueue_ns * 1e-3, d.st_wait_ns * 1e-3);
    }
    return rc;
}

// Pipelined list of problems: lanes (single-device child contexts, each with its own stream, buffers and host thread)
// take problems off a shared counter; see psa_b200.h.
int psa_search_many(psa_context* ctx, psa_problem* problems, int32_t nproblems, int lanes_per_device)
{
    if (!ctx) return PSA_ERR_ARG;
    if (nproblems < 0 || (nproblems > 0 && !problems) || lanes_per_device < 0 || lanes_per_device > kMaxLanes)
        return fail(ctx, PSA_ERR_ARG, "psa_search_many: bad argument");
    if (nproblems == 0) return PSA_OK;
    const int per_dev = lanes_per_device ? lanes_per_device : 2;
    const int ndev = (int)ctx->devs.size();
    const size_t want = (size_t)ndev * per_dev;
    // lane l of device slot g is lanes[g * kMaxLanes + l]; created on first use, kept for the life of the context
    if (ctx->lanes.size() < (size_t)ndev * kMaxLanes) ctx->lanes.resize((size_t)ndev * kMaxLanes, nullptr);
    std::vector<psa_context*> use;
    for (int l = 0; l < per_dev; l++)                                // lane-major: few problems spread over the GPUs first
        for (int g = 0; g < ndev; g++) {
            psa_context*& lane = ctx->lanes[(size_t)g * kMaxLanes + l];
            if (!lane) {
                const int dev = ctx->devs[g].dev;
                const int rc = psa_create(&lane, &dev, 1);
                if (rc) return fail(ctx, rc, "psa_search_many: cannot create a lane on cuda:%d", dev);
            }
            // the caller's knobs apply to every lane
            lane->opt_engine = ctx->opt_engine; lane->opt_rank_planes = ctx->opt_rank_planes; lane->opt_scan_warps = ctx->opt_scan_warps;
            lane->opt_fused_finish = ctx->opt_fused_finish; lane->opt_derive_rank = ctx->opt_derive_rank; lane->opt_pack_queries = ctx->opt_pack_queries;
            lane->opt_zero_copy = ctx->opt_zero_copy; lane->opt_stream_queries = ctx->opt_stream_queries; lane->opt_slices = ctx->opt_slices;
            lane->opt_gather_small = ctx->opt_gather_small;
            lane->opt_sliced_keys = ctx->opt_sliced_keys; lane->opt_batch_mode = ctx->opt_batch_mode; lane->opt_stripe_mode = ctx->opt_stripe_mode;
            lane->opt_single_launch = ctx->opt_single_launch;
            use.push_back(lane);
        }
    const size_t nlanes = std::min(want, (size_t)nproblems);
    std::atomic<int32_t> next{ 0 };
    auto work = [&](psa_context* lane) {
        for (;;) {
            const int32_t k = next.fetch_add(1, std::memory_order_relaxed);
            if (k >= nproblems) return;
            psa_problem& p = problems[k];
            p.status = psa_search_batch(lane, p.weights, p.is_max, p.seq1, p.len1, p.seq2s, p.q_off, p.nq, p.out);
        }
    };
    // lane 0 runs here, the others on threads that stay with the context (waking one costs less than starting one)
    if (ctx->lane_workers.size() < want) ctx->lane_workers.resize(want);
    for (size_t l = 1; l < nlanes; l++) {
        if (!ctx->lane_workers[l]) {
            ctx->lane_workers[l].reset(new Worker());
            ctx->lane_workers[l]->th = std::thread(worker_loop, ctx->lane_workers[l].get());
        }
        Worker& w = *ctx->lane_workers[l];
        std::lock_guard<std::mutex> lk(w.mu);
        psa_context* lane = use[l];
        w.job = [&work, lane]() { work(lane); return 0; };
        w.pending = true;
        w.cv.notify_one();
    }
    work(use[0]);
    for (size_t l = 1; l < nlanes; l++) {
        Worker& w = *ctx->lane_workers[l];
        std::unique_lock<std::mutex> lk(w.mu);
        w.cv.wait(lk, [&w] { return !w.pending; });
    }
    for (int32_t k = 0; k < nproblems; k++)
        if (problems[k].status != PSA_OK) {
            // the message of the lane that failed is gone with the next problem it took: name the problem instead
            return fail(ctx, problems[k].status, "psa_search_many: problem %d failed: %s", (int)k, psa_strerror(problems[k].status));
        }
    return PSA_OK;
}

int psa_search_range(psa_context* ctx, const double weights[4], int is_max, const char* seq1, int64_t len1,
                     const char* seq2, int64_t len2, int64_t first, int64_t last, psa_result* out)
{
    if (first < 0 || last < 0) return ctx ? fail(ctx, PSA_ERR_ARG, "negative offset") : PSA_ERR_ARG;
    const int64_t q_off[2] = { 0, len2 };
    int rc = prepare_common(ctx, weights, is_max, seq1, len1, seq2, q_off, 1, first, last);
    if (rc) return rc;
    return search_prepared(ctx, out);
}

int psa_offset_scores(psa_context* ctx, const double weights[4], int is_max, const char* seq1, int64_t len1, const char* seq2,
                      int64_t len2, int64_t first, int64_t last, double* scores, int32_t* char_offsets, char* letters)
{
    if (!ctx) return PSA_ERR_ARG;
    if (!weights || !seq1 || !seq2 || !scores) return fail(ctx, PSA_ERR_ARG, "null argument");
    if (len2 < 1 || len2 > len1 || len1 > 0x7FFF0000ll || len2 > kExactMaxLen2 || first < 0 || first >= last ||
        last > offsets_of(len1, len2))
        return fail(ctx, PSA_ERR_ARG, "lengths or offset range invalid");
    ctx->prepared = ctx->ran = false;
    DeviceTable T;
    int rc = build_tables(weights, is_max, len2, nullptr, &T);
    if (rc) return fail(ctx, rc, "%s", psa_strerror(rc));
    ctx->table_valid = false;                    // the cached table belongs to the batch entry points
    DeviceState& d = ctx->devs[0];
    PSA_CUDA(ctx, cudaSetDevice(d.dev));
    const int64_t n = last - first;
    if ((rc = ensure_dev(ctx, d.seq1, (size_t)len1 + 64))) return rc;
    if ((rc = ensure_dev(ctx, d.seq2s, (size_t)len2 + 64))) return rc;
    if ((rc = ensure_dev(ctx, d.out, 64))) return rc;
    if ((rc = ensure_dev(ctx, d.partial, (size_t)n * 13 + 64))) return rc;         // scores | char offsets | letters
    if ((rc = ensure_pin(ctx, d.h_out, 64))) return rc;
    double* d_scores = (double*)d.partial.p;
    int32_t* d_coff = (int32_t*)(d_scores + n);
    uint8_t* d_let = (uint8_t*)(d_coff + n);
    BatchGeom G{};
    G.len1 = len1; G.first = first; G.last = last; G.nq = 1; G.uniform_len2 = (int32_t)len2;
    BatchPtrs P{};
    P.seq1 = (const uint8_t*)d.seq1.p; P.seq2s = (const uint8_t*)d.seq2s.p;
    P.cand_count = (int32_t*)d.out.p; P.err_flag = d.h_err; P.run_tag = next_run_tag(d);
    PSA_CUDA(ctx, cudaMemcpyAsync(d.seq1.p, seq1, (size_t)len1, cudaMemcpyHostToDevice, d.stream));
    PSA_CUDA(ctx, cudaMemcpyAsync(d.seq2s.p, seq2, (size_t)len2, cudaMemcpyHostToDevice, d.stream));
    launch_offset_profile(T, G, P, d_scores, d_coff, d_let, d.stream);
    PSA_CUDA(ctx, cudaGetLastError());
    PSA_CUDA(ctx, cudaMemcpyAsync(scores, d_scores, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream));
    if (char_offsets) PSA_CUDA(ctx, cudaMemcpyAsync(char_offsets, d_coff, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, d.stream));
    if (letters) PSA_CUDA(ctx, cudaMemcpyAsync(letters, d_let, (size_t)n, cudaMemcpyDeviceToHost, d.stream));
    PSA_CUDA(ctx, cudaStreamSynchronize(d.stream));
    if (bad_symbol_seen(d)) return fail(ctx, PSA_ERR_ALPHABET, "%s", psa_strerror(PSA_ERR_ALPHABET));
    return PSA_OK;
}

int psa_topk_offsets(psa_context* ctx, const double weights[4], int is_max, const char* seq1, int64_t len1, const char* seq2,
                     int64_t len2, int64_t first, int64_t last, int32_t k, int32_t* offsets, double* scores, int32_t* char_offsets,
                     char* letters, int32_t* found)
{
    if (!ctx) return PSA_ERR_ARG;
    if (!weights || !seq1 || !seq2 || !offsets || !scores || !found || k < 1) return fail(ctx, PSA_ERR_ARG, "null argument or k < 1");
    if (len2 < 1 || len2 > len1 || len1 > 0x7FFF0000ll || len2 > kExactMaxLen2 || first < 0 || first >= last ||
        last > offsets_of(len1, len2))
        return fail(ctx, PSA_ERR_ARG, "lengths or offset range invalid");
    *found = 0;
    ctx->prepared = ctx->ran = false;
    DeviceTable T;
    int rc = build_tables(weights, is_max, len2, nullptr, &T);
    if (rc) return fail(ctx, rc, "%s", psa_strerror(rc));
    ctx->table_valid = false;                    // the cached table belongs to the batch entry points
    DeviceState& d = ctx->devs[0];
    PSA_CUDA(ctx, cudaSetDevice(d.dev));
    const int64_t n = last - first;
    const int kk = int(std::min<int64_t>(k, n));
    if ((rc = ensure_dev(ctx, d.seq1, (size_t)len1 + 64))) return rc;
    if ((rc = ensure_dev(ctx, d.seq2s, (size_t)len2 + 64))) return rc;
    if ((rc = ensure_dev(ctx, d.out, 64))) return rc;
    if ((rc = ensure_dev(ctx, d.partial, (size_t)n * 13 + 64))) return rc;         // scores | char offsets | letters
    if ((rc = ensure_dev(ctx, d.tiles, sizeof(TopkRec) * (size_t)kk + 64))) return rc;
    if ((rc = ensure_pin(ctx, d.h_out, sizeof(TopkRec) * (size_t)kk + 64))) return rc;
    double* d_scores = (double*)d.partial.p;
    int32_t* d_coff = (int32_t*)(d_scores + n);
    uint8_t* d_let = (uint8_t*)(d_coff + n);
    TopkRec* d_top = (TopkRec*)d.tiles.p;
    int32_t* d_found = (int32_t*)((char*)d.tiles.p + sizeof(TopkRec) * (size_t)kk);
    BatchGeom G{};
    G.len1 = len1; G.first = first; G.last = last; G.nq = 1; G.uniform_len2 = (int32_t)len2;
    BatchPtrs P{};
    P.seq1 = (const uint8_t*)d.seq1.p; P.seq2s = (const uint8_t*)d.seq2s.p;
    P.cand_count = (int32_t*)d.out.p; P.err_flag = d.h_err; P.run_tag = next_run_tag(d);
    PSA_CUDA(ctx, cudaMemcpyAsync(d.seq1.p, seq1, (size_t)len1, cudaMemcpyHostToDevice, d.stream));
    PSA_CUDA(ctx, cudaMemcpyAsync(d.seq2s.p, seq2, (size_t)len2, cudaMemcpyHostToDevice, d.stream));
    launch_offset_profile(T, G, P, d_scores, d_coff, d_let, d.stream);
    launch_topk(T.is_max, first, n, kk, d_scores, d_coff, d_let, d_top, d_found, d.stream);
    PSA_CUDA(ctx, cudaGetLastError());
    PSA_CUDA(ctx, cudaMemcpyAsync(d.h_out.p, d_top, sizeof(TopkRec) * (size_t)kk + sizeof(int32_t), cudaMemcpyDeviceToHost, d.stream));
    PSA_CUDA(ctx, cudaStreamSynchronize(d.stream));
    if (bad_symbol_seen(d)) return fail(ctx, PSA_ERR_ALPHABET, "%s", psa_strerror(PSA_ERR_ALPHABET));
    const TopkRec* h = (const TopkRec*)d.h_out.p;
    const int32_t nf = *(const int32_t*)((const char*)d.h_out.p + sizeof(TopkRec) * (size_t)kk);
    for (int r = 0; r < nf; r++) {
        offsets[r] = h[r].offset; scores[r] = h[r].score;
        if (char_offsets) char_offsets[r] = h[r].char_offset;
        if (letters) letters[r] = (char)h[r].letter;
    }
    *found = nf;
    return PSA_OK;
}

int psa_search_batch_mutants(psa_context* ctx, const double weights[4], int is_max, const char* seq1, int64_t len1,
                             const char* seq2s, const int64_t* q_off, int32_t nq, psa_result* out, char* out_mutants)
{
    if (!ctx) return PSA_ERR_ARG;
    if (!out_mutants && nq > 0) return fail(ctx, PSA_ERR_ARG, "null mutant buffer");
    // the records have to stay on the device for the emission kernel: no stores into host memory for this call
    const int saved_zc = ctx->opt_zero_copy;
    ctx->opt_zero_copy = 0;
    int rc = psa_search_batch(ctx, weights, is_max, seq1, len1, seq2s, q_off, nq, out);
    ctx->opt_zero_copy = saved_zc;
    if (rc || nq == 0) return rc;
    if (nq == 1) {
        // one query (possibly split by offset range over the GPUs and merged on the host): one poke, like cpu_funcs.c:96-98
        std::memcpy(out_mutants, seq2s + q_off[0], (size_t)(q_off[1] - q_off[0]));
        if (out[0].mutant.char_offset >= 0 && out[0].mutant.ch) out_mutants[out[0].mutant.char_offset] = out[0].mutant.ch;
        return PSA_OK;
    }
    return for_each_device(ctx, [ctx, q_off, out_mutants](DeviceState& d) {
        if (!d.active) return (int)PSA_OK;
        PSA_CUDA(ctx, cudaSetDevice(d.dev));
        const int64_t byte0 = q_off[d.q_begin], nbytes = q_off[d.q_end] - byte0;
        int r = ensure_dev(ctx, d.mutants, (size_t)nbytes + 64);
        if (r) return r;
        launch_emit_mutants(d.G, d.P, (const QueryRec*)d.out.p, nbytes, (uint8_t*)d.mutants.p, d.sm_count, d.stream);
        PSA_CUDA(ctx, cudaGetLastError());
        PSA_CUDA(ctx, cudaMemcpyAsync(out_mutants + byte0, d.mutants.p, (size_t)nbytes, cudaMemcpyDeviceToHost, d.stream));
        PSA_CUDA(ctx, cudaStreamSynchronize(d.stream));
        return (int)PSA_OK;
    });
}

void* psa_alloc_pinned(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void psa_free_pinned(void* p)
{
    if (p) cudaFreeHost(p);
}

} // extern "C"
