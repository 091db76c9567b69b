// psa_device.cuh -- small device helpers shared by psa_kernels.cu and psa_scan.cu.
#pragma once
#include <cuda_runtime.h>
#include "psa_common.h"

// PSA_CHECK: index / range invariants of the kernels, compiled in only by `make DEBUG=1` (device assert).  The pool
// refuses compute-sanitizer, so a debug run of the GPU tests with these on is the memory-safety net.
#if defined(PSA_DEBUG)
#include <assert.h>
#define PSA_CHECK(cond) assert(cond)
#else
#define PSA_CHECK(cond) ((void)0)
#endif

namespace psa {
namespace {

__device__ __forceinline__ uint32_t symbol_of(uint8_t c)
{
    uint32_t d = uint32_t(c) - uint32_t('A');
    return d < 26u ? d : (c == uint8_t('-') ? uint32_t(kGap) : 0xFFu);
}

// (key, offset) ordering used everywhere: larger key wins, ties go to the lower offset
// (is_swapable, cuda_funcs.cu:290-307, with the score already goal-signed into the key).
struct Cand {
    int64_t key;
    int32_t off;
};

__device__ __forceinline__ bool better(int64_t k2, int32_t o2, int64_t k1, int32_t o1)
{
    return k2 > k1 || (k2 == k1 && o2 < o1);
}

__device__ __forceinline__ Cand warp_best(Cand c)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        int64_t k = __shfl_xor_sync(0xFFFFFFFFu, c.key, d);
        int32_t o = __shfl_xor_sync(0xFFFFFFFFu, c.off, d);
        if (better(k, o, c.key, c.off)) { c.key = k; c.off = o; }
    }
    return c;
}

template <int THREADS>
__device__ __forceinline__ Cand block_best(Cand c, Cand* s_part)
{
    c = warp_best(c);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_part[warp] = c;
    __syncthreads();
    if (warp == 0) {
        Cand v = lane < THREADS / 32 ? s_part[lane] : Cand{ kKeyNone, 0x7FFFFFFF };
        v = warp_best(v);
        if (lane == 0) s_part[0] = v;
    }
    __syncthreads();
    return s_part[0];
}

// Order-preserving map double -> int64 (after -0.0 has been folded into +0.0).
__device__ __forceinline__ int64_t sortable_from_double(double v)
{
    int64_t b = __double_as_longlong(v + 0.0);
    return b ^ ((b >> 63) & 0x7FFFFFFFFFFFFFFFll);
}
__device__ __forceinline__ double double_from_sortable(int64_t k)
{
    return __longlong_as_double(k ^ ((k >> 63) & 0x7FFFFFFFFFFFFFFFll));
}

// query that owns global tile id `tile` (tile_start is a non-decreasing prefix sum, nq+1 entries)
__device__ __forceinline__ int query_of_tile(const int32_t* __restrict__ tile_start, int nq, int tile, int tiles_per_query)
{
    if (tiles_per_query > 0) return tile / tiles_per_query;
    int lo = 0, hi = nq;            // invariant: tile_start[lo] <= tile < tile_start[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tile_start[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// Query geometry: batches of equal-length queries (the common case) carry no per-query arrays at all.
struct QueryGeom {
    int64_t qbeg;     // byte offset of the query in seq2s
    int32_t len2;
    int32_t tile0;    // first tile record of the query
};

__device__ __forceinline__ int32_t first_tile_of(const BatchGeom& G, const int32_t* __restrict__ tile_start, int q)
{
    return G.uniform_len2 > 0 ? q * G.tiles_per_query : tile_start[q];
}

__device__ __forceinline__ QueryGeom query_geom(const BatchGeom& G, const int64_t* __restrict__ qoff,
                                                const int32_t* __restrict__ tile_start, int q)
{
    QueryGeom g;
    if (G.uniform_len2 > 0) {
        g.len2 = G.uniform_len2;
        g.qbeg = int64_t(q) * G.uniform_len2;
        g.tile0 = q * G.tiles_per_query;
    } else {
        g.qbeg = qoff[q];
        g.len2 = int32_t(qoff[q + 1] - g.qbeg);
        g.tile0 = tile_start[q];
    }
    return g;
}

// ---- programmatic dependent launch (sm_90+): a kernel launched with the PDL attribute may start while its
// predecessor in the stream is still running; it must call pdl_wait() before touching anything the predecessor
// writes.  The predecessor calls pdl_launch_dependents() once its blocks no longer need the SM to themselves.
// Both are no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- mbarrier + TMA bulk copy (cp.async.bulk, global -> shared), sm_90+/sm_100a -----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// The same for a warp that can afford to wait: one lane polls, with a pause between polls, so that a ring of waiting warps
// does not keep the shared-memory pipe busy with barrier traffic that the one warp doing the real work has to queue behind.
__device__ __forceinline__ void mbar_wait_patient(uint64_t* bar, uint32_t parity)
{
    if ((threadIdx.x & 31) == 0) {
        for (;;) {
            uint32_t done;
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
            if (done) break;
            __nanosleep(200);
        }
    }
    __syncwarp();
}

// bytes % 16 == 0, both addresses 16-byte aligned; completion is signalled on `bar` (complete_tx)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

} // namespace
} // namespace psa
