// psa_single.cu -- one query, one launch (sm_100a): the gpu_run_program case.
//
// The reference runs ONE (Seq1, Seq2) problem per program run (cpu_funcs.c:353-368 reads a single block of input.txt;
// gpu_run_program cuda_funcs.cu:6-146 launches fill_hashtable_gpu + calc_mutants_scores + reduction x2 for it).  At B200
// speed such a problem is a few microseconds of counting, so the chain k_profile -> k_scan (slices) -> k_combine of
// psa_scan.cu spends most of its ~31 us on three launches, a profile pass over all of Seq1 and bulk copies of windows it
// mostly does not need.  k_single does the whole search in one cooperative launch:
//
//   phase 1  the (1024-offset tile) x (step slice) units go round the blocks.  A block stages just the ~1 KB of Seq1 its
//            unit can touch, builds the unit's STRIPED bit-plane window in shared memory (bit t of lane l = offset
//            tile + l + 32 t, so a step reads one aligned word: the layout of psa_stripe.cu with S = 32), and one warp
//            counts the slice's sign classes and top-rank bits; per-offset partial counts go to global memory in the
//            format of k_scan's slice mode.
//   barrier  all blocks are resident (grid <= SM count): one atomic counter.
//   phase 2  the 256-offset combine tiles go round the blocks (combine_tile, shared with k_combine): slices summed, keys,
//            unresolved offsets settled, one tile record each.
//   phase 3  the block that finishes last picks the winner and writes the result record (finish_body, as k_combine's
//            fused tail does) and resets the two counters for the next launch.
//
// Exact integer order only (the re-score of non-summable weights keeps the k_finish chain).
#include "psa_kernels.cuh"
#include "psa_device.cuh"
#include "psa_finish.cuh"
#include "psa_bitslice.h"
#include "psa_scan_core.cuh"

#include <algorithm>

namespace psa {

#if defined(PSA_SINGLE_TRACE)
// debug build only (make EXTRA=-DPSA_SINGLE_TRACE): SM clock + global timer at the phase boundaries of k_single, block 0
__device__ long long g_single_trace[32];
#define SGL_MARK(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { g_single_trace[k] = clock64(); unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_single_trace[16 + k] = (long long)gt; } } while (0)
extern "C" int psa_debug_single_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_single_trace, sizeof(long long) * 32); }
#define SGL_MARK_ANY(k) do { if (threadIdx.x == 0) { g_single_trace[k] = clock64(); unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_single_trace[16 + k] = (long long)gt; } } while (0)
#else
#define SGL_MARK(k) ((void)0)
#define SGL_MARK_ANY(k) ((void)0)
#endif

namespace {

constexpr int kSingleThreads = 256;             // = kCombineThreads = kFinishThreads

__device__ __forceinline__ void single_class_group(VCounter<5>& A, VCounter<5>& B, VCounter<5>& C, const char* pw, const uint32_t* ro)
{
    uint32_t pa[5], pb[5], pn[5];
#pragma unroll
    for (int s4 = 0; s4 < 32; s4 += 4) {
        const uint4 o4 = *reinterpret_cast<const uint4*>(ro + s4);
        const uint32_t offs[4] = { o4.x, o4.y, o4.z, o4.w };
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint2 x = *reinterpret_cast<const uint2*>(pw + offs[u]);
            vc_feed(A, pa, x.x, s4 + u);
            vc_feed(B, pb, x.y, s4 + u);
            vc_feed(C, pn, x.x & x.y, s4 + u);
        }
    }
}

template <int K>
__global__ void __launch_bounds__(kSingleThreads, 1)
k_single(const __grid_constant__ DeviceTable T, const BatchGeom G, const BatchPtrs P, const SingleGeom SG)
{
    static_assert(kSingleThreads == kCombineThreads && kSingleThreads == kFinishThreads, "one block shape for all phases");
    extern __shared__ __align__(128) unsigned char smem[];
    const int Wn = SG.Wn;                                               // 32 + slice steps
    unsigned char* s_cls = smem;                                        // uint2 [28][Wn]
    unsigned char* s_rnk = s_cls + size_t(kPlaneRows) * Wn * 8;         // uint32 [28][Wn]   (K == 1)
    uint8_t* s_sym = s_rnk + (K > 0 ? size_t(kPlaneRows) * Wn * 4 : 0); // Seq1 symbols the unit can touch
    uint32_t* s_ro = reinterpret_cast<uint32_t*>(s_sym + SG.span);      // per step: (row * Wn + step) * 8
    uint32_t* s_ror = s_ro + SG.slice_steps;                            //           (row * Wn + step) * 4
    __shared__ uint32_t s_col[3][32];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int len2 = G.uniform_len2;
    const int64_t first = G.first, last = G.last, base = tile_base(first);
    const int steps_all = (len2 + 31) & ~31;
    SGL_MARK(0);
    for (int k = tid; k < 3 * 32; k += kSingleThreads) s_col[k >> 5][k & 31] = T.col[k >> 5][k & 31];
    if (blockIdx.x == 0 && tid == 0) P.cand_count[0] = 0;               // statistic: nothing is re-scored in exact order

    // ---- phase 1: partial counts of the (tile, slice) units ----------------------------------------------
    for (int unit = blockIdx.x; unit < SG.units; unit += gridDim.x) {
        const int tile = unit / SG.slices, sl = unit - tile * SG.slices;
        const int64_t tb = base + int64_t(tile) * 1024;                 // offset of lane 0, bit 0
        const int i0 = sl * SG.slice_steps;
        const int ssu = (steps_all - i0) < SG.slice_steps ? (steps_all - i0) : SG.slice_steps;     // multiple of 32
        __syncthreads();                                                // the previous unit is done with shared memory
        {
            // Seq1 positions tb + i0 + [0, span) as symbol indices (31 = past the end / not a symbol: all-zero columns)
            bool bad = false;
            const int64_t p0 = tb + i0;
            for (int k = tid; k < SG.span; k += kSingleThreads) {
                uint32_t c = 31u;
                if (p0 + k < G.len1) {
                    c = symbol_of(P.seq1[p0 + k]);
                    if (c == 0xFFu) { bad = true; c = 31u; }
                }
                s_sym[k] = uint8_t(c);
            }
            // per-step row offsets of the slice; steps at or past len2 read the all-zero row
            for (int k = tid; k < ssu; k += kSingleThreads) {
                uint32_t row = kZeroRow;
                if (i0 + k < len2) {
                    row = symbol_of(P.seq2s[i0 + k]);
                    if (row == 0xFFu) { bad = true; row = 0; }
                }
                s_ro[k] = (row * uint32_t(Wn) + uint32_t(k)) * 8u;
                s_ror[k] = (row * uint32_t(Wn) + uint32_t(k)) * 4u;
            }
            if (bad) report_bad_symbol(P);
        }
        __syncthreads();
        SGL_MARK(1);
        {
            // the striped window of the unit: 32 columns x plane kinds, each a 32x32 bit transpose; the words 32, 64, ... of a
            // column follow by one-bit shifts (psa_stripe.cu)
            constexpr int nkinds = 2 + K;
            for (int task = tid; task < 32 * nkinds; task += kSingleThreads) {
                const int kind = task >> 5, p = task & 31;
                const uint32_t* col = s_col[kind];
                uint32_t m[32];
#pragma unroll
                for (int t = 0; t < 32; t++) m[t] = col[s_sym[p + 32 * t]];
                transpose32(m);
                uint32_t* dst = kind < 2 ? reinterpret_cast<uint32_t*>(s_cls) + kind : reinterpret_cast<uint32_t*>(s_rnk);
                const int wstep = kind < 2 ? 2 : 1;
                for (int k = 0;; k++) {
                    const int word = p + 32 * k;
                    if (word >= 32 + ssu) break;
                    if (k > 0) {
                        const uint32_t c = col[s_sym[word + 31 * 32]];
#pragma unroll
                        for (int r = 0; r < kPlaneRows; r++) m[r] = __funnelshift_r(m[r], c >> r, 1);
                    }
#pragma unroll
                    for (int r = 0; r < kPlaneRows; r++) dst[(size_t(r) * Wn + word) * wstep] = m[r];
                }
            }
        }
        __syncthreads();
        SGL_MARK(2);
        if (warp == 0) {
            // bits of this lane that are offsets of the range: n = tb + lane + 32 t in [first, last)
            uint32_t vmask = 0u;
#pragma unroll
            for (int t = 0; t < 32; t++) {
                const int64_t n = tb + lane + 32 * t;
                vmask |= uint32_t(n >= first && n < last) << t;
            }
            uint32_t racc = ~vmask;
            const int groups = ssu >> 5;
            if (K > 0) {
                const char* pr = reinterpret_cast<const char*>(s_rnk) + size_t(lane) * 4;
                for (int g = 0; g < groups; g++) {
#pragma unroll
                    for (int s4 = 0; s4 < 32; s4 += 4) {
                        const uint4 o4 = *reinterpret_cast<const uint4*>(s_ror + g * 32 + s4);
                        racc |= *reinterpret_cast<const uint32_t*>(pr + o4.x) | *reinterpret_cast<const uint32_t*>(pr + o4.y);
                        racc |= *reinterpret_cast<const uint32_t*>(pr + o4.z) | *reinterpret_cast<const uint32_t*>(pr + o4.w);
                    }
                    if (__all_sync(0xFFFFFFFFu, racc == 0xFFFFFFFFu)) break;
                }
            }
            VCounter<5> A, B, C;                                        // slices have at most 1023 steps
            A.clear(); B.clear(); C.clear();
            const char* pw = reinterpret_cast<const char*>(s_cls) + size_t(lane) * 8;
            for (int g = 0; g < groups; g++) single_class_group(A, B, C, pw, s_ro + g * 32);
            // per offset: {N(b0) | N(b1) << 16, N(b0&b1) | top-rank bit << 16} -- k_scan's slice format, read by combine_tile
            uint32_t m[32], m2[32];
#pragma unroll
            for (int k = 0; k < 32; k++) { m[k] = 0; m2[k] = 0; }
#pragma unroll
            for (int k = 0; k < 10; k++) { m[k] = A.plane(k); m[16 + k] = B.plane(k); m2[k] = C.plane(k); }
            m2[16] = K > 0 ? (racc & vmask) : 0u;
            transpose32(m);
            transpose32(m2);
            uint2* dst = P.partial + int64_t(sl) * P.partial_stride + (int64_t(tile) * 1024 + lane);
#pragma unroll
            for (int t = 0; t < 32; t++)
                if ((vmask >> t) & 1u) dst[32 * t] = make_uint2(m[t], m2[t]);
        }
    }

    // ---- barrier: every unit's partial counts are in global memory --------------------------------------
    __syncthreads();
    SGL_MARK(3);
    if (tid == 0) {
        __threadfence();
        atomicAdd(P.sync, 1);
        while (*reinterpret_cast<volatile int32_t*>(P.sync) < int(gridDim.x)) {}
        __threadfence();
    }
    __syncthreads();

    SGL_MARK(4);
    // ---- phase 2: combine tiles round the blocks --------------------------------------------------------
    for (int tile = blockIdx.x; tile < G.total_tiles; tile += gridDim.x) {
        __syncthreads();
        combine_tile<K, false>(T, G, P, SG.slices, tile);
    }

    // ---- phase 3: the last block finishes the query -----------------------------------------------------
    __syncthreads();
    SGL_MARK(5);
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(P.sync + 1, 1) == int(gridDim.x) - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) { P.sync[0] = 0; P.sync[1] = 0; }                     // every block is past both counters: ready for the next launch
    SGL_MARK_ANY(6);
    finish_body<kFinishWarps, false>(T, G, P, 1, 0);
    SGL_MARK_ANY(7);
}

} // namespace

// Shape of a one-launch single-query search over offsets [first, last) of a len2-symbol query, or ok = 0.
SingleGeom single_plan(int64_t len1, int64_t len2, int64_t first, int64_t last, int sm_count)
{
    SingleGeom g{};
    if (len2 < 1 || len2 > kScanMaxLen2 || last <= first || sm_count < 1) return g;
    const int64_t span_off = last - tile_base(first);
    const int64_t tiles = (span_off + 1023) / 1024;
    const int64_t steps_all = (len2 + 31) & ~int64_t(31);
    if (tiles > 4 * sm_count) return g;                                 // that many warp-tiles fill the GPU the ordinary way
    // slices: enough units for every SM to have one, at least 64 steps each, at most what the window's shared memory allows
    int64_t slices = std::max<int64_t>(1, sm_count / tiles);
    int64_t ss = ((steps_all + slices - 1) / slices + 31) & ~int64_t(31);
    ss = std::max<int64_t>(ss, std::min<int64_t>(64, steps_all));
    ss = std::min<int64_t>(ss, 384);                                    // window of 28 x (32 + 384) words x 12 bytes = 140 KB
    slices = (steps_all + ss - 1) / ss;
    g.slice_steps = int(ss);
    g.slices = int(slices);
    g.tiles = int(tiles);
    g.units = int(tiles * slices);
    g.Wn = 32 + int(ss);
    g.span = (int(ss) + 32 + 31 * 32 + 32 + 15) & ~15;
    g.blocks = int(std::min<int64_t>(sm_count, std::max<int64_t>(g.units, (span_off + 255) / 256)));
    g.smem = size_t(kPlaneRows) * g.Wn * 12 + size_t(g.span) + size_t(ss) * 8;
    if (g.smem > 160 * 1024) return g;
    g.ok = 1;
    return g;
}

void launch_single(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int rank_planes, const SingleGeom& SG, cudaStream_t stream)
{
    auto go = [&](auto kernel, bool (&done)[64]) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (!(dev >= 0 && dev < 64 && done[dev])) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
        // cooperative: the grid barrier needs every block resident, also when other streams share the GPU
        void* args[] = { (void*)&T, (void*)&G, (void*)&P, (void*)&SG };
        // (measured: a cooperative launch costs no more than an ordinary one here -- 25.0 vs 24.7 us per config-1 step)
        cudaLaunchCooperativeKernel((const void*)kernel, dim3(SG.blocks), dim3(kSingleThreads), args, SG.smem, stream);
    };
    static bool done0[64], done1[64];
    if (rank_planes == 0) go(k_single<0>, done0);
    else go(k_single<1>, done1);
}

} // namespace psa
