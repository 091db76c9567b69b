// psa_scan.cu -- the hot kernels of the scan engine (sm_100a): Seq1 bit-plane profile + bit-sliced scan.
//
// What the reference does per pair per offset (cuda_funcs.cu:171-198: two table signs, a substitution
// search, two weights, a global-memory FP64 read-modify-write) becomes here:
//
//   k_profile  once per batch: for every Seq2 symbol a (27 rows) and every Seq1 position j, three facts
//              about the pair (a, Seq1[j]) as bit planes over j:
//                 b0, b1  the 2-bit sign class ('*'=0 ':'=1 '.'=2 '_'=3)
//                 r_p     "the best substitution here has the p-th best rank" for the top K ranks
//   k_scan     a lane owns 32 consecutive offsets as the 32 bits of a register; a warp owns 1024.
//              Step i of the alignment needs, for all 32 offsets at once, bits [n+i, n+i+32) of row
//              Seq2[i] -- two aligned shared-memory words and one funnel shift per plane.  The sign
//              counts N(b0), N(b1), N(b0&b1) are kept as vertical (bit-sliced) counters updated with
//              carry-save adders: ~2 LOP3 per plane per step for 32 offsets, i.e. ~0.3 integer
//              lane-ops per pair evaluation instead of the >= 2 of a scalar formulation.  The best rank
//              is the OR of the rank planes.  Counts are exact integers; the score key is formed once
//              per offset in the epilogue (after a 32x32 bit transpose) as an int64.
//
// Offsets whose best rank is below the K tracked planes are "unresolved": in exact mode the warp that owns them
// settles the few whose upper bound could still win by walking their alignment once (settle_unresolved); in
// re-score mode the per-word upper estimates go to the finish step (finish_body, psa_finish.cuh), which re-scores
// every candidate word in the reference's summation order.
#include "psa_kernels.cuh"
#include "psa_device.cuh"
#include "psa_finish.cuh"
#include "psa_bitslice.h"
#include "psa_scan_core.cuh"

#include <algorithm>
#include <cmath>
#include <type_traits>

namespace psa {

namespace {

constexpr int kProfileThreads = 64;      // few, fat threads: a 3000-letter Seq1 is only ~300 words
// bytes between consecutive words' rows in a staged rank window: K = 1 keeps the class window's 8-byte pitch so the
// per-step row offsets (row * nwords * 8) address both windows without a shift
__host__ __device__ constexpr int rank_pitch(int K) { return K == 1 ? 8 : 4 * K; }
constexpr int kScanChunkMax = 1024;     // alignment steps staged per shared-memory window

// -------------------------------------------------------------------------------------------------
// k_profile: bit planes of Seq1 against every Seq2 symbol.  One thread per 32-position word: each position
// contributes, per plane kind, a 28-bit column (one bit per row symbol) looked up by its Seq1 symbol; a
// 32x32 bit transpose then turns 32 columns into the 28 row words.  (2+K) transposes per word instead of
// 28 x 32 table lookups.
// -------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kProfileThreads)
k_profile(const __grid_constant__ DeviceTable T, const BatchGeom G, const BatchPtrs P)
{
    __shared__ uint32_t s_col[2 + (K > 0 ? K : 1)][32];           // [plane kind][Seq1 symbol] -> bit r = row symbol r
    pdl_launch_dependents();                                        // the scan may set itself up while we run
    if (blockIdx.x == 0 && threadIdx.x == 0) {                      // first kernel of the chain:
        P.cand_count[0] = 0;                                        //   only the finish step adds to the re-score counter
        P.cand_count[2] = 0;                                        //   k_combine's "last block" ticket
    }
    for (int k = threadIdx.x; k < (2 + K) * 32; k += kProfileThreads)       // columns resolved on the host (psa_table.cpp)
        s_col[k >> 5][k & 31] = T.col[k >> 5][k & 31];
    __syncthreads();
    for (int64_t w = int64_t(blockIdx.x) * kProfileThreads + threadIdx.x; w < P.plane_words;
         w += int64_t(gridDim.x) * kProfileThreads) {
        const int64_t base = w * 32;
        uint8_t sym[32];
        int n = 0;
        if (base < G.len1) {
            // 32 bytes of Seq1; the buffer is padded so the vector loads stay inside the allocation
            const uint4 v0 = *reinterpret_cast<const uint4*>(P.seq1 + base);
            const uint4 v1 = *reinterpret_cast<const uint4*>(P.seq1 + base + 16);
            const uint32_t words[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
            n = (G.len1 - base) < 32 ? int(G.len1 - base) : 32;
#pragma unroll
            for (int t = 0; t < 32; t++) {
                uint32_t c = symbol_of(uint8_t(words[t >> 2] >> (8 * (t & 3))));
                if (t < n && c == 0xFFu) report_bad_symbol(P);
                sym[t] = uint8_t(c == 0xFFu ? 0u : c);
            }
        }
        uint32_t m[32], b0rows[kPlaneRows];
#pragma unroll
        for (int kind = 0; kind < 2 + K; kind++) {
#pragma unroll
            for (int t = 0; t < 32; t++) m[t] = t < n ? s_col[kind][sym[t]] : 0u;
            transpose32(m);
            if (kind == 0) {
#pragma unroll
                for (int r = 0; r < kPlaneRows; r++) b0rows[r] = m[r];
            } else if (kind == 1) {
#pragma unroll
                for (int r = 0; r < kPlaneRows; r++) P.cls_planes[int64_t(r) * P.plane_words + w] = make_uint2(b0rows[r], m[r]);
            } else {
#pragma unroll
                for (int r = 0; r < kPlaneRows; r++) P.rank_planes[(int64_t(r) * P.plane_words + w) * K + (kind - 2)] = m[r];
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// k_scan (long mode: one query tile per block, len2 of any size up to 32767)
// NB : counter planes (len2 < 2^NB), K : rank planes tracked, KEY32 : keys fit 26 bits (packed compare)
// block = warps x 32 threads, tile = warps x 1024 offsets, one block per tile.
//
// Two passes over the alignment share one shared-memory window [28 rows][nwords] (+ [chunk] row offsets):
//   pass R  rank planes ([K] uint32 per word): OR-accumulate until every offset of the block has met the
//           top rank (a few dozen steps on long queries) -- then the pass stops;
//   pass C  class planes (uint2 per word): the three vertical counters over all len2 steps.
// Windows are filled by TMA bulk copies (one per plane row) completing on an mbarrier while the block
// turns its slice of Seq2 into row offsets.
// -------------------------------------------------------------------------------------------------
// SLICE = true (single query, few warp-tiles): blockIdx.y selects a slice of the alignment steps
// [slice * slice_len, ...); the block leaves its partial counts and rank bits per offset in P.partial and
// k_combine adds the slices up -- this multiplies the warps in flight when one query cannot fill the GPU.
// DR = true (K = 1 only): the top-rank bit is derived from the class counts (derive_top_rank), pass R is skipped.
template <int NB, int K, bool BS, bool SLICE, bool DR>
__global__ void __launch_bounds__(128, NB <= 10 ? 6 : 4)
k_scan(const __grid_constant__ DeviceTable T, const BatchGeom G, const BatchPtrs P, const int nwords, const int chunk,
       const int key_planes, const int64_t key_bias, const int slice_len, const int fused_finish)
{
    constexpr int NUP = NB - 5;
    constexpr int kEntry = (4 * K > 8) ? 4 * K : 8;                 // bytes per window word (max of the two passes)
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t* s_ro = reinterpret_cast<uint32_t*>(smem + size_t(kPlaneRows) * nwords * kEntry);
    __shared__ Cand s_res[4];
    __shared__ int64_t s_top[4];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, warps = nthreads >> 5;
    const int tile_id = blockIdx.x;
    const int q = query_of_tile(P.tile_start, G.nq, tile_id, G.tiles_per_query);
    const QueryGeom qg = query_geom(G, P.qoff, P.tile_start, q);
    const int t = tile_id - qg.tile0;
    const int64_t qbeg = qg.qbeg;
    const int len2 = qg.len2;
    const int64_t first = G.last >= 0 ? G.first : 0;
    const int64_t last = G.last >= 0 ? G.last : G.len1 - len2 + 1;
    const int64_t tb = tile_base(first) + int64_t(t) * G.tile;      // multiple of 128
    const int64_t ln0 = tb + warp * 1024 + lane * 32;               // this lane's first offset

    if (tid == 0) mbar_init(&s_bar, 1);
    const uint32_t vmask = valid_mask(ln0, first, last);
    const bool warp_active = __any_sync(0xFFFFFFFFu, vmask != 0);

    uint32_t racc[K > 0 ? K : 1];
#pragma unroll
    for (int k = 0; k < (K > 0 ? K : 1); k++) racc[k] = 0;
    racc[0] = ~vmask;                       // offsets outside the range count as saturated
    uint32_t parity = 0;
    pdl_wait();                                                     // the bit planes come from k_profile
    const int steps_all = (len2 + 31) & ~31;
    const int step_begin = SLICE ? int(blockIdx.y) * slice_len : 0;                    // multiple of 128
    const int steps_total = SLICE ? ((step_begin + slice_len) < steps_all ? (step_begin + slice_len) : steps_all) : steps_all;

    // Seq2 slice -> per-step row offsets (row * nwords * 8 bytes); padding steps use the all-zero row
    auto fill_row_offsets = [&](int c0, int cl) {
        if (fill_rows<8>(s_ro, P.seq2s + qbeg, c0, cl, len2, uint32_t(nwords) * 8u, tid, nthreads)) report_bad_symbol(P);
    };

    // ---- pass R: best rank per offset -----------------------------------------------------------------
    if (K > 0 && !DR) {
        bool rank_on = warp_active;
        for (int c0 = step_begin; c0 < steps_total; c0 += chunk) {
            const int cl = (steps_total - c0) < chunk ? (steps_total - c0) : chunk;
            const int need = round_up4(warps * 32 + (cl >> 5));
            PSA_CHECK(need <= nwords && cl <= chunk && ((tb + c0) >> 5) + need <= P.plane_words && ((tb + c0) & 127) == 0);
            __syncthreads();                                        // previous window consumed / barrier initialised
            if (warp == 0) {
                const int64_t g0 = (tb + c0) >> 5;                  // multiple of 4: every row starts on 16 bytes
                if (lane == 0) mbar_arrive_expect_tx(&s_bar, uint32_t(kPlaneRows) * uint32_t(need) * uint32_t(4 * K));
                __syncwarp();
                if (lane < kPlaneRows)
                    tma_load_1d(smem + size_t(lane) * nwords * rank_pitch(K), P.rank_planes + (int64_t(lane) * P.plane_words + g0) * K,
                                uint32_t(need) * 4u * K, &s_bar);
            }
            fill_row_offsets(c0, cl);
            __syncthreads();
            mbar_wait(&s_bar, parity);
            parity ^= 1u;
            if (rank_on) {
                const int groups = cl >> 5;
                for (int g = 0; g < groups; g++) {
                    rank_group<K>(racc, reinterpret_cast<const char*>(smem) + size_t(warp * 32 + lane + g) * 4 * K, s_ro + g * 32);
                    // every offset of the warp carries the top rank: the lower planes and later steps are moot
                    if (__all_sync(0xFFFFFFFFu, racc[0] == 0xFFFFFFFFu)) { rank_on = false; break; }
                }
            }
            if (__syncthreads_and(!rank_on)) break;                 // the whole block is saturated
        }
    }

    // ---- pass C: sign-class counts --------------------------------------------------------------------
    VCounter<NUP> A, B, C;
    A.clear(); B.clear(); C.clear();
    for (int c0 = step_begin; c0 < steps_total; c0 += chunk) {
        const int cl = (steps_total - c0) < chunk ? (steps_total - c0) : chunk;
        const int need = round_up4(warps * 32 + (cl >> 5));
        PSA_CHECK(need <= nwords && cl <= chunk && ((tb + c0) >> 5) + need <= P.plane_words && ((tb + c0) & 127) == 0);
        __syncthreads();
        if (warp == 0) {
            const int64_t g0 = (tb + c0) >> 5;
            if (lane == 0) mbar_arrive_expect_tx(&s_bar, uint32_t(kPlaneRows) * uint32_t(need) * 8u);
            __syncwarp();
            if (lane < kPlaneRows)
                tma_load_1d(smem + size_t(lane) * nwords * 8, P.cls_planes + int64_t(lane) * P.plane_words + g0,
                            uint32_t(need) * 8u, &s_bar);
        }
        fill_row_offsets(c0, cl);
        __syncthreads();
        mbar_wait(&s_bar, parity);
        parity ^= 1u;
        if (warp_active) {
            const int groups = cl >> 5;
            for (int g = 0; g < groups; g++) {
                const char* pw = reinterpret_cast<const char*>(smem) + size_t(warp * 32 + lane + g) * 8;
                class_group<NUP>(A, B, C, pw, s_ro + g * 32);
            }
        }
    }
    if (DR) racc[0] |= derive_top_rank<NB, NUP>(T, A, B, C);

    // ---- epilogue -------------------------------------------------------------------------------------
    pdl_launch_dependents();
    if (SLICE) {
        // partial counts of this slice, one uint2 per offset: {N(b0) | N(b1) << 16, N(b0&b1) | rank bits << 16}
        if (warp_active) {
            uint32_t m[32], m2[32];
#pragma unroll
            for (int k = 0; k < 32; k++) { m[k] = 0; m2[k] = 0; }
#pragma unroll
            for (int k = 0; k < NB; k++) { m[k] = A.plane(k); m[16 + k] = B.plane(k); m2[k] = C.plane(k); }
#pragma unroll
            for (int k = 0; k < K; k++) m2[16 + k] = racc[k];
            transpose32(m);
            transpose32(m2);
            PSA_CHECK(ln0 - tile_base(first) >= 0 && ln0 - tile_base(first) + 32 <= P.partial_stride);
            uint2* dst = P.partial + int64_t(blockIdx.y) * P.partial_stride + (ln0 - tile_base(first));
#pragma unroll
            for (int tt = 0; tt < 32; tt++)
                if ((vmask >> tt) & 1u) dst[tt] = make_uint2(m[tt], m2[tt]);
        }
        return;
    }
    Cand mine{ kKeyNone, 0x7FFFFFFF }, ub{ kKeyNone, 0x7FFFFFFF };
    uint32_t umask = 0;
    typename std::conditional<BS, SlicedKeys<NB, K>, OffsetKeys<NB, K, false>>::type keys;
    if (warp_active) {
        keys.build(T, len2, A, B, C, racc, key_planes, key_bias);
        umask = keys.scan(vmask, ln0, mine, ub);
    }
    if (T.exact) {
        Cand wbest{ kKeyNone, 0x7FFFFFFF };
        if (warp_active) wbest = settle_unresolved(T, P, keys, mine, ub, umask, ln0, qbeg, len2);
        if (lane == 0) { s_res[warp] = wbest; s_top[warp] = kKeyNone; }
    } else {
        // Re-score mode: keys only pre-select; record an upper estimate per 32-offset word for k_finish.
        const int64_t top = mine.key > ub.key ? mine.key : ub.key;
        P.lane_keys[int64_t(tile_id) * (G.tile >> 5) + warp * 32 + lane] = top;
        const Cand wbest = warp_best(mine);
        int64_t wtop = top;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int64_t o = __shfl_xor_sync(0xFFFFFFFFu, wtop, d);
            wtop = o > wtop ? o : wtop;
        }
        if (lane == 0) { s_res[warp] = wbest; s_top[warp] = wtop; }
    }
    __syncthreads();
    Cand r = s_res[0];
    int64_t top = s_top[0];
    for (int w = 1; w < warps; w++) {
        take(r, s_res[w].key, s_res[w].off);
        top = s_top[w] > top ? s_top[w] : top;
    }
    PSA_CHECK(tile_id < G.total_tiles && q < G.nq);
    if (!fused_finish) {
        if (tid == 0) {
            TileRec rec;
            rec.key = r.key; rec.offset = r.off;
            rec.ub_key = top; rec.ub_offset = 0x7FFFFFFF;
            rec.score = 0.0; rec.flags = 0; rec.pad = 0;
            P.tiles[tile_id] = rec;
        }
        return;
    }
    // Fused finish (exact order, one tile per query): this block already holds the query's winner, so it also
    // does what k_finish would -- sign counts, first position carrying the best rank, replacement letter, score --
    // and the extra launch disappears.
    __shared__ unsigned long long s_pos;
    __shared__ int s_cnt[4];
    if (tid == 0) s_pos = 0ull;
    if (tid < 4) s_cnt[tid] = 0;
    __syncthreads();
    QueryRec out;
    out.score = T.is_max ? -INFINITY : INFINITY;
    out.offset = -1; out.char_offset = -1; out.ch = 0; out.rank = 0;
    out.counts[0] = out.counts[1] = out.counts[2] = out.counts[3] = 0;
    if (r.key == kKeyNone) {
        if (tid == 0) P.out[q] = out;
        return;
    }
    const uint8_t* a = P.seq1 + r.off;
    const uint8_t* b = P.seq2s + qbeg;
    int cnt[4] = { 0, 0, 0, 0 };
    unsigned long long pos = 0ull;        // (rank << 32) | ~i  -> max = best rank, then lowest i
    walk_alignment<6>(a, b, P.code_table, len2, tid, nthreads, cnt, pos);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, pos, d);
        pos = o > pos ? o : pos;
#pragma unroll
        for (int c = 0; c < 4; c++) cnt[c] += __shfl_xor_sync(0xFFFFFFFFu, cnt[c], d);
    }
    if (lane == 0) {
        atomicMax(&s_pos, pos);
#pragma unroll
        for (int c = 0; c < 4; c++) atomicAdd(&s_cnt[c], cnt[c]);
    }
    __syncthreads();
    if (tid == 0) {
        const int rank = int(s_pos >> 32);
        const int i = int(~uint32_t(s_pos));
        uint32_t c1 = symbol_of(a[i]), c2 = symbol_of(b[i]);
        if (c1 == 0xFFu || c2 == 0xFFu) { c1 = 0; c2 = 0; }
        out.offset = r.off;
        out.char_offset = i;
        out.ch = T.sub[c2][c1];
        out.rank = rank;
        double sc = 0.0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            out.counts[c] = s_cnt[c];
            sc = __dadd_rn(sc, __dmul_rn(double(s_cnt[c]), T.wcls[c]));        // exact (psa_table.cpp), same as k_finish
        }
        out.score = __dadd_rn(__dadd_rn(sc, T.wdiff[rank]), 0.0);
        if (rank <= 0) { out.offset = -1; out.char_offset = -1; out.ch = 0; out.score = T.is_max ? -INFINITY : INFINITY; }
        P.out[q] = out;
    }
}

// -------------------------------------------------------------------------------------------------
// k_combine (slice mode): one thread per offset adds the slices' partial counts, forms the key, and the
// block (256 offsets = one tile record for k_finish) reduces to its best.  Offsets that met no tracked rank
// plane are settled here when the order must be exact: those whose bound could beat the block's best walk
// the alignment for their true best rank (rare, and only for the few that matter).
// -------------------------------------------------------------------------------------------------
// FUSE (small grids only): the block that finishes last also runs the finish step -- its registers would cost the
// many-block case its occupancy, and there one more launch does not matter.
template <int K, bool FUSE>
__global__ void __launch_bounds__(kCombineThreads)
k_combine(const __grid_constant__ DeviceTable T, const BatchGeom G, const BatchPtrs P, const int slices)
{
    pdl_wait();                                                     // partial counts come from the scan slices
    combine_tile<K, true>(T, G, P, slices, int(blockIdx.x));
    if (FUSE) {
        // the last block to get here has every tile record (and lane key) of the query in front of it: the ticket is taken
        // after this block's stores + __threadfence, and finish_body reads the records with ld.cg (never from L1)
        __shared__ int s_last;
        const int tid = threadIdx.x;
        __syncthreads();                                        // this block's lane keys and record are written
        if (tid == 0) {
            __threadfence();
            s_last = atomicAdd(P.cand_count + 2, 1) == int(gridDim.x) - 1;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        finish_body<kFinishWarps, false>(T, G, P, 1, 0);
    }
}

// -------------------------------------------------------------------------------------------------
// k_scan_batch (batch mode: every query fits one window, len2 <= 1023)
// The bit-plane window of a tile depends only on the offsets, not on the query, so a block stages the
// windows of ONE 1024-offset tile once (class + rank planes, TMA) and then each of its warps walks its own
// queries against them: no block-wide barrier after the staging, warps drift apart freely, and the L2 -> SM
// traffic per pair evaluation drops by the number of queries per block.
//   grid = (query groups, offset tiles); warp w of block (g, t) handles queries g*qpb + w, + warps, ...
//   shared memory: [28][nwords] uint2 | [28][nwords][K] uint32 | per warp [chunk] uint32 row offsets
// One TileRec per (query, tile) is written by the warp that computed it.
// -------------------------------------------------------------------------------------------------
template <int NB, int K, bool BS, bool DR>
__global__ void __launch_bounds__(128, NB <= 10 ? 6 : 4)
k_scan_batch(const __grid_constant__ DeviceTable T, const BatchGeom G, const BatchPtrs P, const int nwords, const int chunk,
             const int queries_per_block, const int key_planes, const int64_t key_bias)
{
    constexpr int NUP = NB - 5;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_cls = smem;
    unsigned char* s_rnk = smem + size_t(kPlaneRows) * nwords * 8;
    uint32_t* s_ro_all = reinterpret_cast<uint32_t*>(s_rnk + size_t(kPlaneRows) * nwords * rank_pitch(K));
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, warps = blockDim.x >> 5;
    uint32_t* s_ro = s_ro_all + size_t(warp) * chunk;
    const int tile = blockIdx.y;
    const int64_t tb = int64_t(tile) * 1024;                        // batch mode always scans from offset 0
    const int64_t ln0 = tb + lane * 32;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_arrive_expect_tx(&s_bar, uint32_t(kPlaneRows) * uint32_t(nwords) * uint32_t(8 + (DR ? 0 : 4 * K)));
    }
    __syncthreads();
    pdl_wait();                                                     // the bit planes come from k_profile
    PSA_CHECK((tb >> 5) + nwords <= P.plane_words && (nwords & 3) == 0);
    if (warp == 0 && lane < kPlaneRows) {
        const int64_t g0 = tb >> 5;
        tma_load_1d(s_cls + size_t(lane) * nwords * 8, P.cls_planes + int64_t(lane) * P.plane_words + g0, uint32_t(nwords) * 8u, &s_bar);
        if (K > 0 && !DR)
            tma_load_1d(s_rnk + size_t(lane) * nwords * rank_pitch(K), P.rank_planes + (int64_t(lane) * P.plane_words + g0) * K,
                        uint32_t(nwords) * 4u * K, &s_bar);
    }
    bool staged = false;

    const int q_begin = blockIdx.x * queries_per_block;
    const int q_end = (q_begin + queries_per_block) < G.nq ? (q_begin + queries_per_block) : G.nq;
    // The first 64 bytes of a warp's NEXT query are fetched while it works on the current one, so the global round
    // trip of the row-offset set-up is off the critical path (most of a short query fits those two loads).
    auto fetch_head = [&](int qq, uint8_t& h0, uint8_t& h1) {
        h0 = h1 = uint8_t('A');
        if (qq < q_end) {
            const QueryGeom g = query_geom(G, P.qoff, P.tile_start, qq);
            if (lane < g.len2) h0 = P.seq2s[g.qbeg + lane];
            if (lane + 32 < g.len2) h1 = P.seq2s[g.qbeg + lane + 32];
        }
    };
    uint8_t head0, head1;
    fetch_head(q_begin + warp, head0, head1);
    for (int q = q_begin + warp; q < q_end; q += warps) {
        const QueryGeom qg = query_geom(G, P.qoff, P.tile_start, q);
        const int64_t qbeg = qg.qbeg;
        const int len2 = qg.len2;
        const int64_t last = G.len1 - len2 + 1;
        const uint8_t cur0 = head0, cur1 = head1;
        fetch_head(q + warps, head0, head1);
        if (tb >= last) continue;                                   // this query does not reach the tile
        const uint32_t vmask = valid_mask(ln0, 0, last);
        const int steps_total = (len2 + 31) & ~31;
        PSA_CHECK(steps_total <= chunk && 32 + (steps_total >> 5) <= nwords && qg.tile0 + tile < G.total_tiles);
        __syncwarp();                                               // previous query's row offsets are no longer read
        // short queries, one warp: the plain loop beats batched predicated loads here (config 5: 2.07 vs 2.31 ms)
        for (int s = lane; s < steps_total; s += 32) {
            uint32_t row = kZeroRow;
            if (s < len2) {
                row = symbol_of(s < 32 ? cur0 : s < 64 ? cur1 : P.seq2s[qbeg + s]);
                if (row == 0xFFu) { report_bad_symbol(P); row = 0; }
            }
            s_ro[s] = row * uint32_t(nwords) * 8u;
        }
        __syncwarp();
        if (!staged) { mbar_wait(&s_bar, 0); staged = true; }

        uint32_t racc[K > 0 ? K : 1];
#pragma unroll
        for (int k = 0; k < (K > 0 ? K : 1); k++) racc[k] = 0;
        racc[0] = ~vmask;
        const int groups = steps_total >> 5;
        if (K > 0 && !DR) {
            for (int g = 0; g < groups; g++) {
                rank_group<K>(racc, reinterpret_cast<const char*>(s_rnk) + size_t(lane + g) * 4 * K, s_ro + g * 32);
                if (__all_sync(0xFFFFFFFFu, racc[0] == 0xFFFFFFFFu)) break;
            }
        }
        VCounter<NUP> A, B, C;
        A.clear(); B.clear(); C.clear();
        for (int g = 0; g < groups; g++)
            class_group<NUP>(A, B, C, reinterpret_cast<const char*>(s_cls) + size_t(lane + g) * 8, s_ro + g * 32);
        if (DR) racc[0] |= derive_top_rank<NB, NUP>(T, A, B, C);

        Cand mine{ kKeyNone, 0x7FFFFFFF }, ub{ kKeyNone, 0x7FFFFFFF };
        typename std::conditional<BS, SlicedKeys<NB, K>, OffsetKeys<NB, K, false>>::type keys;
        keys.build(T, len2, A, B, C, racc, key_planes, key_bias);
        const uint32_t umask = keys.scan(vmask, ln0, mine, ub);
        const int rec_id = qg.tile0 + tile;
        TileRec rec;
        rec.score = 0.0; rec.flags = 0; rec.pad = 0; rec.ub_offset = 0x7FFFFFFF;
        if (T.exact) {
            const Cand wbest = settle_unresolved(T, P, keys, mine, ub, umask, ln0, qbeg, len2);
            rec.key = wbest.key; rec.offset = wbest.off; rec.ub_key = kKeyNone;
        } else {
            const int64_t top = mine.key > ub.key ? mine.key : ub.key;
            P.lane_keys[int64_t(rec_id) * 32 + lane] = top;
            const Cand wbest = warp_best(mine);
            int64_t wtop = top;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const int64_t o = __shfl_xor_sync(0xFFFFFFFFu, wtop, d);
                wtop = o > wtop ? o : wtop;
            }
            rec.key = wbest.key; rec.offset = wbest.off; rec.ub_key = wtop;
        }
        if (lane == 0) P.tiles[rec_id] = rec;
    }
    if (!staged) mbar_wait(&s_bar, 0);      // never leave with a bulk copy in flight
    pdl_launch_dependents();
}

// -------------------------------------------------------------------------------------------------
// k_scan_packed (packed mode: equal-length queries whose whole offset range fits one window, len2 <= 1023)
//
// A query with `noff` offsets needs L = ceil(noff / 32) lanes; k_scan gives it whole warps, so e.g. 2501 offsets
// (L = 79) occupy 3 warps = 96 lanes and every sixth lane-instruction is idle.  Here a block owns Q whole queries and
// lays their lanes end to end: thread t works on query t / L, offset word t % L, and the block is ceil(Q L / 32) warps
// (config 3: Q = 2, 5 warps, 158 of 160 lanes busy).  All queries of a batch read the same Seq1 window, which is staged
// once per block; only the per-step row offsets differ per query, and a warp that straddles two queries simply reads
// two rows (the 64-bit loads are served per half-warp anyway).  Per-query results come from a segmented reduction:
// each warp reduces (and, in exact mode, settles) one query segment at a time, a thread per query merges the warps.
// -------------------------------------------------------------------------------------------------

template <int NB, int K, bool BS, bool DR>
// 96 registers: four 5-warp blocks (config 3's shape) still fit an SM, and it measured 2.7 % faster than the 80 that
// __launch_bounds__(256, 3) allows
__global__ void __maxnreg__(96)
k_scan_packed(const __grid_constant__ DeviceTable T, const BatchGeom G, const BatchPtrs P, const int nwords, const int steps,
              const int Q, const int L, const int key_planes, const int64_t key_bias, const int fused_finish)
{
    constexpr int NUP = NB - 5;
    constexpr bool kRankPass = K > 0 && !DR;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_cls = smem;
    unsigned char* s_rnk = smem + size_t(kPlaneRows) * nwords * 8;
    uint32_t* s_ro_all = reinterpret_cast<uint32_t*>(s_rnk + (kRankPass ? size_t(kPlaneRows) * nwords * rank_pitch(K) : 0));
    const int ro_stride = steps + 4;            // + 16 bytes: the row-offset vectors of different queries fall into different banks
    __shared__ Cand s_part[kPackMaxWarps][kPackMaxQ];
    __shared__ int64_t s_ptop[kPackMaxWarps][kPackMaxQ];
    __shared__ Cand s_best[kPackMaxQ];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, warps = nthreads >> 5;
    const int len2 = G.uniform_len2;
    const int64_t noff = G.len1 - len2 + 1;
    const int q0 = blockIdx.x * Q;
    const int nqb = (G.nq - q0) < Q ? (G.nq - q0) : Q;              // queries of this block
    const int j = tid / L, l = tid - j * L;                         // my query in the block, my offset word in the query
    const bool lane_on = j < nqb;
    const int jc = lane_on ? j : 0, lc = lane_on ? l : 0;           // idle lanes shadow lane 0 (addresses stay valid)
    const int64_t ln0 = int64_t(lc) * 32;
    const uint32_t vmask = lane_on ? valid_mask(ln0, 0, noff) : 0u;
    PSA_CHECK(G.uniform_len2 > 0 && G.tiles_per_query == 1 && G.last < 0 && Q <= kPackMaxQ && warps <= kPackMaxWarps &&
              Q * L <= nthreads && steps <= kScanChunkMax && L + (steps >> 5) + 1 <= nwords && nwords <= P.plane_words);

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_arrive_expect_tx(&s_bar, uint32_t(kPlaneRows) * uint32_t(nwords) * uint32_t(8 + (kRankPass ? 4 * K : 0)));
    }
    __syncthreads();
    pdl_wait();                                                     // the bit planes come from k_profile
    if (warp == 0 && lane < kPlaneRows) {
        tma_load_1d(s_cls + size_t(lane) * nwords * 8, P.cls_planes + int64_t(lane) * P.plane_words, uint32_t(nwords) * 8u, &s_bar);
        if (kRankPass)
            tma_load_1d(s_rnk + size_t(lane) * nwords * rank_pitch(K), P.rank_planes + int64_t(lane) * P.plane_words * K,
                        uint32_t(nwords) * 4u * K, &s_bar);
    }
    // per-step row offsets of the block's queries: their bytes are one contiguous run of nqb * len2 (equal lengths);
    // eight loads per thread in flight, then the padding steps (the all-zero row)
    {
        const uint8_t* src = P.seq2s + int64_t(q0) * len2;
        const int nbytes = nqb * len2;
        bool bad = false;
        for (int base = tid; base < nbytes; base += 8 * nthreads) {
            uint8_t v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = (base + u * nthreads) < nbytes ? src[base + u * nthreads] : uint8_t('A');
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int e = base + u * nthreads;
                if (e < nbytes) {
                    const int jj = e / len2, st = e - jj * len2;
                    uint32_t row = symbol_of(v[u]);
                    if (row == 0xFFu) { bad = true; row = 0; }
                    s_ro_all[jj * ro_stride + st] = row * uint32_t(nwords) * 8u;
                }
            }
        }
        const int pad = steps - len2;
        for (int e = tid; e < nqb * pad; e += nthreads) {
            const int jj = e / pad, st = len2 + (e - jj * pad);
            s_ro_all[jj * ro_stride + st] = uint32_t(kZeroRow) * uint32_t(nwords) * 8u;
        }
        if (bad) report_bad_symbol(P);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);

    const uint32_t* ro = s_ro_all + jc * ro_stride;
    const int groups = steps >> 5;
    uint32_t racc[K > 0 ? K : 1];
#pragma unroll
    for (int k = 0; k < (K > 0 ? K : 1); k++) racc[k] = 0;
    racc[0] = ~vmask;
    if (kRankPass) {
        for (int g = 0; g < groups; g++) {
            rank_group<K>(racc, reinterpret_cast<const char*>(s_rnk) + size_t(lc + g) * 4 * K, ro + g * 32);
            if (__all_sync(0xFFFFFFFFu, racc[0] == 0xFFFFFFFFu)) break;
        }
    }
    VCounter<NUP> A, B, C;
    A.clear(); B.clear(); C.clear();
    for (int g = 0; g < groups; g++)
        class_group<NUP>(A, B, C, reinterpret_cast<const char*>(s_cls) + size_t(lc + g) * 8, ro + g * 32);
    if (DR) racc[0] |= derive_top_rank<NB, NUP>(T, A, B, C);
    pdl_launch_dependents();

    // ---- per-lane keys, then one query segment of the warp at a time ------------------------------------
    const Cand none{ kKeyNone, 0x7FFFFFFF };
    Cand mine = none, ub = none;
    typename std::conditional<BS, SlicedKeys<NB, K>, OffsetKeys<NB, K, false>>::type keys;
    keys.build(T, len2, A, B, C, racc, key_planes, key_bias);
    const uint32_t umask = keys.scan(vmask, ln0, mine, ub);
    const int words_per_tile = G.tile >> 5;
    if (!T.exact) {
        // re-score mode: an upper estimate per 32-offset word for k_finish; words past the query's last one hold nothing
        const int64_t top = mine.key > ub.key ? mine.key : ub.key;
        if (lane_on) P.lane_keys[int64_t(q0 + j) * words_per_tile + l] = top;
        for (int e = tid; e < nqb * (words_per_tile - L); e += nthreads) {
            const int jj = e / (words_per_tile - L), w = e - jj * (words_per_tile - L);
            P.lane_keys[int64_t(q0 + jj) * words_per_tile + L + w] = kKeyNone;
        }
    }
    const int jlo = (warp * 32) / L;
    const int jhi = ((warp * 32 + 31) / L) < (nqb - 1) ? ((warp * 32 + 31) / L) : (nqb - 1);
    for (int jj = jlo; jj <= jhi; jj++) {                           // warp-uniform
        const bool in = lane_on && j == jj;
        Cand wb;
        int64_t wtop = kKeyNone;
        if (T.exact) {
            wb = settle_unresolved(T, P, keys, in ? mine : none, in ? ub : none, in ? umask : 0u, ln0, int64_t(q0 + jj) * len2, len2);
        } else {
            wb = warp_best(in ? mine : none);
            wtop = in ? (mine.key > ub.key ? mine.key : ub.key) : kKeyNone;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const int64_t o = __shfl_xor_sync(0xFFFFFFFFu, wtop, d);
                wtop = o > wtop ? o : wtop;
            }
        }
        if (lane == 0) { s_part[warp][jj - jlo] = wb; s_ptop[warp][jj - jlo] = wtop; }
    }
    __syncthreads();
    if (tid < nqb) {
        const int wlo = (tid * L) >> 5, whi = ((tid + 1) * L - 1) >> 5;
        Cand r = none;
        int64_t top = kKeyNone;
        for (int w = wlo; w <= whi; w++) {
            const int slot = tid - (w * 32) / L;
            take(r, s_part[w][slot].key, s_part[w][slot].off);
            top = s_ptop[w][slot] > top ? s_ptop[w][slot] : top;
        }
        s_best[tid] = r;
        if (!fused_finish) {
            TileRec rec;
            rec.key = r.key; rec.offset = r.off;
            rec.ub_key = top; rec.ub_offset = 0x7FFFFFFF;
            rec.score = 0.0; rec.flags = 0; rec.pad = 0;
            P.tiles[q0 + tid] = rec;                                // one tile per query: tile id == query id
        }
    }
    if (!fused_finish) return;
    __syncthreads();
    for (int jj = warp; jj < nqb; jj += warps) finish_query_warp(T, P, q0 + jj, int64_t(q0 + jj) * len2, len2, s_best[jj]);
}

size_t packed_smem_bytes(int rank_bytes_per_word, int nwords, int steps, int Q)
{
    return size_t(kPlaneRows) * nwords * (8 + size_t(rank_bytes_per_word)) + size_t(Q) * (steps + 4) * 4;
}

size_t batch_smem_bytes(int rank_planes, int chunk, int warps)
{
    const size_t nwords = size_t(round_up4(32 + chunk / 32));
    return size_t(kPlaneRows) * nwords * (8 + size_t(rank_pitch(rank_planes))) + size_t(warps) * chunk * 4;
}

// function attributes are per device and sticky: set them once per (kernel, device) instead of per launch
template <class Kernel>
void allow_big_smem(Kernel kernel, bool (&done)[64])
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && done[dev]) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (dev >= 0 && dev < 64) done[dev] = true;
}

template <int NB, int K, bool BS>
void launch_scan_inst(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int chunk, bool batch, int sm_count,
                      int key_planes, int64_t key_bias, const SliceGeom& SG, cudaStream_t stream)
{
    // the top-rank bit can be derived from the class planes when it is a function of the sign class that excludes
    // class 0 (the all-zero padding row reads as class 0 and must never count as a hit)
    const bool derive = K == 1 && T.top_rank_lut > 0 && (T.top_rank_lut & 1) == 0 && SG.allow_derive;
    if (SG.slices > 1) {
        // G is the finish geometry (256-offset tiles); the scan runs on its own warp tiles
        BatchGeom Gs = G;
        Gs.tile = SG.scan_tile;
        Gs.total_tiles = SG.scan_tiles;
        Gs.tiles_per_query = SG.scan_tiles;
        const int warps = Gs.tile / 1024;
        const int nwords = round_up4(warps * 32 + chunk / 32);
        const size_t smem = scan_smem_bytes(K, chunk, warps);
        static bool done[64];
        if constexpr (K == 1) {
            if (derive) {
                static bool done_dr[64];
                allow_big_smem(k_scan<NB, K, false, true, true>, done_dr);
                launch_dependent(k_scan<NB, K, false, true, true>, dim3(Gs.total_tiles, SG.slices), dim3(warps * 32), smem, stream, T, Gs, P,
                                 nwords, chunk, 0, int64_t(0), SG.slice_len, 0);
                if (SG.fused_combine) launch_dependent(k_combine<K, true>, dim3(G.total_tiles), dim3(kCombineThreads), 0, stream, T, G, P, SG.slices);
                else launch_dependent(k_combine<K, false>, dim3(G.total_tiles), dim3(kCombineThreads), 0, stream, T, G, P, SG.slices);
                return;
            }
        }
        allow_big_smem(k_scan<NB, K, false, true, false>, done);
        launch_dependent(k_scan<NB, K, false, true, false>, dim3(Gs.total_tiles, SG.slices), dim3(warps * 32), smem, stream, T, Gs, P, nwords, chunk, 0,
                         int64_t(0), SG.slice_len, 0);
        if (SG.fused_combine) launch_dependent(k_combine<K, true>, dim3(G.total_tiles), dim3(kCombineThreads), 0, stream, T, G, P, SG.slices);
        else launch_dependent(k_combine<K, false>, dim3(G.total_tiles), dim3(kCombineThreads), 0, stream, T, G, P, SG.slices);
        return;
    }
    if constexpr (NB <= 10) {
        if (SG.pack_q > 0 && !batch) {
            const int steps = int((G.uniform_len2 + 31) & ~31);
            const int64_t noff = G.len1 - G.uniform_len2 + 1;
            const int L = int((noff + 31) / 32);
            const int nwords = round_up4(L + steps / 32 + 1);
            const dim3 grid((G.nq + SG.pack_q - 1) / SG.pack_q), block(SG.pack_warps * 32);
            if constexpr (K == 1) {
                if (derive) {
                    static bool done_dr[64];
                    allow_big_smem(k_scan_packed<NB, K, BS, true>, done_dr);
                    launch_dependent(k_scan_packed<NB, K, BS, true>, grid, block, packed_smem_bytes(0, nwords, steps, SG.pack_q), stream, T, G, P,
                                     nwords, steps, SG.pack_q, L, key_planes, key_bias, SG.fused_finish ? 1 : 0);
                    return;
                }
            }
            static bool done[64];
            allow_big_smem(k_scan_packed<NB, K, BS, false>, done);
            launch_dependent(k_scan_packed<NB, K, BS, false>, grid, block, packed_smem_bytes(rank_pitch(K), nwords, steps, SG.pack_q), stream, T, G, P,
                             nwords, steps, SG.pack_q, L, key_planes, key_bias, SG.fused_finish ? 1 : 0);
            return;
        }
    }
    if (batch) {
        const int warps = 4;
        const int nwords = round_up4(32 + chunk / 32);
        const size_t smem = batch_smem_bytes(K, chunk, warps);
        const int tiles = int((G.len1 + 1023) / 1024);              // upper bound: tiles past a query's last offset are skipped
        // about 8 blocks per SM in flight over the whole launch; more queries per block = better use of a staged window
        int64_t qpb = (int64_t(G.nq) * tiles + int64_t(sm_count) * 8 - 1) / (int64_t(sm_count) * 8);
        qpb = ((qpb + warps - 1) / warps) * warps;
        if (qpb < warps) qpb = warps;
        if (qpb > 4096) qpb = 4096;
        dim3 grid((G.nq + qpb - 1) / qpb, tiles);
        static bool done[64];
        if constexpr (K == 1) {
            if (derive) {
                static bool done_dr[64];
                allow_big_smem(k_scan_batch<NB, K, BS, true>, done_dr);
                launch_dependent(k_scan_batch<NB, K, BS, true>, grid, dim3(warps * 32), smem, stream, T, G, P, nwords, chunk, int(qpb), key_planes,
                                 key_bias);
                return;
            }
        }
        allow_big_smem(k_scan_batch<NB, K, BS, false>, done);
        launch_dependent(k_scan_batch<NB, K, BS, false>, grid, dim3(warps * 32), smem, stream, T, G, P, nwords, chunk, int(qpb), key_planes, key_bias);
    } else {
        const int warps = G.tile / 1024;
        const int nwords = round_up4(warps * 32 + chunk / 32);
        const size_t smem = scan_smem_bytes(K, chunk, warps);
        static bool done[64];
        if constexpr (K == 1) {
            if (derive) {
                static bool done_dr[64];
                allow_big_smem(k_scan<NB, K, BS, false, true>, done_dr);
                launch_dependent(k_scan<NB, K, BS, false, true>, dim3(G.total_tiles), dim3(warps * 32), smem, stream, T, G, P, nwords, chunk,
                                 key_planes, key_bias, 0, SG.fused_finish ? 1 : 0);
                return;
            }
        }
        allow_big_smem(k_scan<NB, K, BS, false, false>, done);
        launch_dependent(k_scan<NB, K, BS, false, false>, dim3(G.total_tiles), dim3(warps * 32), smem, stream, T, G, P, nwords, chunk, key_planes, key_bias, 0, SG.fused_finish ? 1 : 0);
    }
}

template <int NB>
void launch_scan_nb(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int K, int chunk, bool batch, int sm_count,
                    int key_planes, int64_t key_bias, const SliceGeom& SG, cudaStream_t stream)
{
#define PSA_SCAN_CASE(KK)                                                                                                  \
    if (key_planes > 0) launch_scan_inst<NB, KK, true>(T, G, P, chunk, batch, sm_count, key_planes, key_bias, SG, stream); \
    else launch_scan_inst<NB, KK, false>(T, G, P, chunk, batch, sm_count, 0, 0, SG, stream);                              \
    break;
    switch (K) {
    case 0: PSA_SCAN_CASE(0)
    case 1: PSA_SCAN_CASE(1)
    case 2: PSA_SCAN_CASE(2)
    default: PSA_SCAN_CASE(4)
    }
#undef PSA_SCAN_CASE
}

} // namespace

// The bit-sliced epilogue applies when keys are small integers: exact mode, |multipliers| < 2^kSlicedMaxBits, rank
// terms spread <= 255 and every biased key below 2^sliced_planes(NB).  Returns the plane count (0 = not applicable).
int sliced_key_planes(const DeviceTable& T, int64_t max_len2, int nb, int64_t* bias_out)
{
    if (!T.exact) return 0;
    const int64_t k0 = T.kcls[0], ka = T.kcls[1] - T.kcls[0], kb = T.kcls[2] - T.kcls[0];
    const int64_t kc = T.kcls[3] - T.kcls[1] - T.kcls[2] + T.kcls[0];
    auto mag = [](int64_t v) { return v < 0 ? -v : v; };
    const int64_t lim = (int64_t(1) << kSlicedMaxBits) - 1;
    if (mag(ka) > lim || mag(kb) > lim || mag(kc) > lim || mag(k0) > (int64_t(1) << 20)) return 0;
    int64_t dmin = INT64_MAX, dmax = INT64_MIN;
    for (int r = 1; r <= T.nranks; r++) { dmin = std::min(dmin, T.kdiff[r]); dmax = std::max(dmax, T.kdiff[r]); }
    if (T.nranks < 1 || dmax - dmin > 255) return 0;
    const int64_t bias = max_len2 * (mag(ka) + mag(kb) + mag(kc) + mag(k0)) + std::max(mag(dmin), mag(dmax)) + 1;
    if (2 * bias + 512 >= (int64_t(1) << sliced_planes(nb))) return 0;
    *bias_out = bias;
    return sliced_planes(nb);
}

int scan_chunk_steps(int, int64_t max_len2)
{
    const int64_t padded = (max_len2 + 127) & ~int64_t(127);      // multiple of 128: window rows start on 16 bytes
    return int(padded < kScanChunkMax ? padded : kScanChunkMax);
}

// batch mode: every query of the batch fits one staged window and no offset range is imposed
// and there are enough (query, tile) tasks that sharing a staged window among the queries of a block pays off
bool scan_batch_mode(const BatchGeom& G, int64_t max_len2, int sm_count)
{
    const int64_t tiles = (G.len1 + 1023) / 1024;                   // gridDim.y of k_scan_batch
    const int64_t tasks = int64_t(G.nq) * tiles;
    return max_len2 <= 1023 && G.last < 0 && tiles <= 65535 && tasks >= int64_t(sm_count) * 64;
}

size_t scan_smem_bytes(int rank_planes, int chunk, int warps)
{
    const size_t nwords = size_t(round_up4(warps * 32 + chunk / 32));
    const size_t entry = 4 * size_t(rank_planes) > 8 ? 4 * size_t(rank_planes) : 8;      // the two passes share the window
    return size_t(kPlaneRows) * nwords * entry + size_t(chunk) * 4;
}

int64_t scan_plane_words(int64_t len1)
{
    // every window read stays inside the row: last tile start < len1 + 128, + tile + chunk + rounding
    const int64_t bits = len1 + 128 + kScanTile + kScanChunkMax + 256;
    return ((bits + 31) / 32 + 3) & ~int64_t(3);
}

bool scan_packed_fits(int64_t len1, int64_t len2)
{
    const int64_t noff = len1 - len2 + 1;
    if (noff < 1 || len2 < 1 || len2 > 1023) return false;
    const int64_t lanes = (noff + 31) / 32, steps = (len2 + 31) & ~int64_t(31);
    const int64_t nwords = round_up4(int(lanes + steps / 32 + 1));
    // worst case shared memory: 4 rank planes, 8 queries
    const size_t smem = size_t(kPlaneRows) * nwords * (8 + 16) + size_t(kPackMaxQ) * (steps + 4) * 4;
    return lanes <= 32 * kPackMaxWarps && nwords <= scan_plane_words(len1) && smem <= 160 * 1024;
}

void launch_profile(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int rank_planes, int sm_count,
                    cudaStream_t stream)
{
    int64_t blocks = (P.plane_words + kProfileThreads - 1) / kProfileThreads;
    const int64_t cap = int64_t(sm_count) * 16;
    if (blocks > cap) blocks = cap;
    switch (rank_planes) {
    case 0: k_profile<0><<<(int)blocks, kProfileThreads, 0, stream>>>(T, G, P); break;
    case 1: k_profile<1><<<(int)blocks, kProfileThreads, 0, stream>>>(T, G, P); break;
    case 2: k_profile<2><<<(int)blocks, kProfileThreads, 0, stream>>>(T, G, P); break;
    default: k_profile<4><<<(int)blocks, kProfileThreads, 0, stream>>>(T, G, P); break;
    }
}

void launch_scan(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int rank_planes, int64_t max_len2, bool batch,
                 bool sliced_ok, int sm_count, const SliceGeom& SG, cudaStream_t stream)
{
    if (G.total_tiles < 1) return;
    const int chunk = scan_chunk_steps(rank_planes, max_len2);
    int64_t key_bias = 0;
    const int nb = max_len2 <= 127 ? 7 : max_len2 <= 1023 ? 10 : 15;
    const int key_planes = sliced_ok ? sliced_key_planes(T, max_len2, nb, &key_bias) : 0;
    batch = batch && max_len2 <= 1023 && G.last < 0 && G.tile == 1024;
    if (max_len2 <= 127) launch_scan_nb<7>(T, G, P, rank_planes, chunk, batch, sm_count, key_planes, key_bias, SG, stream);
    else if (max_len2 <= 1023) launch_scan_nb<10>(T, G, P, rank_planes, chunk, batch, sm_count, key_planes, key_bias, SG, stream);
    else launch_scan_nb<15>(T, G, P, rank_planes, chunk, batch, sm_count, key_planes, key_bias, SG, stream);
}

} // namespace psa
