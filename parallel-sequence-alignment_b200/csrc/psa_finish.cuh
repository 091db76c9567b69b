// psa_finish.cuh -- the finish step (winner over the tile records, re-score of candidate words, result record) as a
// device function shared by k_finish (psa_kernels.cu) and the last block of k_combine (psa_scan.cu).
#pragma once
#include "psa_kernels.cuh"
#include "psa_device.cuh"

namespace psa {
#if defined(PSA_FINISH_TRACE)
// debug build only (make EXTRA=-DPSA_FINISH_TRACE): SM clock at the phase boundaries of the finish step, thread 0 of block 0
__device__ long long g_finish_trace[16];
#define FIN_MARK(k) do { if (threadIdx.x == 0 && block_index == 0) g_finish_trace[k] = clock64(); } while (0)
#else
#define FIN_MARK(k) ((void)0)
#endif
namespace {

// -------------------------------------------------------------------------------------------------
// Finish: one warp per query, 8 queries per block (WPQ = 1), or -- for small batches, where a single warp
// would crawl through dependent global loads -- the whole block on one query (WPQ = 8).
//  (1) winner over the query's tile records under (key desc, offset asc);
//  (2) only when the weights are not exactly summable and the records come from the bit-sliced scan:
//      the integer keys order offsets like the reference only up to key_slack, so every 32-offset word
//      whose key estimate is within key_slack of the best key is re-scored here -- one lane per offset,
//      a double accumulated over i = 0..len2-1 in the reference's order (cpu_funcs.c:271-299), symbols
//      staged through shared memory 256 steps at a time -- and the winner is chosen among those doubles
//      under is_swapable's order (cuda_funcs.cu:290-307);
//  (3) one pass over the winning alignment for the sign counts, the first position carrying the best
//      rank (cpu_funcs.c:287-294: strict compare, so the lowest i wins ties) and its replacement letter.
// -------------------------------------------------------------------------------------------------
constexpr int kFinishWarps = 8;
constexpr int kFinishThreads = kFinishWarps * 32;
constexpr int kFinishChunk = 1024;
constexpr int kFinishList = 32;       // candidate tiles remembered per query before falling back to a full walk

// Tile records and lane keys are read with ld.global.cg (L2): when this runs as the tail of k_combine's last block they
// were written by OTHER blocks of the same launch, and an L1 line filled earlier in the launch would be stale.
// The body is a device function so that k_combine's last block can run it too (slice mode on a small grid: one launch
// fewer per single query).  `block_index` stands for blockIdx.x of a k_finish launch; WAIT: the caller has not yet waited
// for the kernels before it.
// COOP (block per query, k_finish<8> only): when a query has only a few candidate words, each is re-scored by the whole
// block instead of by one warp.  Warps 1..7 take turns producing the 32 lanes' addends, 32 steps per slot of an 8-slot ring
// in dynamic shared memory (symbol staging, pair index, weight lookup: ~1300 cycles per slot for a lone warp, which is what
// made a one-warp re-score run at 41 cycles per step); warp 0 only loads its addends and runs the chain of dependent double
// adds -- the one part whose order is the reference's (cpu_funcs.c:278) and has to be serial -- at the pace of the FP64 pipe.
constexpr int kFinishRingSteps = 32;
constexpr int kFinishRingSlots = 8;
constexpr int kFinishCoopMaxWords = 4;                                   // more candidate words than this: one warp per word, side by side
constexpr size_t kFinishRingBytes = size_t(kFinishRingSlots) * kFinishRingSteps * 32 * sizeof(double);
constexpr size_t kFinishRankBytes = size_t(kFinishWarps) * 32 * sizeof(uint32_t);
constexpr size_t kFinishDynBytes = kFinishRingBytes + kFinishRankBytes + 2 * kFinishRingSlots * sizeof(uint64_t);

template <int WPQ, bool WAIT, bool COOP = false>
__device__ __forceinline__ void finish_body(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, const int scan_records,
                                            const int block_index)
{
    extern __shared__ __align__(16) unsigned char fin_dyn[];            // COOP only: ring | best ranks per producer | full[8], empty[8]
    constexpr int kGroup = WPQ * 32;                                    // threads working on one query
    __shared__ Cand s_part[kFinishWarps];
    __shared__ unsigned long long s_pos;
    __shared__ int s_cnt[4];
    __shared__ __align__(16) uint8_t s_code[kSymbols * kRowPad];
    __shared__ double s_w[4];
    __shared__ int s_list[kFinishWarps][kFinishList];
    __shared__ int s_nlist[kFinishWarps];
    __shared__ uint16_t s_q[kFinishWarps][kFinishChunk];               // Seq2 symbol * kRowPad
    __shared__ uint8_t s_win[kFinishWarps][kFinishChunk + 32];          // Seq1 symbols under the 32 offsets
    __shared__ double s_wtab[kSymbols * kRowPad];                       // pair weight by (Seq2 symbol * kRowPad + Seq1 symbol)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FIN_MARK(0);
    if (tid < 4) s_w[tid] = T.wcls[tid];
    if (COOP && tid < 2 * kFinishRingSlots) mbar_init(reinterpret_cast<uint64_t*>(fin_dyn + kFinishRingBytes + kFinishRankBytes) + tid, 1);
    if (tid == 0) s_pos = 0ull;
    if (tid < 4) s_cnt[tid] = 0;
    if (WAIT) pdl_wait();                                               // tile records come from the kernels before us
    for (int k = tid; k < kSymbols * kRowPad / 4; k += kFinishThreads)
        reinterpret_cast<uint32_t*>(s_code)[k] = reinterpret_cast<const uint32_t*>(P.code_table)[k];
    if (!T.exact)
        for (int k = tid; k < kSymbols * kRowPad; k += kFinishThreads) s_wtab[k] = T.wcls[P.code_table[k] & 3u];
    __syncthreads();
    FIN_MARK(1);
    const int q = WPQ == 1 ? block_index * kFinishWarps + warp : block_index;
    if (q >= G.nq) return;
    const int gtid = WPQ == 1 ? lane : tid;                             // index within the query's thread group
    const int gwarp = WPQ == 1 ? 0 : warp;

    const QueryGeom qg = query_geom(G, P.qoff, P.tile_start, q);
    const int t0 = qg.tile0, t1 = q + 1 < G.nq ? first_tile_of(G, P.tile_start, q + 1) : G.total_tiles;
    PSA_CHECK(t0 >= 0 && t0 <= t1 && t1 <= G.total_tiles);
    const int64_t qbeg = qg.qbeg;
    const int len2 = qg.len2;
    const int64_t first = G.last >= 0 ? G.first : 0;
    const int64_t last = G.last >= 0 ? G.last : G.len1 - len2 + 1;
    const uint8_t* b = P.seq2s + qbeg;

    // A query can have thousands of tile records and this is a latency-bound walk: 8 records per thread are fetched
    // together (one round trip), then compared.
    Cand mine{ kKeyNone, 0x7FFFFFFF };
    for (int tbase = t0 + gtid; tbase < t1; tbase += 8 * kGroup) {
        int64_t k8[8];
        int32_t o8[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int t = tbase + u * kGroup;
            k8[u] = t < t1 ? __ldcg(&P.tiles[t].key) : kKeyNone;
            o8[u] = t < t1 ? __ldcg(&P.tiles[t].offset) : 0x7FFFFFFF;
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (better(k8[u], o8[u], mine.key, mine.off)) { mine.key = k8[u]; mine.off = o8[u]; }
    }
    Cand win = WPQ == 1 ? warp_best(mine) : block_best<kFinishThreads>(mine, s_part);

    if (!T.exact && scan_records) {
    FIN_MARK(2);
        const int64_t threshold = win.key == kKeyNone ? kKeyNone : win.key - T.key_slack;   // |key| < 2^61: no wrap
        const int tile_words = G.tile >> 5;
        Cand mine2{ kKeyNone, 0x7FFFFFFF };
        int words = 0, cand_seq = 0;
        // candidate tiles, found in parallel (a query can have thousands of tiles and one or two candidates)
        int* my_list = s_list[WPQ == 1 ? warp : 0];
        int* my_count = &s_nlist[WPQ == 1 ? warp : 0];
        if (gtid == 0) *my_count = 0;
        if (WPQ == 1) __syncwarp(); else __syncthreads();
        for (int tbase = t0 + gtid; tbase < t1; tbase += 8 * kGroup) {
            int64_t top8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int t = tbase + u * kGroup;
                const int64_t k = t < t1 ? __ldcg(&P.tiles[t].key) : kKeyNone, ub = t < t1 ? __ldcg(&P.tiles[t].ub_key) : kKeyNone;
                top8[u] = k > ub ? k : ub;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (top8[u] == kKeyNone || top8[u] < threshold) continue;
                const int slot = atomicAdd(my_count, 1);
                if (slot < kFinishList) my_list[slot] = tbase + u * kGroup;
            }
        }
        if (WPQ == 1) __syncwarp(); else __syncthreads();
        FIN_MARK(3);
        const int nlist = *my_count;
        const bool listed = nlist <= kFinishList;               // else: walk every tile (correct, just slower)
        const int ntry = listed ? nlist : (t1 - t0);
        FIN_MARK(4);
        // COOP: how many candidate words are there?  (every warp counts the same sequence; a handful of tiles at most)
        bool coop = false;
        int coop_blk = 0;                                               // ring blocks so far (the same on every warp)
        if (COOP && listed) {
            int total_words = 0;
            for (int k = 0; k < nlist; k++) {
                const int64_t* lk = P.lane_keys + int64_t(my_list[k]) * tile_words;
                for (int w0 = 0; w0 < tile_words; w0 += 32) {
                    const int64_t kk = (w0 + lane) < tile_words ? __ldcg(lk + w0 + lane) : kKeyNone;
                    total_words += __popc(__ballot_sync(0xFFFFFFFFu, kk != kKeyNone && kk >= threshold));
                }
            }
            coop = total_words >= 1 && total_words <= kFinishCoopMaxWords;
        }
        for (int k = 0; k < ntry; k++) {
            const int t = listed ? my_list[k] : t0 + k;
            if (!listed) {
                const int64_t rk = __ldcg(&P.tiles[t].key), rub = __ldcg(&P.tiles[t].ub_key);
                const int64_t top = rk > rub ? rk : rub;
                if (top == kKeyNone || top < threshold) continue;
            }
            const int64_t* lk = P.lane_keys + int64_t(t) * tile_words;
            const int64_t tb = tile_base(first) + int64_t(t - t0) * G.tile;
            for (int w0 = 0; w0 < tile_words; w0 += 32) {
                // which of the next 32 words are candidates (one ballot instead of 32 broadcast loads)
                const int64_t k = (w0 + lane) < tile_words ? __ldcg(lk + w0 + lane) : kKeyNone;
                uint32_t cand = __ballot_sync(0xFFFFFFFFu, k != kKeyNone && k >= threshold);
                while (cand) {
                    const int w = w0 + __ffs(int(cand)) - 1;
                    cand &= cand - 1u;
                    // candidate words go round the warps of the block in the order they are met (every warp walks the same
                    // sequence), so that k <= WPQ words are always re-scored side by side -- by word index two candidates of
                    // different tiles could land on the same warp and double the longest chain
                    if (!(COOP && coop) && WPQ > 1 && (cand_seq++ % WPQ) != gwarp) continue;
                    const int64_t n0 = tb + 32 * w;                    // the word's first offset; this lane owns n0 + lane
                    const int64_t n = n0 + lane;
                    double total = 0.0;
                    uint32_t best_rank = 0;
                    if (COOP && coop) __syncthreads();                 // the previous word is done with the ring and the rank slots
                    // COOP, queries of up to 8192 symbols: the whole query and the whole Seq1 window under the word are staged ONCE,
                    // by all eight warps together, into the per-warp staging arrays taken as one flat array each -- every load of
                    // both is in flight at the same time (one trip to L2 instead of two per 1024-step chunk and warp), and the
                    // producers then run through the query without a chunk boundary at which the consumer would starve.
                    const bool flat = COOP && coop && len2 <= kFinishWarps * kFinishChunk;
                    const int chunk = flat ? len2 : kFinishChunk;
                    uint16_t* const flat_q = &s_q[0][0];
                    uint8_t* const flat_win = &s_win[0][0];
                    if (flat) {
                        const uint8_t* qsrc = b;
                        const int ft = int(threadIdx.x);
                        if ((reinterpret_cast<uintptr_t>(qsrc) & 15u) == 0) {
                            for (int i = ft * 16; i < len2; i += kFinishThreads * 16) {
                                const uint4 v = *reinterpret_cast<const uint4*>(qsrc + i);
                                const uint32_t w4[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                                for (int t = 0; t < 16; t++) {
                                    uint32_t c2 = symbol_of(uint8_t(w4[t >> 2] >> (8 * (t & 3))));
                                    if (c2 == 0xFFu) c2 = 0;           // flagged by the kernels before us
                                    if (i + t < len2) flat_q[i + t] = uint16_t(c2 * kRowPad);
                                }
                            }
                        } else {
                            for (int i = ft; i < len2; i += kFinishThreads) {
                                uint32_t c2 = symbol_of(qsrc[i]);
                                if (c2 == 0xFFu) c2 = 0;
                                flat_q[i] = uint16_t(c2 * kRowPad);
                            }
                        }
                        for (int i = ft * 16; i < len2 + 31; i += kFinishThreads * 16) {
                            uint4 v = make_uint4(0u, 0u, 0u, 0u);
                            if (n0 + i < G.len1) v = *reinterpret_cast<const uint4*>(P.seq1 + n0 + i);      // n0 is a multiple of 32
                            const uint32_t w4[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                            for (int t = 0; t < 16; t++) {
                                uint32_t c1 = (n0 + i + t) < G.len1 ? symbol_of(uint8_t(w4[t >> 2] >> (8 * (t & 3)))) : 0u;
                                if (c1 == 0xFFu) c1 = 0;
                                if (i + t < len2 + 31) flat_win[i + t] = uint8_t(c1);
                            }
                        }
                        __syncthreads();
                    }
                    if (COOP && coop && gwarp == 0) {
                        // ---- consumer: addends from the ring, adds in step order --------------------------------------
                        const double* ring = reinterpret_cast<const double*>(fin_dyn);
                        uint64_t* bars = reinterpret_cast<uint64_t*>(fin_dyn + kFinishRingBytes + kFinishRankBytes);      // full[8], empty[8]
                        int nblk = 0;
                        for (int c0 = 0; c0 < len2; c0 += chunk)
                            nblk += (((len2 - c0) < chunk ? (len2 - c0) : chunk) + kFinishRingSteps - 1) / kFinishRingSteps;
                        for (int kb = 0; kb < nblk; kb++) {
                            const int slot = coop_blk % kFinishRingSlots, use = coop_blk / kFinishRingSlots;
                            mbar_wait(bars + slot, uint32_t(use) & 1u);                  // produced?
                            const double* src = ring + size_t(slot) * kFinishRingSteps * 32 + lane;
                            double a[kFinishRingSteps];
#pragma unroll
                            for (int u = 0; u < kFinishRingSteps; u++) a[u] = src[u * 32];
#pragma unroll
                            for (int u = 0; u < kFinishRingSteps; u++) total += a[u];
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bars + kFinishRingSlots + slot);  // consumed
                            coop_blk++;
                        }
                    } else
                    for (int c0 = 0; c0 < len2; c0 += chunk) {
                        const int cl = (len2 - c0) < chunk ? (len2 - c0) : chunk;
                        __syncwarp();
                        // symbols of this chunk, 16 bytes per lane per load (both device buffers carry 64 bytes of padding;
                        // the Seq1 window starts on a multiple of 32, the query start may be anywhere)
                        const uint8_t* qsrc = b + c0;
                        if (flat) {
                            // staged above, once, by the whole block
                        } else
                        if ((reinterpret_cast<uintptr_t>(qsrc) & 15u) == 0) {
                            for (int i = lane * 16; i < cl; i += 512) {
                                const uint4 v = *reinterpret_cast<const uint4*>(qsrc + i);
                                const uint32_t w4[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                                for (int t = 0; t < 16; t++) {
                                    uint32_t c2 = symbol_of(uint8_t(w4[t >> 2] >> (8 * (t & 3))));
                                    if (c2 == 0xFFu) c2 = 0;           // flagged by the kernels before us
                                    if (i + t < cl) s_q[warp][i + t] = uint16_t(c2 * kRowPad);
                                }
                            }
                        } else {
#pragma unroll 8
                            for (int i = lane; i < cl; i += 32) {
                                uint32_t c2 = symbol_of(qsrc[i]);
                                if (c2 == 0xFFu) c2 = 0;
                                s_q[warp][i] = uint16_t(c2 * kRowPad);
                            }
                        }
                        const int64_t p0 = n0 + c0;                    // multiple of 32
                        if (!flat)
                        for (int i = lane * 16; i < cl + 31; i += 512) {
                            uint4 v = make_uint4(0u, 0u, 0u, 0u);
                            if (p0 + i < G.len1) v = *reinterpret_cast<const uint4*>(P.seq1 + p0 + i);
                            const uint32_t w4[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                            for (int t = 0; t < 16; t++) {
                                uint32_t c1 = (p0 + i + t) < G.len1 ? symbol_of(uint8_t(w4[t >> 2] >> (8 * (t & 3)))) : 0u;
                                if (c1 == 0xFFu) c1 = 0;
                                if (i + t < cl + 31) s_win[warp][i + t] = uint8_t(c1);
                            }
                        }
                        __syncwarp();
                        const uint8_t* wv = flat ? flat_win + lane : &s_win[warp][lane];
                        const uint16_t* qv = flat ? flat_q : &s_q[warp][0];
                        // The sum must be the reference's: one rounding per step, in step order -- a chain of dependent
                        // double adds run by one warp, with nothing to hide latency behind.  Measured on B200
                        // (tools/probes/dadd_probe.cu): 8 cycles per step when each addend is a plain shared-memory load
                        // issued well ahead, 36-45 when a select or a constant-bank operand sits on the chain.  So 32 steps
                        // at a time: all symbol loads, then all weight loads (one table indexed by the symbol pair), then
                        // the 32 adds.  (A hand-pipelined version with the next block's loads between the adds was slower:
                        // ptxas places the address adds right behind their loads and a lone warp stalls on each.)
                        if (COOP && coop) {
                            // ---- producers (warps 1..7): every seventh ring block each -------------------------------------
                            double* ring = reinterpret_cast<double*>(fin_dyn);
                            uint64_t* bars = reinterpret_cast<uint64_t*>(fin_dyn + kFinishRingBytes + kFinishRankBytes);
                            for (int i0 = 0; i0 < cl; i0 += kFinishRingSteps, coop_blk++) {
                                if (coop_blk % (kFinishWarps - 1) != gwarp - 1) continue;
                                const int slot = coop_blk % kFinishRingSlots, use = coop_blk / kFinishRingSlots;
                                if (use >= 1) mbar_wait_patient(bars + kFinishRingSlots + slot, uint32_t(use - 1) & 1u);    // the slot's previous content is consumed
                                double* dst = ring + size_t(slot) * kFinishRingSteps * 32 + lane;
                                if (i0 + kFinishRingSteps <= cl) {
                                    uint32_t idx[kFinishRingSteps];
                                    double w32[kFinishRingSteps];
#pragma unroll
                                    for (int u = 0; u < kFinishRingSteps; u++) idx[u] = uint32_t(qv[i0 + u]) + wv[i0 + u];
#pragma unroll
                                    for (int u = 0; u < kFinishRingSteps; u++) {
                                        w32[u] = s_wtab[idx[u]];
                                        best_rank = max(best_rank, uint32_t(s_code[idx[u]]) >> 2);
                                    }
#pragma unroll
                                    for (int u = 0; u < kFinishRingSteps; u++) dst[u * 32] = w32[u];
                                } else {
                                    for (int u = 0; u < kFinishRingSteps; u++) {
                                        double wgt = 0.0;                             // steps past the end add +0.0: exact
                                        if (i0 + u < cl) {
                                            const uint32_t idx = uint32_t(qv[i0 + u]) + wv[i0 + u];
                                            wgt = s_wtab[idx];
                                            best_rank = max(best_rank, uint32_t(s_code[idx]) >> 2);
                                        }
                                        dst[u * 32] = wgt;
                                    }
                                }
                                __threadfence_block();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(bars + slot);              // produced
                            }
                            continue;
                        }
                        for (int i0 = 0; i0 < cl; i0 += 32) {
                            if (i0 + 32 <= cl) {
                                uint32_t idx[32];
                                double w32[32];
#pragma unroll
                                for (int u = 0; u < 32; u++) idx[u] = uint32_t(qv[i0 + u]) + wv[i0 + u];
#pragma unroll
                                for (int u = 0; u < 32; u++) {
                                    w32[u] = s_wtab[idx[u]];
                                    best_rank = max(best_rank, uint32_t(s_code[idx[u]]) >> 2);
                                }
#pragma unroll
                                for (int u = 0; u < 32; u++) total += w32[u];
                            } else {
                                for (int i = i0; i < cl; i++) {
                                    const uint32_t code = s_code[uint32_t(qv[i]) + wv[i]];
                                    total += s_w[code & 3u];
                                    best_rank = max(best_rank, code >> 2);
                                }
                            }
                        }
                    }
                    if (COOP && coop) {
                        // the producers' best ranks meet in shared memory; warp 0 owns the word's result
                        uint32_t* rank_x = reinterpret_cast<uint32_t*>(fin_dyn + kFinishRingBytes);
                        if (gwarp > 0) rank_x[gwarp * 32 + lane] = best_rank;
                        __syncthreads();
                        if (gwarp > 0) continue;
#pragma unroll
                        for (int g = 1; g < kFinishWarps; g++) best_rank = max(best_rank, rank_x[g * 32 + lane]);
                    }
                    words++;
                    if (best_rank && n >= first && n < last) {
                        const double score = total + T.wdiff[best_rank];                   // cpu_funcs.c:299
                        const int64_t key = sortable_from_double(T.is_max ? score : -score);
                        if (better(key, int32_t(n), mine2.key, mine2.off)) { mine2.key = key; mine2.off = int32_t(n); }
                    }
                }
            }
        }
        FIN_MARK(5);
        if (lane == 0 && words) atomicAdd(P.cand_count, words);
        win = WPQ == 1 ? warp_best(mine2) : block_best<kFinishThreads>(mine2, s_part);
    }

    QueryRec out;
    out.score = T.is_max ? -INFINITY : INFINITY;      // what the reference returns when nothing can be mutated
    out.offset = -1; out.char_offset = -1; out.ch = 0; out.rank = 0;
    out.counts[0] = out.counts[1] = out.counts[2] = out.counts[3] = 0;
    if (win.key == kKeyNone) {            // no mutation possible at any offset (never seen in practice)
        if (gtid == 0) P.out[q] = out;
        return;
    }

    PSA_CHECK(win.off >= first && win.off < last && int64_t(win.off) + len2 <= G.len1);
    FIN_MARK(6);
    const uint8_t* a = P.seq1 + win.off;
    int cnt[4] = { 0, 0, 0, 0 };
    unsigned long long pos = 0ull;        // (rank << 32) | ~i  -> max = best rank, then lowest i
#pragma unroll 4
    for (int i = gtid; i < len2; i += kGroup) {
        uint32_t c1 = symbol_of(a[i]), c2 = symbol_of(b[i]);
        if (c1 == 0xFFu || c2 == 0xFFu) { c1 = 0; c2 = 0; }
        const uint32_t code = s_code[c2 * kRowPad + c1];
        cnt[code & 3u]++;
        const unsigned long long p = (uint64_t(code >> 2) << 32) | uint32_t(~uint32_t(i));
        pos = p > pos ? p : pos;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, pos, d);
        pos = o > pos ? o : pos;
#pragma unroll
        for (int c = 0; c < 4; c++) cnt[c] += __shfl_xor_sync(0xFFFFFFFFu, cnt[c], d);
    }
    if (WPQ > 1) {
        if (lane == 0) {
            atomicMax(&s_pos, pos);
#pragma unroll
            for (int c = 0; c < 4; c++) atomicAdd(&s_cnt[c], cnt[c]);
        }
        __syncthreads();
    FIN_MARK(7);
        pos = s_pos;
#pragma unroll
        for (int c = 0; c < 4; c++) cnt[c] = s_cnt[c];
    }
    if (gtid == 0) {
        const int rank = int(pos >> 32);
        const int i = int(~uint32_t(pos));
        uint32_t c1 = symbol_of(a[i]), c2 = symbol_of(b[i]);
        if (c1 == 0xFFu || c2 == 0xFFu) { c1 = 0; c2 = 0; }
        out.offset = win.off;
        out.char_offset = i;
        out.ch = T.sub[c2][c1];
        out.rank = rank;
        for (int c = 0; c < 4; c++) out.counts[c] = cnt[c];
        if (T.exact) {
            // every product and partial sum is exactly representable (psa_table.cpp), so this IS the reference's
            // sequential sum + difference, bit for bit; explicit _rn ops keep the compiler from contracting to FMA
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < 4; c++) s = __dadd_rn(s, __dmul_rn(double(cnt[c]), T.wcls[c]));
            out.score = __dadd_rn(__dadd_rn(s, T.wdiff[rank]), 0.0);
        } else {
            // engine 1 tile keys and re-scored keys are sortable images of the reference's double
            out.score = __dadd_rn((T.is_max ? 1.0 : -1.0) * double_from_sortable(win.key), 0.0);
        }
        if (rank <= 0) { out.offset = -1; out.char_offset = -1; out.ch = 0; out.score = T.is_max ? -INFINITY : INFINITY; }
        P.out[q] = out;
    }
}


} // namespace
} // namespace psa
