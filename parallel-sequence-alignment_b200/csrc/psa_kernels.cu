// psa_kernels.cu -- exact scalar kernel, candidate selection and per-query finalisation (sm_100a).
//
// Together with psa_scan.cu these replace the reference's calc_mutants_scores / reduction /
// reduce_last_results kernels (cuda_funcs.cu:149-264).  Differences in kind, not in result:
//   * no per-pair substitution search on the device: one byte of a host-resolved table per pair;
//   * no global-memory scratch of offsets x block_size records (the reference allocates
//     offsets*B*32 bytes per call, cuda_funcs.cu:57-64): partial results live in registers and
//     are reduced with warp shuffles, one 40-byte record per tile of offsets reaches HBM;
//   * integer sign counts and an int64 key instead of tree-summed doubles, so the order of
//     offsets does not depend on the reduction shape (the reference's GPU path differs from its
//     own CPU path there, and races across blocks, cuda_funcs.cu:239-264);
//   * when the weights are not exactly summable, candidate tiles are re-scored with a double
//     accumulated in the reference's CPU order (cpu_funcs.c:271-299) so the winner is the CPU
//     reference's winner.
#include "psa_kernels.cuh"
#include "psa_device.cuh"
#include "psa_finish.cuh"

#include <algorithm>

namespace psa {

namespace {

constexpr int kExactThreads = 256;
constexpr int kExactChunk = 2048;       // Seq2 symbols staged per pass

// -------------------------------------------------------------------------------------------------
// Exact scalar kernel: one thread per offset, all of Seq2 in ascending i.
//   EXACT = true : integer sign counts -> int64 fixed-point key
//   EXACT = false: double accumulated sequentially in i exactly like find_best_mutant_offset
//                  (cpu_funcs.c:271-299); the key is the goal-signed double mapped to a sortable int64
// -------------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void __launch_bounds__(kExactThreads)
k_exact_tiles(const __grid_constant__ DeviceTable T, const BatchGeom G, const BatchPtrs P)
{
    __shared__ uint64_t s_tab[kSymbols * kRowPad];
    __shared__ __align__(16) uint8_t s_win[kExactThreads + kExactChunk];
    __shared__ __align__(16) uint8_t s_q[kExactChunk];
    __shared__ double s_w[4];
    __shared__ Cand s_part[kExactThreads / 32];

    const int tid = threadIdx.x;
    for (int k = tid; k < kSymbols * kRowPad; k += kExactThreads) {
        uint32_t code = T.code[k / kRowPad][k % kRowPad];
        uint32_t cls = code & 3u, rank = code >> 2;
        uint64_t e = uint64_t(rank) << 60;
        if (EXACT) { if (cls) e |= 1ull << (20 * (cls - 1)); }
        else e |= cls;
        s_tab[k] = e;
    }
    if (tid < 4) s_w[tid] = T.wcls[tid];
    if (blockIdx.x == 0 && tid == 0) *P.cand_count = 0;             // first kernel of the chain: only k_finish adds to it
    __syncthreads();

    for (int tile_id = blockIdx.x; tile_id < G.total_tiles; tile_id += gridDim.x) {
        const int q = query_of_tile(P.tile_start, G.nq, tile_id, G.tiles_per_query);
        const QueryGeom qg = query_geom(G, P.qoff, P.tile_start, q);
        const int t = tile_id - qg.tile0;
        const int64_t qbeg = qg.qbeg;
        const int len2 = qg.len2;
        const int64_t first = G.last >= 0 ? G.first : 0;
        const int64_t last = G.last >= 0 ? G.last : G.len1 - len2 + 1;
        const int64_t tb = tile_base(first) + int64_t(t) * G.tile;
        const int64_t n0 = tb > first ? tb : first;
        const int64_t n1 = (tb + G.tile) < last ? (tb + G.tile) : last;

        Cand mine{ kKeyNone, 0x7FFFFFFF };
        for (int64_t sub = n0; sub < n1; sub += kExactThreads) {
            const int64_t n = sub + tid;
            const bool valid = n < n1;
            uint64_t acc = 0;          // EXACT: three 20-bit counts (':' '.' '_')
            double total = 0.0;        // !EXACT: the reference's running sum
            uint32_t best_hi = 0;      // rank in bits 28-31
            for (int c0 = 0; c0 < len2; c0 += kExactChunk) {
                const int cl = (len2 - c0) < kExactChunk ? (len2 - c0) : kExactChunk;
                __syncthreads();
                for (int k = tid; k < cl; k += kExactThreads) {
                    uint32_t s = symbol_of(P.seq2s[qbeg + c0 + k]);
                    if (s == 0xFFu) { report_bad_symbol(P); s = 0; }
                    s_q[k] = uint8_t(s);
                }
                for (int k = tid; k < kExactThreads + cl - 1; k += kExactThreads) {
                    const int64_t p = sub + c0 + k;
                    uint32_t s = 0;
                    if (p < G.len1) {
                        s = symbol_of(P.seq1[p]);
                        if (s == 0xFFu) { report_bad_symbol(P); s = 0; }
                    }
                    s_win[k] = uint8_t(s);
                }
                __syncthreads();
                if (valid) {
                    const uint8_t* w = s_win + tid;
                    if (EXACT) {
#pragma unroll 4
                        for (int i = 0; i < cl; i++) {
                            uint64_t e = s_tab[uint32_t(s_q[i]) * kRowPad + w[i]];
                            acc += e & 0x0FFFFFFFFFFFFFFFull;
                            best_hi = max(best_hi, uint32_t(e >> 32));
                        }
                    } else {
#pragma unroll 4
                        for (int i = 0; i < cl; i++) {
                            uint64_t e = s_tab[uint32_t(s_q[i]) * kRowPad + w[i]];
                            total += s_w[uint32_t(e) & 3u];
                            best_hi = max(best_hi, uint32_t(e >> 32));
                        }
                    }
                }
            }
            if (valid) {
                const uint32_t rank = best_hi >> 28;
                int64_t key = kKeyNone;
                if (rank) {
                    if (EXACT) {
                        const int64_t c1 = int64_t(acc & 0xFFFFF), c2 = int64_t((acc >> 20) & 0xFFFFF),
                                      c3 = int64_t((acc >> 40) & 0xFFFFF);
                        const int64_t c0n = int64_t(len2) - c1 - c2 - c3;
                        key = c0n * T.kcls[0] + c1 * T.kcls[1] + c2 * T.kcls[2] + c3 * T.kcls[3] + T.kdiff[rank];
                    } else {
                        const double score = total + T.wdiff[rank];       // cpu_funcs.c:299
                        key = sortable_from_double(T.is_max ? score : -score);
                    }
                }
                if (better(key, int32_t(n), mine.key, mine.off)) { mine.key = key; mine.off = int32_t(n); }
            }
        }
        const Cand win = block_best<kExactThreads>(mine, s_part);
        if (tid == 0) {
            TileRec r;
            r.key = win.key;
            r.ub_key = kKeyNone;
            r.score = EXACT ? 0.0 : (win.key == kKeyNone ? 0.0 : (T.is_max ? 1.0 : -1.0) * double_from_sortable(win.key));
            r.offset = win.off;
            r.ub_offset = 0x7FFFFFFF;
            r.flags = kTileExact;
            r.pad = 0;
            P.tiles[tile_id] = r;
        }
    }
    pdl_launch_dependents();
}

// -------------------------------------------------------------------------------------------------
// Per-offset profile (reporting): for every offset of ONE query the reference's score
// (find_best_mutant_offset, cpu_funcs.c:257-300: sequential double + best single-substitution difference), the
// position it would mutate and the replacement letter.  One thread per offset, Seq2 staged in chunks.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kExactThreads)
k_offset_profile(const __grid_constant__ DeviceTable T, const BatchGeom G, const BatchPtrs P, double* __restrict__ scores,
                 int32_t* __restrict__ char_offsets, uint8_t* __restrict__ letters)
{
    __shared__ uint8_t s_code[kSymbols * kRowPad];
    __shared__ __align__(16) uint8_t s_win[kExactThreads + kExactChunk];
    __shared__ __align__(16) uint8_t s_q[kExactChunk];
    __shared__ double s_w[4];
    const int tid = threadIdx.x;
    for (int k = tid; k < kSymbols * kRowPad; k += kExactThreads) s_code[k] = T.code[k / kRowPad][k % kRowPad];
    if (tid < 4) s_w[tid] = T.wcls[tid];
    const int len2 = G.uniform_len2;
    const int64_t sub = G.first + int64_t(blockIdx.x) * kExactThreads;
    const int64_t n = sub + tid;
    const bool valid = n < G.last;
    double total = 0.0;
    uint32_t best_rank = 0;
    int32_t best_i = -1;
    for (int c0 = 0; c0 < len2; c0 += kExactChunk) {
        const int cl = (len2 - c0) < kExactChunk ? (len2 - c0) : kExactChunk;
        __syncthreads();
        for (int k = tid; k < cl; k += kExactThreads) {
            uint32_t c = symbol_of(P.seq2s[c0 + k]);
            if (c == 0xFFu) { report_bad_symbol(P); c = 0; }
            s_q[k] = uint8_t(c);
        }
        for (int k = tid; k < kExactThreads + cl - 1; k += kExactThreads) {
            const int64_t p = sub + c0 + k;
            uint32_t c = 0;
            if (p < G.len1) {
                c = symbol_of(P.seq1[p]);
                if (c == 0xFFu) { report_bad_symbol(P); c = 0; }
            }
            s_win[k] = uint8_t(c);
        }
        __syncthreads();
        if (valid) {
            const uint8_t* w = s_win + tid;
#pragma unroll 4
            for (int i = 0; i < cl; i++) {
                const uint32_t code = s_code[uint32_t(s_q[i]) * kRowPad + w[i]];
                total += s_w[code & 3u];
                const uint32_t r = code >> 2;
                if (r > best_rank) { best_rank = r; best_i = c0 + i; }          // strict: the first position wins ties
            }
        }
    }
    if (valid) {
        const int64_t k = n - G.first;
        if (best_rank) {
            scores[k] = __dadd_rn(__dadd_rn(total, T.wdiff[best_rank]), 0.0);
            char_offsets[k] = best_i;
            uint32_t c1 = symbol_of(P.seq1[n + best_i]), c2 = symbol_of(P.seq2s[best_i]);
            if (c1 == 0xFFu || c2 == 0xFFu) { c1 = 0; c2 = 0; }
            letters[k] = T.sub[c2][c1];
        } else {
            scores[k] = T.is_max ? -INFINITY : INFINITY;
            char_offsets[k] = -1;
            letters[k] = 0;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Top-k offsets of one query (reporting) from the per-offset profile above: one block repeats k times "the
// best offset that comes after the previous pick in the reference order" (score best first, equal scores by ascending
// offset; is_swapable, cuda_funcs.cu:290-307).  k passes over n scores in L2: a reporting path, not a hot one.
// -------------------------------------------------------------------------------------------------
constexpr int kTopkThreads = 1024;
__global__ void __launch_bounds__(kTopkThreads)
k_topk(const int is_max, const int64_t first, const int64_t n, const int k, const double* __restrict__ scores,
       const int32_t* __restrict__ char_offsets, const uint8_t* __restrict__ letters, TopkRec* __restrict__ out, int32_t* __restrict__ found)
{
    __shared__ Cand s_part[kTopkThreads / 32];
    Cand prev{ INT64_MAX, -1 };                                     // nothing picked yet: everything comes after it
    int r = 0;
    for (; r < k; r++) {
        Cand mine{ kKeyNone, 0x7FFFFFFF };
        for (int64_t i = threadIdx.x; i < n; i += kTopkThreads) {
            if (letters[i] == 0) continue;                          // no mutation possible at this offset
            const double sc = scores[i];
            const int64_t key = sortable_from_double(is_max ? sc : -sc);
            const int32_t off = int32_t(first + i);
            const bool after_prev = key < prev.key || (key == prev.key && off > prev.off);
            if (after_prev && better(key, off, mine.key, mine.off)) { mine.key = key; mine.off = off; }
        }
        const Cand win = block_best<kTopkThreads>(mine, s_part);
        if (win.key == kKeyNone) break;
        if (threadIdx.x == 0) {
            const int64_t i = int64_t(win.off) - first;
            TopkRec rec;
            rec.score = scores[i]; rec.offset = win.off; rec.char_offset = char_offsets[i]; rec.letter = letters[i]; rec.pad = 0;
            out[r] = rec;
        }
        prev = win;
        __syncthreads();
    }
    if (threadIdx.x == 0) *found = r;
}

// -------------------------------------------------------------------------------------------------
// Mutant strings of a batch (reporting; the reference builds the winner's string on the host, cpu_funcs.c:96-98):
// a copy of every query with its one substitution applied.  Thread per byte of the concatenated queries.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_emit_mutants(const BatchGeom G, const BatchPtrs P, const QueryRec* __restrict__ recs, const int64_t nbytes, uint8_t* __restrict__ out)
{
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < nbytes; e += int64_t(gridDim.x) * blockDim.x) {
        int q;
        int64_t qbeg;
        if (G.uniform_len2 > 0) { q = int(e / G.uniform_len2); qbeg = int64_t(q) * G.uniform_len2; }
        else {
            int lo = 0, hi = G.nq;                                  // qoff[lo] <= e < qoff[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (P.qoff[mid] <= e) lo = mid; else hi = mid;
            }
            q = lo; qbeg = P.qoff[lo];
        }
        const QueryRec& r = recs[q];
        uint8_t c = P.seq2s[e];
        if (r.char_offset >= 0 && e - qbeg == r.char_offset && (r.ch & 0xFF) != 0) c = uint8_t(r.ch & 0xFF);
        out[e] = c;
    }
}

// (block per query: min blocks = 1, so ptxas may spend registers on keeping the re-score loop's loads well ahead of its add chain)
template <int WPQ>
__global__ void __launch_bounds__(kFinishThreads, WPQ == 1 ? 4 : 1)
k_finish(const __grid_constant__ DeviceTable T, const BatchGeom G, const BatchPtrs P, const int scan_records)
{
    finish_body<WPQ, true, WPQ == kFinishWarps>(T, G, P, scan_records, int(blockIdx.x));
}

} // namespace

#if defined(PSA_FINISH_TRACE)
extern "C" int psa_debug_finish_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_finish_trace, sizeof(long long) * 16); }
#endif

void launch_exact_tiles(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, cudaStream_t stream)
{
    if (G.total_tiles < 1) return;
    if (T.exact) k_exact_tiles<true><<<G.total_tiles, kExactThreads, 0, stream>>>(T, G, P);
    else k_exact_tiles<false><<<G.total_tiles, kExactThreads, 0, stream>>>(T, G, P);
}

void launch_offset_profile(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, double* scores, int32_t* char_offsets,
                           uint8_t* letters, cudaStream_t stream)
{
    const int64_t n = G.last - G.first;
    if (n < 1) return;
    k_offset_profile<<<(unsigned)((n + kExactThreads - 1) / kExactThreads), kExactThreads, 0, stream>>>(T, G, P, scores, char_offsets,
                                                                                                        letters);
}

void launch_topk(int is_max, int64_t first, int64_t n, int k, const double* scores, const int32_t* char_offsets, const uint8_t* letters,
                 TopkRec* out, int32_t* found, cudaStream_t stream)
{
    k_topk<<<1, kTopkThreads, 0, stream>>>(is_max, first, n, k, scores, char_offsets, letters, out, found);
}

void launch_emit_mutants(const BatchGeom& G, const BatchPtrs& P, const QueryRec* recs, int64_t nbytes, uint8_t* out, int sm_count,
                         cudaStream_t stream)
{
    if (nbytes < 1) return;
    const int64_t blocks = std::min<int64_t>((nbytes + 255) / 256, int64_t(sm_count) * 8);
    k_emit_mutants<<<(unsigned)blocks, 256, 0, stream>>>(G, P, recs, nbytes, out);
}

void launch_finish(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, bool scan_records, cudaStream_t stream)
{
    if (G.nq < 1) return;
    if (G.nq >= 256) launch_dependent(k_finish<1>, dim3((G.nq + kFinishWarps - 1) / kFinishWarps), dim3(kFinishThreads), 0, stream, T, G, P, scan_records ? 1 : 0);
    else {
        static bool done[64];
        int dev = 0;
        cudaGetDevice(&dev);
        if (!(dev >= 0 && dev < 64 && done[dev])) {
            cudaFuncSetAttribute(k_finish<kFinishWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFinishDynBytes));
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
        launch_dependent(k_finish<kFinishWarps>, dim3(G.nq), dim3(kFinishThreads), kFinishDynBytes, stream, T, G, P, scan_records ? 1 : 0);
    }
}

} // namespace psa
