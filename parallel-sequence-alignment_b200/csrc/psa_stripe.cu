// psa_stripe.cu -- stripe mode: the whole batch in ONE launch (sm_100a).
//
// For a batch of equal-length queries whose striped window fits shared memory, this kernel replaces the chain
// k_profile -> k_scan_packed | k_scan_batch -> k_finish (psa_scan.cu, psa_kernels.cu), i.e. everything the reference does
// in fill_hashtable_gpu + calc_mutants_scores + reduction x2 (cuda_funcs.cu:149-278) per query.
//
// Striped bit planes.  A lane still owns 32 offsets as the 32 bits of a register, but they are S apart instead of
// adjacent: lane l of a query (l < S = ceil(offsets / 32)) holds offsets l, l + S, l + 2S, ...  Word p of plane row r is
// built so that bit t = fact about (row symbol r, Seq1[p + t*S]); alignment step i of lane l then needs exactly ONE
// aligned shared-memory word, W[r][l + i] -- no second word and no funnel shift (the linear layout of psa_scan.cu pays
// 2 loads + 1 SHF per plane per step: 62 of its 327 ALU instructions per 32 steps, and twice the shared-memory traffic).
// The window has S + steps words per row instead of len1 / 32 (positions repeat across words), which is why it only
// pays when it can stay resident: one persistent block per SM builds the window ONCE, straight from Seq1 (a 32x32 bit
// transpose per word and plane kind, as k_profile does for the linear layout), and then serves many queries from it.
//
// Work decomposition.  A task is Q consecutive queries = Q*S lanes laid end to end = ceil(Q*S/32) passes of one warp
// (config 3: S = 79, Q = 2 -> 5 passes, 158 of 160 lanes busy; config 5: S = 311, Q = 1 -> 10 passes).  A team of T warps
// owns a task (its passes go round the team's warps); a block holds as many teams as fit 20 warps and each team walks
// its own list of tasks, so teams drift apart and one team's set-up or epilogue overlaps another team's counting --
// there is no block-wide barrier after the window is built.  Per-(pass, query) bests meet in shared-memory slots, and
// the team's warps then finish one query each (sign counts, first position of the best rank, letter, score).
#include "psa_kernels.cuh"
#include "psa_device.cuh"
#include "psa_bitslice.h"
#include "psa_scan_core.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace psa {

#if defined(PSA_STRIPE_TRACE)
// Debug build only (make EXTRA=-DPSA_STRIPE_TRACE): per block, SM clock at the phase boundaries of k_stripe as seen by
// warp 0 (tools/stripe_trace.py prints them).  Not compiled into the product.
__device__ long long g_stripe_trace[256][10];
#define PSA_TRACE_MARK(k) do { if (threadIdx.x == 0 && blockIdx.x < 256) g_stripe_trace[blockIdx.x][k] = clock64(); } while (0)
extern "C" int psa_debug_stripe_trace(long long* out, int blocks)
{
    return (int)cudaMemcpyFromSymbol(out, g_stripe_trace, sizeof(long long) * 10 * (blocks < 256 ? blocks : 256));
}
#else
#define PSA_TRACE_MARK(k) ((void)0)
#endif

namespace {

// One group = 32 alignment steps on striped planes.
//   pw : s_cls + l * 8 (this lane's word for step 0 of the query, row 0)
//   ro : 32 byte offsets (row * Wn * 8 + step * 8), one per step; the same for every lane of a query
template <int NUP>
__device__ __forceinline__ void stripe_class_group(VCounter<NUP>& A, VCounter<NUP>& B, VCounter<NUP>& C, const char* pw,
                                                   const uint32_t* ro)
{
    uint32_t pa[5], pb[5], pn[5];
#pragma unroll
    for (int s4 = 0; s4 < 32; s4 += 4) {
        const uint4 o4 = *reinterpret_cast<const uint4*>(ro + s4);
        const uint32_t offs[4] = { o4.x, o4.y, o4.z, o4.w };
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int s = s4 + u;
            const uint2 x = *reinterpret_cast<const uint2*>(pw + offs[u]);
            vc_feed(A, pa, x.x, s);
            vc_feed(B, pb, x.y, s);
            vc_feed(C, pn, x.x & x.y, s);
        }
    }
}

// The same with two rank planes ("best rank here is the top / the second rank") riding along: their uint2 entries sit
// kStripeRankBase bytes behind the class entries, so the step's one address serves both loads and the rank planes cost one
// LDS.64 and one LOP3 per step -- no pass of their own (short queries: there is no early exit to lose).
template <int NUP>
__device__ __forceinline__ void stripe_class_rank_group(VCounter<NUP>& A, VCounter<NUP>& B, VCounter<NUP>& C, uint32_t (&racc)[2],
                                                        const char* pw, const uint32_t* ro)
{
    uint32_t pa[5], pb[5], pn[5];
#pragma unroll
    for (int s4 = 0; s4 < 32; s4 += 4) {
        const uint4 o4 = *reinterpret_cast<const uint4*>(ro + s4);
        const uint32_t offs[4] = { o4.x, o4.y, o4.z, o4.w };
        uint2 r[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int s = s4 + u;
            const uint2 x = *reinterpret_cast<const uint2*>(pw + offs[u]);
            r[u] = *reinterpret_cast<const uint2*>(pw + offs[u] + kStripeRankBase);
            vc_feed(A, pa, x.x, s);
            vc_feed(B, pb, x.y, s);
            vc_feed(C, pn, x.x & x.y, s);
        }
        racc[0] |= r[0].x | r[1].x;
        racc[0] |= r[2].x | r[3].x;
        racc[1] |= r[0].y | r[1].y;
        racc[1] |= r[2].y | r[3].y;
    }
}

//   pr : s_rnk + l * 4;  ro : byte offsets (row * Wn * 4 + step * 4)
__device__ __forceinline__ void stripe_rank_group(uint32_t& racc, const char* pr, const uint32_t* ro)
{
#pragma unroll
    for (int s4 = 0; s4 < 32; s4 += 4) {
        const uint4 o4 = *reinterpret_cast<const uint4*>(ro + s4);
        racc |= *reinterpret_cast<const uint32_t*>(pr + o4.x);
        racc |= *reinterpret_cast<const uint32_t*>(pr + o4.y);
        racc |= *reinterpret_cast<const uint32_t*>(pr + o4.z);
        racc |= *reinterpret_cast<const uint32_t*>(pr + o4.w);
    }
}

// bytes of the Seq1 symbol area: the window build gathers positions up to (Wn - 1) + 31 S (rounded up to whole 16-byte stores)
__host__ __device__ inline int stripe_seq1_span(const StripeGeom& g) { return (g.Wn + 31 * g.S + 16 + 15) & ~15; }

__device__ __forceinline__ void team_sync(int team, int team_threads)
{
    if (team_threads == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(team_threads) : "memory");
}

// bits t of lane l whose offset l + t*S is a real offset (< noff)
__device__ __forceinline__ uint32_t stripe_valid_mask(int l, int S, int64_t noff)
{
    if (l >= noff) return 0u;
    const int64_t cnt = (noff - l + S - 1) / S;
    return cnt >= 32 ? 0xFFFFFFFFu : ((1u << int(cnt)) - 1u);
}

// What a pass leaves per query segment: the segment's best (key, offset) and -- so that the finish step need not walk the
// alignment again -- the winner's vertical counters read out at its bit, and whether it met the top rank.
struct StripeSlot {
    int64_t key;
    int32_t off;
    uint32_t top;           // 1: the winner's best rank is the top rank (nranks)
    uint32_t na, nb, nc;    // N(b0), N(b1), N(b0 & b1) at the winning offset
    uint32_t pad;
};

// The finish step of one query by one warp (what finish_query_warp does for the linear kernels): sign counts, first
// position carrying the best rank, replacement letter and score of the winning offset.  Seq1 symbols and the pair table
// come from shared memory (the block staged both for the window), so the only global round trip is the query itself.
//   ro / Wn : the query's per-step row offsets ((row * Wn + step) * 8, still in shared memory from the counting) -- the top-rank
//             search reads the query's symbols out of them instead of going back to global memory
__device__ __forceinline__ void stripe_finish_query(const DeviceTable& T, const BatchPtrs& P, const uint8_t* s_seq1, const uint8_t* s_code,
                                                    int q, int64_t qbeg, int len2, const StripeSlot& w, const uint32_t* ro, int Wn)
{
    const int lane = threadIdx.x & 31;
    const Cand r{ w.key, w.off };
    QueryRec out;
    out.score = T.is_max ? -INFINITY : INFINITY;
    out.offset = -1; out.char_offset = -1; out.ch = 0; out.rank = 0;
    out.counts[0] = out.counts[1] = out.counts[2] = out.counts[3] = 0;
    if (r.key == kKeyNone) {
        warp_store_record(P.out + q, out);
        return;
    }
    if (w.top) {
        // The usual case: the winner met the top rank, and its sign-class counts came with the slot.  All that is left is
        // the FIRST position carrying that rank (strict compare in the reference, cpu_funcs.c:287-294) -- 32 positions per
        // round, normally found in the first round or two -- and the replacement letter there.
        const uint8_t* a = s_seq1 + r.off;
        const uint32_t want = uint32_t(T.nranks);
        int found = -1;
        uint32_t found_c2 = 0;                                          // the query symbol at the position found
        for (int base = 0; base < len2 && found < 0; base += 32) {
            const int i = base + lane;
            bool hit = false;
            uint32_t c2 = 0;
            if (i < len2) {
                uint32_t c1 = a[i];
                c2 = ((ro[i] >> 3) - uint32_t(i)) / uint32_t(Wn);        // the row of step i = the query's symbol (a bad one was mapped to row 0)
                if (c1 > 26u || c2 > 26u) { c1 = 0; c2 = 0; }
                hit = (uint32_t(s_code[c2 * kRowPad + c1]) >> 2) == want;
            }
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
            if (m) {
                found = base + __ffs(int(m)) - 1;
                found_c2 = __shfl_sync(0xFFFFFFFFu, c2, __ffs(int(m)) - 1);
            }
        }
        PSA_CHECK(found >= 0);
        if (found >= 0) {                                               // warp-uniform; lane 0's record is the one stored
            uint32_t c1 = a[found], c2 = found_c2;
            if (c1 > 26u) { c1 = 0; c2 = 0; }
            const int64_t n1 = int64_t(w.na) - w.nc, n2 = int64_t(w.nb) - w.nc, n3 = w.nc;
            out.offset = r.off;
            out.char_offset = found;
            out.ch = T.sub[c2][c1];
            out.rank = int(want);
            out.counts[0] = len2 - n1 - n2 - n3; out.counts[1] = n1; out.counts[2] = n2; out.counts[3] = n3;
            double sc = 0.0;
#pragma unroll
            for (int c = 0; c < 4; c++) sc = __dadd_rn(sc, __dmul_rn(double(out.counts[c]), T.wcls[c]));     // exact (psa_table.cpp)
            out.score = __dadd_rn(__dadd_rn(sc, T.wdiff[want]), 0.0);
            warp_store_record(P.out + q, out);
            return;
        }
    }
    const uint8_t* a = s_seq1 + r.off;
    const uint8_t* b = P.seq2s + qbeg;
    int cnt[4] = { 0, 0, 0, 0 };
    unsigned long long pos = 0ull;        // (rank << 32) | ~i  -> max = best rank, then lowest i
    for (int base = lane; base < len2; base += 8 * 32) {
        uint8_t vb[8];
#pragma unroll
        for (int u = 0; u < 8; u++) vb[u] = (base + u * 32) < len2 ? b[base + u * 32] : uint8_t('A');
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int i = base + u * 32;
            if (i < len2) {
                uint32_t c1 = a[i], c2 = symbol_of(vb[u]);
                if (c1 > 26u || c2 == 0xFFu) { c1 = 0; c2 = 0; }                           // a bad symbol: the batch is rejected anyway
                const uint32_t code = s_code[c2 * kRowPad + c1];
                cnt[code & 3u]++;
                const unsigned long long pp = (uint64_t(code >> 2) << 32) | uint32_t(~uint32_t(i));
                pos = pp > pos ? pp : pos;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, pos, d);
        pos = o > pos ? o : pos;
#pragma unroll
        for (int c = 0; c < 4; c++) cnt[c] += __shfl_xor_sync(0xFFFFFFFFu, cnt[c], d);
    }
    {                                                                   // every lane holds the reduced values; lane 0's record is stored
        const int rank = int(pos >> 32);
        const int i = pos ? int(~uint32_t(pos)) : 0;                    // pos == 0: rank 0, the record is overwritten below
        uint32_t c1 = a[i], c2 = symbol_of(b[i]);
        if (c1 > 26u || c2 == 0xFFu) { c1 = 0; c2 = 0; }
        out.offset = r.off;
        out.char_offset = i;
        out.ch = T.sub[c2][c1];
        out.rank = rank;
        double sc = 0.0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            out.counts[c] = cnt[c];
            sc = __dadd_rn(sc, __dmul_rn(double(cnt[c]), T.wcls[c]));          // exact (psa_table.cpp), same as k_finish
        }
        out.score = __dadd_rn(__dadd_rn(sc, T.wdiff[rank]), 0.0);
        if (rank <= 0) { out.offset = -1; out.char_offset = -1; out.ch = 0; out.score = T.is_max ? -INFINITY : INFINITY; }
        warp_store_record(P.out + q, out);
    }
}

// One pass over the alignment of query q at offset `off` by one warp: sign counts and the best (rank, lowest i), the same on
// every lane.  Seq1 symbols and the pair table come from shared memory; the query is read from global memory in rounds of 8
// loads per lane.
__device__ __forceinline__ void stripe_walk(const BatchPtrs& P, const uint8_t* s_seq1, const uint8_t* s_code, int64_t qbeg, int len2, int off,
                                            int (&cnt)[4], int& rank, int& first_i)
{
    const int lane = threadIdx.x & 31;
    const uint8_t* a = s_seq1 + off;
    const uint8_t* b = P.seq2s + qbeg;
    cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
    unsigned long long pos = 0ull;        // (rank << 32) | ~i  -> max = best rank, then lowest i
    for (int base = lane; base < len2; base += 8 * 32) {
        uint8_t vb[8];
#pragma unroll
        for (int u = 0; u < 8; u++) vb[u] = (base + u * 32) < len2 ? b[base + u * 32] : uint8_t('A');
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int i = base + u * 32;
            if (i < len2) {
                uint32_t c1 = a[i], c2 = symbol_of(vb[u]);
                if (c1 > 26u || c2 == 0xFFu) { c1 = 0; c2 = 0; }                           // a bad symbol: the batch is rejected anyway
                const uint32_t code = s_code[c2 * kRowPad + c1];
                cnt[code & 3u]++;
                const unsigned long long pp = (uint64_t(code >> 2) << 32) | uint32_t(~uint32_t(i));
                pos = pp > pos ? pp : pos;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, pos, d);
        pos = o > pos ? o : pos;
#pragma unroll
        for (int c = 0; c < 4; c++) cnt[c] += __shfl_xor_sync(0xFFFFFFFFu, cnt[c], d);
    }
    rank = int(pos >> 32);
    first_i = int(~uint32_t(pos));
}

// the record of a query at which no offset has a possible mutation (what the reference returns then: cuda_funcs.cu:143-145)
__device__ __forceinline__ void stripe_emit_none(const DeviceTable& T, const BatchPtrs& P, int q)
{
    QueryRec out;
    out.score = T.is_max ? -INFINITY : INFINITY;
    out.offset = -1; out.char_offset = -1; out.ch = 0; out.rank = 0;
    out.counts[0] = out.counts[1] = out.counts[2] = out.counts[3] = 0;
    warp_store_record(P.out + q, out);
}

// the result record from a finished walk (the walk's results are the same on every lane; the warp stores lane 0's record)
__device__ __forceinline__ void stripe_emit(const DeviceTable& T, const BatchPtrs& P, const uint8_t* s_seq1, int q, int64_t qbeg, int off,
                                            const int (&cnt)[4], int rank, int first_i)
{
    QueryRec out;
    const int fi = first_i < 0 ? 0 : first_i;                           // rank 0 (no mutation anywhere): the record is overwritten below
    uint32_t c1 = s_seq1[off + fi], c2 = symbol_of(P.seq2s[qbeg + fi]);
    if (c1 > 26u || c2 == 0xFFu) { c1 = 0; c2 = 0; }
    out.offset = off;
    out.char_offset = first_i;
    out.ch = T.sub[c2][c1];
    out.rank = rank;
    double sc = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        out.counts[c] = cnt[c];
        sc = __dadd_rn(sc, __dmul_rn(double(cnt[c]), T.wcls[c]));              // exact (psa_table.cpp), same as k_finish
    }
    out.score = __dadd_rn(__dadd_rn(sc, T.wdiff[rank]), 0.0);
    if (rank <= 0) { out.offset = -1; out.char_offset = -1; out.ch = 0; out.score = T.is_max ? -INFINITY : INFINITY; }
    warp_store_record(P.out + q, out);
}

// NB : counter planes (len2 < 2^NB), RANKPASS : a rank plane is read (K = 1 and the top rank is not derivable),
// K : rank planes tracked (0 or 1), DR : top-rank bit derived from the class counts; keys are always bit-sliced;
// SEQ : teams of one warp (T == 1) -- the warp owns all passes of its task and keeps its running best in bit planes
template <int NB, int K, bool DR, bool SEQ>
__global__ void __launch_bounds__(stripe_threads(NB), 1)
k_stripe(const BatchGeom G, const BatchPtrs P, const StripeGeom SG, const __grid_constant__ SlicedPlan SP)
{
    // the resolved table comes from global memory with coalesced loads into shared memory (as a 3.7 KB by-value parameter it made
    // the launch heavier and its per-thread reads of the constant bank serialised)
    // The block starts with three cold reads -- the table, Seq1, the first task's queries -- that do not depend on each other:
    // all three are issued before the first is waited for (one trip to L2 / HBM instead of three).
    __shared__ __align__(16) DeviceTable s_table;
    static_assert(sizeof(DeviceTable) % 16 == 0, "DeviceTable is copied in 16-byte pieces");
    constexpr int kTableVecs = int(sizeof(DeviceTable) / 16);
    static_assert(kTableVecs <= stripe_threads(NB), "one 16-byte piece of the table per thread");
    uint4 table_piece = make_uint4(0u, 0u, 0u, 0u);
    if (int(threadIdx.x) < kTableVecs)
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(table_piece.x), "=r"(table_piece.y), "=r"(table_piece.z), "=r"(table_piece.w)
                     : "l"(reinterpret_cast<const uint4*>(P.table) + threadIdx.x));
    const DeviceTable& T = s_table;
    constexpr int NUP = NB - 5;
    constexpr bool kRankPass = K > 0 && !DR;                        // rank planes are read (not derived from the class counts)
    constexpr bool kRank2 = kRankPass && K == 2;                    // two planes as uint2 at the class pitch (no second offset vector)
    constexpr bool kRor = kRankPass && !kRank2;                     // one plane at a 4-byte pitch: its own offset vector
    extern __shared__ __align__(128) unsigned char smem[];
    const int Wn = SG.Wn, S = SG.S, steps = SG.steps;
    unsigned char* s_cls = smem;                                                            // uint2 [28][Wn]
    unsigned char* s_rnk = s_cls + (K == 2 && !DR ? size_t(kStripeRankBase) : size_t(kPlaneRows) * Wn * 8);   // uint32 [28][Wn], or uint2 [28][Wn] at the fixed base
    unsigned char* s_seq1 = s_rnk + (kRankPass ? size_t(kPlaneRows) * Wn * (kRank2 ? 8 : 4) : 0);   // symbols of Seq1
    const int seq1_span = stripe_seq1_span(SG);                                             // every position the build gathers
    uint32_t* s_ro_all = reinterpret_cast<uint32_t*>(s_seq1 + seq1_span);
    const int ro_team = SG.Q * SG.ro_stride;                                                // words per team (class offsets)
    uint32_t* s_ror_all = s_ro_all + size_t(SG.teams) * ro_team;                            // rank offsets (kRankPass)
    StripeSlot* s_slot_all = reinterpret_cast<StripeSlot*>(s_ror_all + (kRor ? size_t(SG.teams) * ro_team : 0));
    const int slots_team = SG.T * SG.Q;                                                     // [warp of the team][query of the task]
    // split passes: the warp that counts the first part of the steps leaves its counter planes here for the warp that counts the rest
    constexpr int kMergeWords = 3 * NB + 2;                                                 // A, B, C planes + two rank words
    uint32_t* s_merge = reinterpret_cast<uint32_t*>(s_slot_all + size_t(SG.teams) * slots_team);   // [split pass][word][lane]
    const uint8_t* s_code = &T.code[0][0];                                                  // the pair table (in shared memory with T)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const int len2 = G.uniform_len2;
    const int64_t noff = G.len1 - len2 + 1;
    PSA_CHECK(len2 > 0 && G.last < 0 && S * 32 >= noff && steps >= len2 && Wn >= S + steps && SG.T * SG.teams * 32 <= nthreads);

    if (blockIdx.x == 0 && tid == 0) { P.cand_count[0] = 0; P.cand_count[2] = 0; }          // statistics of the run (nothing is re-scored here)

    // ---- teams ---------------------------------------------------------------------------------------------
    const int team = warp / SG.T, tw = warp - team * SG.T;
    const int team_threads = SG.T * 32;
    uint32_t* s_ro = s_ro_all + size_t(team) * ro_team;
    uint32_t* s_ror = s_ror_all + size_t(team) * ro_team;
    StripeSlot* s_slot = s_slot_all + size_t(team) * slots_team;
    const Cand none{ kKeyNone, 0x7FFFFFFF };
    const int groups = steps >> 5;
    const int task_stride = SG.teams * int(gridDim.x);
    const int first_task = team * int(gridDim.x) + int(blockIdx.x);

    // Per-step row offsets of a task's queries (needs nothing but the queries).  The team's threads stride the flattened
    // (query, step) space; a thread's byte loads -- up to 8 per round -- are all issued before the first is used, so a round
    // costs one trip to L2 / HBM, and padding steps read the all-zero row.
    auto build_row_offsets = [&](int task) {
        const int q0 = task * SG.Q;
        const int nqt = (G.nq - q0) < SG.Q ? (G.nq - q0) : SG.Q;
        const uint8_t* src = P.seq2s + int64_t(q0) * len2;
        const int total = nqt * steps;
        bool bad = false;
        int jj = 0, st = tw * 32 + lane;                            // element tt of the flattened space, kept as (query, step)
        while (st >= steps) { st -= steps; jj++; }
        for (int base = tw * 32 + lane; base < total; base += 8 * team_threads) {
            uint8_t v[8];
            int qj[8], qs[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                qj[u] = jj; qs[u] = st;
                v[u] = (base + u * team_threads < total && st < len2) ? src[jj * len2 + st] : uint8_t('A');
                st += team_threads;
                while (st >= steps) { st -= steps; jj++; }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (base + u * team_threads < total) {
                    uint32_t row = kZeroRow;
                    if (qs[u] < len2) {
                        row = symbol_of(v[u]);
                        if (row == 0xFFu) { bad = true; row = 0; }
                    }
                    s_ro[qj[u] * SG.ro_stride + qs[u]] = (row * uint32_t(Wn) + uint32_t(qs[u])) * 8u;
                    if (kRor) s_ror[qj[u] * SG.ro_stride + qs[u]] = (row * uint32_t(Wn) + uint32_t(qs[u])) * 4u;
                }
            }
        }
        if (bad) report_bad_symbol(P);
    };
    // the first task's queries are fetched while Seq1 is on its way too (both are cold reads: one wait instead of two)
    uint4 seq1_first = make_uint4(0u, 0u, 0u, 0u);
    if (tid * 16 < G.len1)                                          // (volatile: the load must be issued HERE, not sunk to its use)
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(seq1_first.x), "=r"(seq1_first.y), "=r"(seq1_first.z), "=r"(seq1_first.w)
                     : "l"(P.seq1 + tid * 16));                     // the buffer is padded: vector loads stay inside
    // (a streamed batch's queries may not have landed yet: its first task is fetched after the window is built)
    const bool streamed = P.ready != nullptr;
    if (!streamed && team < SG.teams && first_task < SG.ntasks) build_row_offsets(first_task);
    if (tid < kTableVecs) reinterpret_cast<uint4*>(&s_table)[tid] = table_piece;
    __syncthreads();                                                // the table is in shared memory from here on
    PSA_TRACE_MARK(0);

    // ---- the striped window, built once per block ------------------------------------------------------
    {
        // Seq1 as symbol indices; everything past len1 (and any byte outside [A-Z-]) becomes 31, whose columns are all zero
        bool bad = false;
        for (int i = tid * 16; i < seq1_span; i += nthreads * 16) {
            uint4 v = seq1_first;
            if (i != tid * 16) {
                v = make_uint4(0u, 0u, 0u, 0u);
                if (i < G.len1) v = *reinterpret_cast<const uint4*>(P.seq1 + i);
            }
            const uint32_t w4[4] = { v.x, v.y, v.z, v.w };
            uint32_t o4[4];
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t o = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint32_t c = symbol_of(uint8_t(w4[w] >> (8 * b)));
                    const bool inside = (i + 4 * w + b) < G.len1;
                    if (c == 0xFFu) { bad = bad || inside; c = 31u; }
                    o |= (inside ? c : 31u) << (8 * b);
                }
                o4[w] = o;
            }
            *reinterpret_cast<uint4*>(s_seq1 + i) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
        if (bad) report_bad_symbol(P);
    }
    __syncthreads();
    PSA_TRACE_MARK(1);
    // The window has S + steps words per row but only S distinct columns.  Step 1, one task per (column p < S, plane kind): the
    // 32 stripes of column p are one 32x32 bit transpose -> word p of every row.  Step 2: the words p + S, p + 2S, ... of a
    // column hold the same stripes moved up by one (word p + kS, bit t = position p + (t + k) S), so each follows from its
    // predecessor by a one-bit shift that takes in ONE new position per row -- 2 ALU instructions per row instead of another
    // transpose.  These chains are independent per row: four tasks per (column, kind), seven rows each, so that every warp of
    // the block has a share (the transposes alone keep only S * kinds / 32 warps busy).
    {
        constexpr int nkinds = kRankPass ? 2 + K : 2;
        for (int task = tid; task < S * nkinds; task += nthreads) {
            const int kind = task / S, p = task - kind * S;
            const uint32_t* col = T.col[kind];
            uint32_t m[32];
#pragma unroll
            for (int t = 0; t < 32; t++) m[t] = col[s_seq1[p + t * S]];
            transpose32(m);
            uint32_t* dst = kind < 2 ? reinterpret_cast<uint32_t*>(s_cls) + kind : reinterpret_cast<uint32_t*>(s_rnk) + (kRank2 ? kind - 2 : 0);
            const int wstep = (kind < 2 || kRank2) ? 2 : 1;                                 // words between neighbours of a row
#pragma unroll
            for (int r = 0; r < kPlaneRows; r++) dst[(size_t(r) * Wn + p) * wstep] = m[r];
        }
        __syncthreads();
        // Plane kinds that sit side by side as uint2 entries -- the two class planes, and the two rank planes where there are
        // two -- are shifted along together: one 8-byte store per entry, and the lanes of a warp take CONSECUTIVE columns of
        // the same rows, so a store is 256 contiguous bytes (the 4-byte stores of one kind at a time were every other word of
        // four scattered rows: two to three shared-memory wavefronts each, and twice as many of them).  Seven groups of four
        // rows per column.
        constexpr int kRowsPer = 4, kRowGroups = kPlaneRows / kRowsPer;                      // 28 rows = 7 x 4
        constexpr int npairs = kRank2 ? 2 : 1;
        for (int task = tid; task < S * npairs * kRowGroups; task += nthreads) {
            const int p = task % S, rest = task / S;
            const int rg = rest % kRowGroups, pair = rest / kRowGroups;
            if (p + S >= Wn) continue;                                                      // a column with a single word
            const uint32_t* col0 = T.col[2 * pair];
            const uint32_t* col1 = T.col[2 * pair + 1];
            uint2* dst = reinterpret_cast<uint2*>(pair == 0 ? s_cls : s_rnk);
            const int r0 = rg * kRowsPer;
            uint2 m[kRowsPer];
#pragma unroll
            for (int r = 0; r < kRowsPer; r++) m[r] = dst[size_t(r0 + r) * Wn + p];
            // four words per round: the (dependent) symbol and column loads of all four are issued before the first shift
            for (int word = p + S; word < Wn; word += 4 * S) {
                uint32_t c0[4], c1[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int w = word + u * S;
                    const uint32_t sym = w < Wn ? s_seq1[w + 31 * S] : 31u;                 // the position that enters at bit 31
                    c0[u] = col0[sym] >> r0;
                    c1[u] = col1[sym] >> r0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int w = word + u * S;
                    if (w < Wn) {
#pragma unroll
                        for (int r = 0; r < kRowsPer; r++) {
                            m[r].x = __funnelshift_r(m[r].x, c0[u] >> r, 1);
                            m[r].y = __funnelshift_r(m[r].y, c1[u] >> r, 1);
                            dst[size_t(r0 + r) * Wn + w] = m[r];
                        }
                    }
                }
            }
        }
        if constexpr (kRor) {
            // the one rank plane at a 4-byte pitch: on its own
            for (int task = tid; task < S * kRowGroups; task += nthreads) {
                const int p = task % S, rg = task / S;
                if (p + S >= Wn) continue;
                const uint32_t* col = T.col[2];
                uint32_t* dst = reinterpret_cast<uint32_t*>(s_rnk);
                const int r0 = rg * kRowsPer;
                uint32_t m[kRowsPer];
#pragma unroll
                for (int r = 0; r < kRowsPer; r++) m[r] = dst[size_t(r0 + r) * Wn + p];
                for (int word = p + S; word < Wn; word += S) {
                    const uint32_t c = col[s_seq1[word + 31 * S]] >> r0;
#pragma unroll
                    for (int r = 0; r < kRowsPer; r++) {
                        m[r] = __funnelshift_r(m[r], c >> r, 1);
                        dst[size_t(r0 + r) * Wn + word] = m[r];
                    }
                }
            }
        }
    }
    __syncthreads();
    PSA_TRACE_MARK(2);

    if (team >= SG.teams) return;                                   // spare warps only helped to build the window

    // bytes of the queries up to the end of `task`: what has to have landed before the task's row offsets are built
    auto task_bytes_end = [&](int task) {
        const int qe = (task + 1) * SG.Q < G.nq ? (task + 1) * SG.Q : G.nq;
        return int64_t(qe) * len2;
    };
    if (streamed && first_task < SG.ntasks) {
        if (!stream_wait(P, task_bytes_end(first_task))) return;    // (every warp of the team gives up at the same point)
        build_row_offsets(first_task);
        team_sync(team, team_threads);
    }

    for (int task = first_task; task < SG.ntasks; task += task_stride) {
        const int q0 = task * SG.Q;
        const int nqt = (G.nq - q0) < SG.Q ? (G.nq - q0) : SG.Q;    // queries of this task
        PSA_TRACE_MARK(3);

        const int lanes_task = nqt * S;
        const int passes = (lanes_task + 31) >> 5;
        // This warp's slots of the task, one per query: best candidate this warp has seen for it (merged at every flush).
        StripeSlot* my_slot = s_slot + tw * SG.Q;
        if (lane < nqt) my_slot[lane] = StripeSlot{ kKeyNone, 0x7FFFFFFF, 0u, 0u, 0u, 0u, 0u };
        __syncwarp();
        if constexpr (SEQ) {
            // ---- one warp owns the task: running best in bit planes ---------------------------------------------------
            // Per bit position (= per stripe of the lane) the best key seen over the warp's passes of the lane's current query
            // stays in the vertical domain: a pass costs the key build plus one bit-sliced compare-select (~70 LOP3) instead of
            // arg-max + warp reduction + settling + slot traffic.  Unresolved offsets take part with their upper bounds and a
            // flag; only when a query's overall winner is such a bound AND the bound turns out not to be tight is the query
            // searched again the exact way (a few per cent of short-query batches).
            constexpr int PK = SlicedKeys<NB, K>::P;
            constexpr int kPassPlanes = 6;                          // kStripeMaxPasses = 64
            const int64_t kfloor = SP.kfl;
            uint32_t best[PK], bpass[kPassPlanes], bres = 0u, bvalid = 0u;
#pragma unroll
            for (int k = 0; k < PK; k++) best[k] = 0u;
#pragma unroll
            for (int k = 0; k < kPassPlanes; k++) bpass[k] = 0u;
            int lq = -1;                                            // query whose candidates this lane's planes hold

            // lanes whose planes belong to a query other than `jnext` hand their best candidate to that query's slot
            auto flush = [&](int jnext) {
                const bool go = lq >= 0 && lq != jnext;
                if (!__any_sync(0xFFFFFFFFu, go)) return;
                uint32_t v = 0;
                int bb = 0;
                const bool have = sliced_argmax<PK>(best, go ? bvalid : 0u, v, bb);
                int pw = 0;
#pragma unroll
                for (int k = 0; k < kPassPlanes; k++) pw |= int((bpass[k] >> bb) & 1u) << k;
                const Cand c = have ? Cand{ int64_t(v) - SP.bias, int32_t(pw * 32 + lane - lq * S + bb * S) } : none;
                const uint32_t cres = have ? (bres >> bb) & 1u : 0u;
                uint32_t todo = __ballot_sync(0xFFFFFFFFu, go);
                while (todo) {                                      // one old query at a time (warp-uniform)
                    const int jj = __shfl_sync(0xFFFFFFFFu, lq, __ffs(int(todo)) - 1);
                    const bool in = go && lq == jj;
                    const Cand wb = warp_best(in ? c : none);
                    const uint32_t own = __ballot_sync(0xFFFFFFFFu, in && wb.key != kKeyNone && c.key == wb.key && c.off == wb.off);
                    const uint32_t wres = own ? __shfl_sync(0xFFFFFFFFu, cres, __ffs(int(own)) - 1) : 0u;
                    PSA_CHECK(jj >= 0 && jj < nqt);
                    if (lane == 0 && wb.key != kKeyNone && better(wb.key, wb.off, my_slot[jj].key, my_slot[jj].off))
                        my_slot[jj] = StripeSlot{ wb.key, wb.off, 0u, wres, 0u, 0u, 0u };   // .na: 1 = the key is exact (resolved offset)
                    __syncwarp();
                    todo &= ~__ballot_sync(0xFFFFFFFFu, in);
                }
                if (go) { bvalid = 0u; bres = 0u; lq = -1; }
            };

            // counters of one pass (shared by the running-best loop and the exact re-search)
            auto count_pass = [&](int jc, int lc, uint32_t vmask, VCounter<NUP>& A, VCounter<NUP>& B, VCounter<NUP>& C, uint32_t (&racc)[K > 0 ? K : 1]) {
                const uint32_t* ro = s_ro + jc * SG.ro_stride;
#pragma unroll
                for (int k = 0; k < (K > 0 ? K : 1); k++) racc[k] = 0u;
                racc[0] = ~vmask;
                if constexpr (kRankPass && !kRank2) {
                    const uint32_t* ror = s_ror + jc * SG.ro_stride;
                    const char* pr = reinterpret_cast<const char*>(s_rnk) + size_t(lc) * 4;
                    for (int g = 0; g < groups; g++) {
                        stripe_rank_group(racc[0], pr, ror + g * 32);
                        if (__all_sync(0xFFFFFFFFu, racc[0] == 0xFFFFFFFFu)) break;
                    }
                }
                A.clear(); B.clear(); C.clear();
                const char* pw = reinterpret_cast<const char*>(s_cls) + size_t(lc) * 8;
                if constexpr (kRank2) {
                    for (int g = 0; g < groups; g++) stripe_class_rank_group<NUP>(A, B, C, racc, pw, ro + g * 32);
                } else {
                    for (int g = 0; g < groups; g++) stripe_class_group<NUP>(A, B, C, pw, ro + g * 32);
                }
                if (DR) racc[0] |= derive_top_rank<NB, NUP>(T, A, B, C);
            };

            int j = 0, l = lane;
            while (l >= S) { l -= S; j++; }
            for (int p = 0; p < passes; p++) {
                const bool lane_on = p * 32 + lane < lanes_task;
                flush(lane_on ? j : -2);
                if (lane_on) lq = j;
                const int jc = lane_on ? j : 0, lc = lane_on ? l : 0;   // idle lanes shadow lane 0 (addresses stay valid)
                const uint32_t vmask = lane_on ? stripe_valid_mask(lc, S, noff) : 0u;
                VCounter<NUP> A, B, C;
                uint32_t racc[K > 0 ? K : 1];
                count_pass(jc, lc, vmask, A, B, C, racc);
                SlicedKeys<NB, K> keys;
                keys.build(SP, A, B, C, racc);
                const uint32_t valid = vmask & ~keys.nokey;
                // new > old, MSB first; a position without an old candidate takes the new one
                uint32_t gt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
                for (int k = PK - 1; k >= 0; k--) {
                    gt |= eq & keys.acc[k] & ~best[k];
                    eq &= ~(keys.acc[k] ^ best[k]);
                }
                gt = ((gt & bvalid) | ~bvalid) & valid;
#pragma unroll
                for (int k = 0; k < PK; k++) best[k] = (gt & keys.acc[k]) | (~gt & best[k]);
                bres = (gt & keys.rmask) | (~gt & bres);
#pragma unroll
                for (int k = 0; k < kPassPlanes; k++) bpass[k] = ((p >> k) & 1) ? (bpass[k] | gt) : (bpass[k] & ~gt);
                bvalid |= valid;
                l += 32;
                while (l >= S) { l -= S; j++; }
            }
            flush(-2);
            // ---- finish the task's queries ---------------------------------------------------------------------------
            for (int jj = 0; jj < nqt; jj++) {
                const StripeSlot r = my_slot[jj];
                const int q = q0 + jj;
                const int64_t qbeg = int64_t(q) * len2;
                if (r.key == kKeyNone) { stripe_emit_none(T, P, q); continue; }
                int cnt[4], rank, first_i, off = r.off;
                stripe_walk(P, s_seq1, s_code, qbeg, len2, off, cnt, rank, first_i);
                if (!r.na && (rank <= 0 || T.kdiff[rank] != kfloor)) {
                    // The winner is an unresolved offset and its bound was not tight: its true key is lower, and candidates the
                    // planes dropped in its favour may beat it.  Search this one query again the exact way (per pass: resolved
                    // arg-max, then settle every unresolved offset that could still win -- what the linear kernels do).
                    Cand lb = none;
                    const int plo = (jj * S) >> 5, phi = ((jj + 1) * S - 1) >> 5;
                    for (int p = plo; p <= phi; p++) {
                        const int f = p * 32 + lane;
                        const bool on = f < lanes_task && f >= jj * S && f < (jj + 1) * S;
                        const int lc = on ? f - jj * S : 0;
                        const uint32_t vmask = on ? stripe_valid_mask(lc, S, noff) : 0u;
                        VCounter<NUP> A, B, C;
                        uint32_t racc[K > 0 ? K : 1];
                        count_pass(jj, lc, vmask, A, B, C, racc);
                        SlicedKeys<NB, K> keys;
                        keys.build(SP, A, B, C, racc);
                        Cand mine = none, ub = none;
                        const uint32_t umask = keys.scan(vmask, lc, mine, ub, S);
                        if (on && better(mine.key, mine.off, lb.key, lb.off)) lb = mine;
                        const Cand wb = settle_unresolved(T, P, keys, on ? lb : none, on ? ub : none, on ? umask : 0u, lc, qbeg, len2, S);
                        if (on && wb.key != kKeyNone && wb.off % S == lc && better(wb.key, wb.off, lb.key, lb.off)) lb = wb;
                    }
                    const Cand fin = warp_best(lb);
                    if (fin.key == kKeyNone) { stripe_emit_none(T, P, q); continue; }
                    off = fin.off;
                    stripe_walk(P, s_seq1, s_code, qbeg, len2, off, cnt, rank, first_i);
                }
                stripe_emit(T, P, s_seq1, q, qbeg, off, cnt, rank, first_i);
            }
            __syncwarp();                                           // this task's row offsets and slots are consumed
            if (task + task_stride < SG.ntasks) {
                if (!stream_wait(P, task_bytes_end(task + task_stride))) return;
                build_row_offsets(task + task_stride);
            }
            __syncwarp();
            continue;
        }
        // Lane-local running best of the lane's current query, carried across this warp's passes: (key, offset), the pass
        // that produced it, and a warp-uniform floor -- the best RESOLVED key seen so far in the query, as the biased 32-bit
        // value of the bit-sliced keys -- under which nothing needs a second look.
        Cand lbest = none;
        int lpass = -1;
        uint32_t lfloor = 0u;
        // Units of work: whole passes, or -- for the last `split` passes of a task in a split plan -- the two parts of a pass
        // (steps [0, gh) and [gh, steps)), each on a warp of its own: the first part's warp leaves its counters in shared
        // memory and is done; the second part's warp adds them to its own and carries on as if it had counted every step.
        const int split = SG.split, full = SG.passes - split;
        const int gh = groups >= 8 ? groups / 2 + 1 : (groups + 1) / 2;        // the second part also runs the pass's epilogue
        for (int u = tw; u < (split ? SG.T : passes); u += SG.T) {
            int p = u, part = 0;                                    // part: 0 whole pass, 1 first steps, 2 the rest
            if (split && u >= full) { p = full + ((u - full) >> 1); part = 1 + ((u - full) & 1); }
            if (p >= passes) continue;                              // (a short last task: both parts skip)
            const int g_begin = part == 2 ? gh : 0, g_end = part == 1 ? gh : groups;
            const int f = p * 32 + lane;
            const bool lane_on = f < lanes_task;
            const int j = lane_on ? f / S : 0, l = lane_on ? f - j * S : 0;        // idle lanes shadow lane 0 (addresses stay valid)
            const uint32_t vmask = lane_on ? stripe_valid_mask(l, S, noff) : 0u;
            const uint32_t* ro = s_ro + j * SG.ro_stride;
            uint32_t racc[K > 0 ? K : 1];
#pragma unroll
            for (int k = 0; k < (K > 0 ? K : 1); k++) racc[k] = 0u;
            racc[0] = ~vmask;                                       // offsets outside the range count as saturated
            if constexpr (kRankPass && !kRank2) {
                const uint32_t* ror = s_ror + j * SG.ro_stride;
                const char* pr = reinterpret_cast<const char*>(s_rnk) + size_t(l) * 4;
                for (int g = g_begin; g < g_end; g++) {
                    stripe_rank_group(racc[0], pr, ror + g * 32);
                    if (__all_sync(0xFFFFFFFFu, racc[0] == 0xFFFFFFFFu)) break;
                }
            }
            VCounter<NUP> A, B, C;
            A.clear(); B.clear(); C.clear();
            const char* pw = reinterpret_cast<const char*>(s_cls) + size_t(l) * 8;
            if constexpr (kRank2) {
                for (int g = g_begin; g < g_end; g++) stripe_class_rank_group<NUP>(A, B, C, racc, pw, ro + g * 32);
            } else {
                for (int g = g_begin; g < g_end; g++) stripe_class_group<NUP>(A, B, C, pw, ro + g * 32);
            }
            if (part != 0) {
                // named barrier 2 + j pairs the two warps of split pass j (barrier 1 is this team's; split plans have one team)
                uint32_t* mg = s_merge + size_t(p - full) * kMergeWords * 32 + lane;
                const int bar = 2 + (p - full);
                if (part == 1) {
#pragma unroll
                    for (int k = 0; k < NB; k++) { mg[k * 32] = A.plane(k); mg[(NB + k) * 32] = B.plane(k); mg[(2 * NB + k) * 32] = C.plane(k); }
                    mg[3 * NB * 32] = racc[0];
                    if constexpr (K > 1) mg[(3 * NB + 1) * 32] = racc[1];
                    __threadfence_block();
                    asm volatile("bar.arrive %0, 64;" ::"r"(bar) : "memory");
                    continue;                                       // nothing of this pass is left for this warp
                }
                asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
                uint32_t a[NB], b[NB], c[NB], o[NB];
#pragma unroll
                for (int k = 0; k < NB; k++) { a[k] = A.plane(k); b[k] = B.plane(k); c[k] = C.plane(k); }
#pragma unroll
                for (int k = 0; k < NB; k++) o[k] = mg[k * 32];
                sliced_add_scaled<NB, NB>(a, o, 1);
#pragma unroll
                for (int k = 0; k < NB; k++) o[k] = mg[(NB + k) * 32];
                sliced_add_scaled<NB, NB>(b, o, 1);
#pragma unroll
                for (int k = 0; k < NB; k++) o[k] = mg[(2 * NB + k) * 32];
                sliced_add_scaled<NB, NB>(c, o, 1);
#pragma unroll
                for (int k = 0; k < NB; k++) {
                    if (k < 5) { A.low[k] = a[k]; B.low[k] = b[k]; C.low[k] = c[k]; }
                    else { A.up[k - 5 < NUP ? k - 5 : 0] = a[k]; B.up[k - 5 < NUP ? k - 5 : 0] = b[k]; C.up[k - 5 < NUP ? k - 5 : 0] = c[k]; }
                }
                // rank words: "saturated outside the range" is in both parts' words, so OR is the union of what the steps met
                racc[0] |= mg[3 * NB * 32];
                if constexpr (K > 1) racc[1] |= mg[(3 * NB + 1) * 32];
            }
            if (DR) racc[0] |= derive_top_rank<NB, NUP>(T, A, B, C);
            PSA_TRACE_MARK(4);

            // ---- keys of the lane's 32 offsets; the lane's best joins its running best ------------------------------
            SlicedKeys<NB, K> keys;                                 // stripe mode is only entered when the bit-sliced keys apply
            keys.build(SP, A, B, C, racc);
            const int jlo = (p * 32) / S;
            const int jhi = ((p * 32 + 31) / S) < (nqt - 1) ? ((p * 32 + 31) / S) : (nqt - 1);
            // One arg-max over ALL valid offsets, resolved keys and unresolved bounds alike.  Resolved winner: it is the lane's
            // candidate and no unresolved offset of the lane can beat it.  Unresolved winner below the floor: nothing in the
            // lane can matter.  Only an unresolved winner at or above the floor needs the exact treatment below.
            uint32_t v = 0;
            int bbit = 0;
            const bool have = sliced_argmax<SlicedKeys<NB, K>::P>(keys.acc, vmask & ~keys.nokey, v, bbit);
            const bool res = have && ((keys.rmask >> bbit) & 1u);
            {
                // raise the floor of the lane's query with this pass's resolved lane maxima BEFORE anything is judged by it
                // (lanes of one query share one value; a pass that straddles queries reduces each segment on its own)
                const uint32_t mine_v = res ? v : 0u;
                for (int jj = jlo; jj <= jhi; jj++) {               // warp-uniform; one segment unless the pass straddles queries
                    const bool in = lane_on && j == jj;
                    const uint32_t seg = __ballot_sync(0xFFFFFFFFu, in);
                    if (in) {
                        const uint32_t fl = __reduce_max_sync(seg, mine_v);
                        lfloor = fl > lfloor ? fl : lfloor;
                    }
                }
            }
            const bool pend = have && !res && v >= lfloor;
            if (__any_sync(0xFFFFFFFFu, pend)) {
                // exact per query segment, as the linear kernels do it: resolved best, then settle every unresolved offset whose
                // bound could still beat the segment's best (the running bests take part, so settling stops early)
                Cand mine = none, ub = none;
                const uint32_t umask = keys.scan(vmask, l, mine, ub, S);
                if (lane_on && better(mine.key, mine.off, lbest.key, lbest.off)) { lbest = mine; lpass = p; }
                for (int jj = jlo; jj <= jhi; jj++) {               // warp-uniform
                    const bool in = lane_on && j == jj;
                    const Cand wb = settle_unresolved(T, P, keys, in ? lbest : none, in ? ub : none, in ? umask : 0u, l,
                                                      int64_t(q0 + jj) * len2, len2, S);
                    // a settled offset that won is carried on by the lane that owns it (it is an offset of this pass)
                    if (in && wb.key != kKeyNone && wb.off % S == l && better(wb.key, wb.off, lbest.key, lbest.off)) { lbest = wb; lpass = p; }
                }
            } else if (res) {
                const int64_t key = int64_t(v) - keys.bias;
                const int32_t off = l + bbit * S;
                if (better(key, off, lbest.key, lbest.off)) { lbest.key = key; lbest.off = off; lpass = p; }
            }
            // ---- flush: lanes whose query ends with this pass hand their running best to the warp's slot of that query ----
            const int pn = split ? passes : p + SG.T;               // (split plans: one unit per warp, every lane leaves)
            const bool leaving = lane_on && (pn >= passes || (pn * 32 + lane) >= lanes_task || (pn * 32 + lane) / S != j);
            if (__any_sync(0xFFFFFFFFu, leaving)) {
                for (int jj = jlo; jj <= jhi; jj++) {               // warp-uniform
                    const bool in = leaving && j == jj;
                    if (!__any_sync(0xFFFFFFFFu, in)) continue;
                    const Cand wb = warp_best(in ? lbest : none);
                    PSA_CHECK(jj < SG.Q);
                    if (wb.key != kKeyNone && better(wb.key, wb.off, my_slot[jj].key, my_slot[jj].off)) {
                        // the lane that owns the winner writes the slot; if the winner comes from THIS pass its counters are
                        // still in registers: read them out at the winning bit (offset = l + t S) for the finish step
                        if (in && lbest.key == wb.key && lbest.off == wb.off) {
                            StripeSlot sl{ wb.key, wb.off, 0u, 0u, 0u, 0u, 0u };
                            if (lpass == p && wb.off % S == l) {
                                const int t = (wb.off - l) / S;
                                uint32_t na = 0, nb = 0, nc = 0;
#pragma unroll
                                for (int k = 0; k < NB; k++) {
                                    na |= ((A.plane(k) >> t) & 1u) << k;
                                    nb |= ((B.plane(k) >> t) & 1u) << k;
                                    nc |= ((C.plane(k) >> t) & 1u) << k;
                                }
                                sl.top = K > 0 ? (racc[0] >> t) & (vmask >> t) & 1u : 0u;
                                sl.na = na; sl.nb = nb; sl.nc = nc;
                            }
                            my_slot[jj] = sl;
                        }
                    }
                    __syncwarp();
                }
                if (leaving) { lbest = none; lpass = -1; lfloor = 0u; }
            }
        }
        PSA_TRACE_MARK(5);
        team_sync(team, team_threads);                              // slots complete; nobody reads this task's row offsets any more
        PSA_TRACE_MARK(6);
        if (task + task_stride < SG.ntasks) {                       // in flight while the queries are finished
            if (!stream_wait(P, task_bytes_end(task + task_stride))) return;
            build_row_offsets(task + task_stride);
        }
        // finish: one warp per query of the task -- best over the team's warps, then the record
        for (int jj = tw; jj < nqt; jj += SG.T) {
            // best over the team's warps: lane w looks at warp w's slot (teams have at most 20 warps), one warp reduction
            const Cand mine = lane < SG.T ? Cand{ s_slot[lane * SG.Q + jj].key, s_slot[lane * SG.Q + jj].off } : none;
            const Cand wb = warp_best(mine);
            const uint32_t own = __ballot_sync(0xFFFFFFFFu, lane < SG.T && mine.key == wb.key && mine.off == wb.off);
            const StripeSlot r = s_slot[(own ? __ffs(int(own)) - 1 : 0) * SG.Q + jj];
            PSA_CHECK(q0 + jj < G.nq);
            stripe_finish_query(T, P, s_seq1, s_code, q0 + jj, int64_t(q0 + jj) * len2, len2, r, s_ro + jj * SG.ro_stride, Wn);
        }
        PSA_TRACE_MARK(7);
        team_sync(team, team_threads);                              // next task's row offsets are in place, this task's slots are consumed
    }
}

size_t stripe_smem_bytes(const StripeGeom& g, int64_t len1, int rank_planes_read)
{
    if (rank_planes_read == 2 && size_t(kPlaneRows) * g.Wn * 8 > size_t(kStripeRankBase)) return size_t(1) << 30;    // class window beyond the fixed rank base
    size_t b = rank_planes_read == 2 ? size_t(kStripeRankBase) + size_t(kPlaneRows) * g.Wn * 8 : size_t(kPlaneRows) * g.Wn * (8 + (rank_planes_read == 1 ? 4 : 0));
    b += size_t(std::max<int64_t>(stripe_seq1_span(g), (len1 + 15) & ~int64_t(15)));
    b += size_t(g.teams) * g.Q * g.ro_stride * 4 * (rank_planes_read == 1 ? 2 : 1);
    b += size_t(g.teams) * g.T * g.Q * sizeof(StripeSlot);
    b += size_t(g.split) * (3 * 10 + 2) * 32 * 4;                   // merge buffers of the split passes (NB <= 10)
    return b;
}

} // namespace

// Shape of a stripe-mode launch for nq queries of len2 symbols against len1, or ok = 0 when the mode does not apply
// (window beyond shared memory, or so few lanes per query that whole warps would idle).
StripeGeom stripe_plan(int64_t len1, int64_t len2, int32_t nq, int rank_planes_read, int sm_count, bool force)
{
    StripeGeom g{};
    const int64_t noff = len1 - len2 + 1;
    if (noff < 1 || len2 < 1 || nq < 1 || len2 > 1023 || sm_count < 1) return g;        // NB <= 10 instantiated
    const int64_t S = (noff + 31) / 32, steps = (len2 + 31) & ~int64_t(31);
    if (S + steps > 4096) return g;
    g.S = int(S); g.steps = int(steps); g.Wn = int(S + steps); g.ro_stride = int(steps) + 4;
    // Candidates: Q queries per task (lane utilisation) x T warps per team (a task's passes in parallel or in sequence).
    // Cost model in cycles for the busiest block: a pass is `groups` unrolled 32-step groups (~265 ALU instructions at one
    // per two cycles per scheduler) plus its epilogue; a task adds row offsets, barriers and the finish.  The block takes
    // the larger of (all its work spread over 4 schedulers) and (the longest chain one team runs in sequence, stretched
    // when fewer than four warps per scheduler are counting and the ALU pipe cannot be kept full).
    // per pass: the unrolled groups plus the epilogue -- ~400 cycles when one warp owns the task (running best in bit planes),
    // ~1100 when several warps share it (arg-max, settling and slot traffic per pass)
    const double pass_groups_c = double(steps / 32) * 560.0;
    double best_cost = 0, best_busy = 0;
    bool have = false;
    g.threads = stripe_threads_for_len2(len2);
    const int warps_max = g.threads / 32;
    static const int dbg_q = std::getenv("PSA_STRIPE_Q") ? std::atoi(std::getenv("PSA_STRIPE_Q")) : 0;      // experiments only
    static const int dbg_t = std::getenv("PSA_STRIPE_T") ? std::atoi(std::getenv("PSA_STRIPE_T")) : 0;
    for (int Q = 1; Q <= kStripeMaxQ && Q <= nq; Q++) {
        if (dbg_q > 0 && Q != dbg_q) continue;
        const int passes = int((Q * S + 31) / 32);
        if (passes > kStripeMaxPasses) break;
        const int ntasks = (nq + Q - 1) / Q;
        const int blocks = std::min(sm_count, ntasks);
        const int tasks_b = (ntasks + blocks - 1) / blocks;                              // tasks of the busiest block
        for (int Tw = 1; Tw <= std::min(passes, warps_max); Tw++) {
            const int teams = Tw == 1 ? warps_max : std::min(warps_max / Tw, 15);        // named barriers 1..15
            if (teams < 1 || (dbg_t > 0 && Tw != dbg_t)) continue;
            StripeGeom c = g;
            c.Q = Q; c.passes = passes; c.T = Tw; c.teams = teams; c.ntasks = ntasks;
            if (stripe_smem_bytes(c, len1, rank_planes_read) > kStripeSmemMax) continue;
            const int ppw = (passes + Tw - 1) / Tw;                                      // passes per warp, in sequence
            const int rounds = (tasks_b + teams - 1) / teams;                            // tasks per team, in sequence
            const double busy = double(std::min(teams, tasks_b)) * Tw;                   // warps counting at the same time
            const double stretch = std::max(1.0, 0.75 + 4.0 / busy);                      // few warps per scheduler: dependent ALU chains show
            const double pass_c = pass_groups_c + (Tw == 1 ? 400.0 : 1100.0);
            const double task_c = 500.0 + (Tw > 1 ? 300.0 : 0.0);
            const double util = double(Q) * double(S) / (32.0 * passes);                 // lanes that hold offsets
            // warp w issues on scheduler w % 4: a team's passes go round its warps, hence round the schedulers -- but a one-warp
            // team's task stays on one scheduler, so there whole tasks are what is dealt out (8 192 config-5 queries, a GPU's share
            // of the batch on eight: 14 four-query tasks per SM put 4 on two schedulers and 3 on the others; 12 five-query tasks, 3 on each)
            const double spread = Tw == 1 ? std::ceil(double(tasks_b) / 4.0) * passes * pass_c : std::ceil(double(tasks_b) * passes / 4.0) * pass_c;
            const double chain = double(rounds) * (ppw * pass_c * stretch + task_c);
            const double cost = std::max(spread, chain) + (1.0 - util) * 0.5 * pass_c;   // time of the busiest block = time of the launch
            const bool better_cost = !have || cost < best_cost * 0.99;
            const bool tie = have && !better_cost && cost <= best_cost * 1.01 && busy > best_busy;
            if (better_cost || tie) {
                best_cost = have && tie ? std::min(best_cost, cost) : cost;
                best_busy = busy;
                g.Q = Q; g.passes = passes; g.T = Tw; g.teams = teams; g.ntasks = ntasks; g.split = 0;
                have = true;
            }
        }
        // Split plans: ONE task per block in ONE round, a warp per pass -- and, where the pass count is not a multiple of four, the
        // last `sp` passes cut in two along their steps, so that warp w % 4 = scheduler carries equal shares (config 3: 7 queries
        // = 18 passes = 16 whole + 2 x 2 parts on 20 warps: 4.5 passes per scheduler on every SM, instead of 5 on the 68 SMs that
        // hold four 5-pass tasks and 3.75 on the other 80).
        static const int dbg_split = std::getenv("PSA_STRIPE_SPLIT") ? std::atoi(std::getenv("PSA_STRIPE_SPLIT")) : -1;   // experiments only
        if (tasks_b == 1 && steps >= 256 && dbg_t <= 0)
            for (int sp = 1; sp <= 3 && sp <= passes; sp++) {
                if (dbg_split >= 0 && sp != dbg_split) continue;
                const int Tw = passes + sp;
                if (Tw > warps_max || Tw <= warps_max / 2 || Tw < 2) continue;      // one team, most of the block's warps
                StripeGeom c = g;
                c.Q = Q; c.passes = passes; c.T = Tw; c.teams = 1; c.ntasks = ntasks; c.split = sp;
                if (stripe_smem_bytes(c, len1, rank_planes_read) > kStripeSmemMax) continue;
                int load[4] = { 0, 0, 0, 0 };                                        // half passes per scheduler
                for (int w = 0; w < Tw; w++) load[w & 3] += w < passes - sp ? 2 : 1;
                const int worst = std::max(std::max(load[0], load[1]), std::max(load[2], load[3]));
                const double pass_c = pass_groups_c + 1100.0;
                const double util = double(Q) * double(S) / (32.0 * passes);
                const double cost = std::max(0.5 * worst * pass_c, pass_c + 800.0) + 300.0 + (1.0 - util) * 0.5 * pass_c;   // + merging
                if (!have || cost < best_cost * 0.99) {
                    best_cost = cost;
                    best_busy = Tw;
                    g.Q = Q; g.passes = passes; g.T = Tw; g.teams = 1; g.ntasks = ntasks; g.split = sp;
                    have = true;
                }
            }
    }
    if (dbg_q > 0 || dbg_t > 0) { /* experiments: whatever the overrides left */ }
    if (!have) return g;
    // lanes that idle in the last pass of a task: below ~70 % the linear kernels' packing does better
    if (!force && double(g.Q) * double(S) / (32.0 * g.passes) < 0.70) return g;
    g.smem = stripe_smem_bytes(g, len1, rank_planes_read);
    g.rank_planes_read = rank_planes_read;
    g.blocks = std::min(sm_count, g.ntasks);
    g.ok = 1;
    return g;
}

namespace {

// the constants SlicedKeys::build derives from the table, for K tracked rank planes (mirrors that function line by line)
SlicedPlan make_sliced_plan(const DeviceTable& T, int64_t len2, int K, int64_t bias)
{
    SlicedPlan S{};
    S.ka = int32_t(T.kcls[1] - T.kcls[0]);
    S.kb = int32_t(T.kcls[2] - T.kcls[0]);
    S.kc = int32_t(T.kcls[3] - T.kcls[1] - T.kcls[2] + T.kcls[0]);
    const int floor_rank = T.nranks - K;
    S.floor_none = floor_rank <= 0;
    S.floor_exact = S.floor_none || (floor_rank == 1 && !T.has_none);
    S.kfl = S.floor_none ? 0 : T.kdiff[floor_rank];
    int64_t ktop[4] = { 0, 0, 0, 0 };
    int64_t dmin = S.floor_none ? INT64_MAX : S.kfl;
    for (int k = 0; k < K; k++) {
        ktop[k] = T.kdiff[(T.nranks - k) > 0 ? (T.nranks - k) : 0];
        dmin = std::min(dmin, ktop[k]);
    }
    if (dmin == INT64_MAX) dmin = 0;
    S.c0 = uint32_t(bias + len2 * T.kcls[0] + dmin);
    for (int k = 0; k < K; k++) S.dv[k] = uint32_t(ktop[k] - dmin);
    S.dv_floor = S.floor_none ? 0u : uint32_t(S.kfl - dmin);
    S.bias = bias;
    return S;
}

template <int NB>
void launch_stripe_nb(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, const StripeGeom& SG, int K, bool derive,
                      int64_t key_bias, cudaStream_t stream)
{
    const SlicedPlan SP = make_sliced_plan(T, G.uniform_len2, K, key_bias);
    auto go = [&](auto kernel, bool (&done)[64]) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (!(dev >= 0 && dev < 64 && done[dev])) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kStripeSmemMax));
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
        kernel<<<SG.blocks, SG.threads, SG.smem, stream>>>(G, P, SG, SP);
    };
    static bool done[8][64];
    const bool seq = SG.T == 1;
    if (K == 0) { if (seq) go(k_stripe<NB, 0, false, true>, done[0]); else go(k_stripe<NB, 0, false, false>, done[1]); }
    else if (derive) { if (seq) go(k_stripe<NB, 1, true, true>, done[2]); else go(k_stripe<NB, 1, true, false>, done[3]); }
    else if (K == 1) { if (seq) go(k_stripe<NB, 1, false, true>, done[4]); else go(k_stripe<NB, 1, false, false>, done[5]); }
    else { if (seq) go(k_stripe<NB, 2, false, true>, done[6]); else go(k_stripe<NB, 2, false, false>, done[7]); }
}

} // namespace

bool stripe_derives_rank(const DeviceTable& T, int rank_planes, bool allow_derive)
{
    return rank_planes == 1 && T.top_rank_lut > 0 && (T.top_rank_lut & 1) == 0 && allow_derive;
}

// stripe mode needs the bit-sliced key epilogue (small integer keys in exact order)
bool stripe_keys_ok(const DeviceTable& T, int64_t len2)
{
    int64_t bias = 0;
    return len2 <= 1023 && sliced_key_planes(T, len2, len2 <= 127 ? 7 : 10, &bias) > 0;
}

void launch_stripe(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, int rank_planes, bool allow_derive,
                   const StripeGeom& SG, cudaStream_t stream)
{
    const int64_t len2 = G.uniform_len2;
    int64_t key_bias = 0;
    const int nb = len2 <= 127 ? 7 : 10;
    sliced_key_planes(T, len2, nb, &key_bias);
    const bool derive = stripe_derives_rank(T, rank_planes, allow_derive);
    // planes the kernel reads: none when the top rank is derived (or no plane is tracked), else what the planner reserved room for
    const int K = rank_planes == 0 ? 0 : derive ? 1 : SG.rank_planes_read;
    if (nb == 7) launch_stripe_nb<7>(T, G, P, SG, K, derive, key_bias, stream);
    else launch_stripe_nb<10>(T, G, P, SG, K, derive, key_bias, stream);
}

} // namespace psa
