// psa_scan_core.cuh -- device building blocks shared by the scan kernels (psa_scan.cu: linear bit planes staged by
// TMA; psa_stripe.cu: striped bit planes built in shared memory): the unrolled 32-step counting groups, the key builders
// (bit-sliced and transposed), in-kernel settling of unresolved offsets, and the per-query finish done by one warp.
#pragma once
#include "psa_kernels.cuh"
#include "psa_device.cuh"
#include "psa_bitslice.h"

#include <type_traits>

namespace psa {
namespace {

// -------------------------------------------------------------------------------------------------
// One group = 32 alignment steps with compile-time shift amounts 0..31.
//   pw : this lane's low word in the staged window for step 0 of the group
//   ro : 32 byte offsets (row * nwords * 8), one per step, warp-uniform
// -------------------------------------------------------------------------------------------------
template <int NUP>
__device__ __forceinline__ void class_group(VCounter<NUP>& A, VCounter<NUP>& B, VCounter<NUP>& C, const char* pw,
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
            const uint32_t off = offs[u];
            const uint2 lo = *reinterpret_cast<const uint2*>(pw + off);
            uint32_t x0 = lo.x, x1 = lo.y;
            if (s != 0) {
                const uint2 hi = *reinterpret_cast<const uint2*>(pw + off + 8);
                x0 = __funnelshift_r(lo.x, hi.x, s);
                x1 = __funnelshift_r(lo.y, hi.y, s);
            }
            vc_feed(A, pa, x0, s);
            vc_feed(B, pb, x1, s);
            vc_feed(C, pn, x0 & x1, s);
        }
    }
}

// Walk one alignment (Seq1 window `a`, query `b`): sign-class counts and the best (rank, lowest i) packed as
// (rank << 32) | ~i.  Thread t of nt; the 2U byte loads of a round are issued together, then the U table lookups.
template <int U>
__device__ __forceinline__ void walk_alignment(const uint8_t* a, const uint8_t* b, const uint8_t* code_table, int len2, int t, int nt,
                                               int (&cnt)[4], unsigned long long& pos)
{
    for (int base = t; base < len2; base += U * nt) {
        uint8_t va[U], vb[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * nt;
            va[u] = i < len2 ? a[i] : uint8_t('A');
            vb[u] = i < len2 ? b[i] : uint8_t('A');
        }
        uint32_t code[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            uint32_t c1 = symbol_of(va[u]), c2 = symbol_of(vb[u]);
            if (c1 == 0xFFu || c2 == 0xFFu) { c1 = 0; c2 = 0; }
            code[u] = __ldg(code_table + c2 * kRowPad + c1);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * nt;
            if (i < len2) {
                cnt[code[u] & 3u]++;
                const unsigned long long p = (uint64_t(code[u] >> 2) << 32) | uint32_t(~uint32_t(i));
                pos = p > pos ? p : pos;
            }
        }
    }
}

// Per-step row offsets (row * row_bytes) of alignment steps [i0, i0 + n) of the query at `src`; steps at or past len2
// read the all-zero row.  Thread t of nt; the U byte loads of a round are issued together (one global round trip).
// Returns true if a byte outside [A-Z-] was met.
template <int U>
__device__ __forceinline__ bool fill_rows(uint32_t* dst, const uint8_t* src, int i0, int n, int len2, uint32_t row_bytes, int t, int nt)
{
    bool bad = false;
    for (int base = t; base < n; base += U * nt) {
        uint8_t v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int st = base + u * nt;
            v[u] = (st < n && i0 + st < len2) ? src[i0 + st] : uint8_t('A');
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int st = base + u * nt;
            if (st < n) {
                uint32_t row = kZeroRow;
                if (i0 + st < len2) {
                    row = symbol_of(v[u]);
                    if (row == 0xFFu) { bad = true; row = 0; }
                }
                dst[st] = row * row_bytes;
            }
        }
    }
    return bad;
}

// When the top-rank bit of a pair is a function of its sign class (DeviceTable::top_rank_lut, never class 0), an offset
// met the top rank iff it met a pair of such a class, i.e. iff that class count is non-zero: N(':') = N(b0) - N(b0&b1),
// N('.') = N(b1) - N(b0&b1), N('_') = N(b0&b1).  So the bit falls out of the three counters after the scan -- no rank
// plane, no second pass, nothing per step.
template <int NB, int NUP>
__device__ __forceinline__ uint32_t derive_top_rank(const DeviceTable& T, const VCounter<NUP>& A, const VCounter<NUP>& B,
                                                    const VCounter<NUP>& C)
{
    uint32_t nz1 = 0, nz2 = 0, nz3 = 0;
#pragma unroll
    for (int k = 0; k < NB; k++) {
        const uint32_t c = C.plane(k);
        nz1 |= A.plane(k) ^ c;
        nz2 |= B.plane(k) ^ c;
        nz3 |= c;
    }
    const int lut = T.top_rank_lut;
    return ((lut & 2) ? nz1 : 0u) | ((lut & 4) ? nz2 : 0u) | ((lut & 8) ? nz3 : 0u);
}

template <int K>
__device__ __forceinline__ void rank_group(uint32_t (&racc)[K > 0 ? K : 1], const char* pw, const uint32_t* ro)
{
#pragma unroll
    for (int s4 = 0; s4 < 32; s4 += 4) {
        const uint4 o4 = *reinterpret_cast<const uint4*>(ro + s4);
        const uint32_t offs[4] = { o4.x, o4.y, o4.z, o4.w };
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int s = s4 + u;
            const uint32_t off = offs[u];                 // row * nwords * 8; rank entries are 4*K bytes wide
            if (K == 1) {
                const uint32_t l = *reinterpret_cast<const uint32_t*>(pw + off);             // rows at the 8-byte pitch (rank_pitch)
                uint32_t x = l;
                if (s != 0) x = __funnelshift_r(l, *reinterpret_cast<const uint32_t*>(pw + off + 4), s);
                racc[0] |= x;
            } else if (K == 2) {
                const uint2 l = *reinterpret_cast<const uint2*>(pw + off);
                uint2 h = l;
                if (s != 0) h = *reinterpret_cast<const uint2*>(pw + off + 8);
                racc[0] |= s ? __funnelshift_r(l.x, h.x, s) : l.x;
                racc[K > 1 ? 1 : 0] |= s ? __funnelshift_r(l.y, h.y, s) : l.y;
            } else if (K == 4) {
                const uint4 l = *reinterpret_cast<const uint4*>(pw + (off << 1));
                uint4 h = l;
                if (s != 0) h = *reinterpret_cast<const uint4*>(pw + (off << 1) + 16);
                racc[0] |= s ? __funnelshift_r(l.x, h.x, s) : l.x;
                racc[K > 1 ? 1 : 0] |= s ? __funnelshift_r(l.y, h.y, s) : l.y;
                racc[K > 2 ? 2 : 0] |= s ? __funnelshift_r(l.z, h.z, s) : l.z;
                racc[K > 3 ? 3 : 0] |= s ? __funnelshift_r(l.w, h.w, s) : l.w;
            }
        }
    }
}

__device__ __forceinline__ void take(Cand& c, int64_t key, int32_t off)
{
    if (better(key, off, c.key, c.off)) { c.key = key; c.off = off; }
}

__host__ __device__ inline int round_up4(int x) { return (x + 3) & ~3; }

__device__ __forceinline__ uint32_t valid_mask(int64_t ln0, int64_t first, int64_t last)
{
    const int64_t lo = first > ln0 ? first - ln0 : 0;
    const int64_t hi = (last - ln0) < 32 ? (last - ln0) : 32;
    if (hi <= lo || lo >= 32) return 0u;
    const uint32_t upto_hi = hi >= 32 ? 0xFFFFFFFFu : ((1u << int(hi)) - 1u);
    return upto_hi & ~((1u << int(lo)) - 1u);
}

// -------------------------------------------------------------------------------------------------
// Epilogue shared by both scan kernels: from the vertical counters of one lane (32 offsets) to per-offset
// keys.  A 32x32 bit transpose turns plane k / offset t into offset t / bit k; the key of an offset is
//   key = len2*k0 + N(b0)*(k1-k0) + N(b1)*(k2-k0) + N(b0&b1)*(k3-k1-k2+k0) + kdiff[rank]
// Offsets that met none of the K tracked rank planes only get an upper bound (kdiff[floor_rank]).
// -------------------------------------------------------------------------------------------------
template <int NB, int K, bool KEY32>
struct OffsetKeys {
    static constexpr bool kSingle = 3 * NB + K <= 32;
    using key_t = typename std::conditional<KEY32, int32_t, int64_t>::type;
    uint32_t m[32], m2[kSingle ? 1 : 32];
    key_t ka, kb, kc, kbase, kfl, ktop[K > 0 ? K : 1];
    bool floor_none, floor_exact;

    template <int NUP>
    __device__ __forceinline__ void build(const DeviceTable& T, int len2, const VCounter<NUP>& A, const VCounter<NUP>& B,
                                          const VCounter<NUP>& C, const uint32_t (&racc)[K > 0 ? K : 1], int /*planes*/,
                                          int64_t /*bias*/)
    {
        ka = key_t(T.kcls[1] - T.kcls[0]);
        kb = key_t(T.kcls[2] - T.kcls[0]);
        kc = key_t(T.kcls[3] - T.kcls[1] - T.kcls[2] + T.kcls[0]);
        kbase = key_t(int64_t(len2) * T.kcls[0]);
        // ranks nranks, nranks-1, .. nranks-K+1 are tracked; anything else is <= floor_rank
        const int floor_rank = T.nranks - K;
        floor_none = floor_rank <= 0;                                  // nothing below the planes but "no substitute"
        floor_exact = floor_none || (floor_rank == 1 && !T.has_none);
        kfl = floor_none ? key_t(0) : key_t(T.kdiff[floor_rank]);
#pragma unroll
        for (int k = 0; k < K; k++) ktop[k] = key_t(T.kdiff[(T.nranks - k) > 0 ? (T.nranks - k) : 0]);
#pragma unroll
        for (int k = 0; k < 32; k++) m[k] = 0;
#pragma unroll
        for (int k = 0; k < NB; k++) { m[k] = A.plane(k); m[NB + k] = B.plane(k); }
        if (kSingle) {
#pragma unroll
            for (int k = 0; k < NB; k++) m[(kSingle ? 2 * NB : 0) + k] = C.plane(k);
#pragma unroll
            for (int k = 0; k < K; k++) m[(kSingle ? 3 * NB : 0) + k] = racc[k];
            transpose32(m);
        } else {
            uint32_t (&mm)[32] = reinterpret_cast<uint32_t (&)[32]>(m2);
#pragma unroll
            for (int k = 0; k < 32; k++) mm[k] = 0;
#pragma unroll
            for (int k = 0; k < NB; k++) mm[k] = C.plane(k);
#pragma unroll
            for (int k = 0; k < K; k++) mm[NB + k] = racc[k];
            transpose32(m);
            transpose32(mm);
        }
    }

    // Visit the offsets in `mask`: resolved ones compete for `res`, unresolved ones for `ub` and are
    // reported in the returned mask.  (key desc, offset asc) order in both.  Bit t of the lane is offset
    // ln0 + t * stride (stride 1: linear planes; stride S: striped planes, psa_stripe.cu).
    __device__ __forceinline__ uint32_t scan(uint32_t mask, int64_t ln0, Cand& res, Cand& ub, int stride = 1) const
    {
        constexpr uint32_t kMask = (1u << NB) - 1u;
        uint32_t unresolved = 0;
        int32_t pres = INT32_MIN, pub = INT32_MIN;      // KEY32: key * 32 + (31 - t): one integer max orders (key desc, offset asc)
#pragma unroll
        for (int tt = 0; tt < 32; tt++) {
            if (!((mask >> tt) & 1u)) continue;
            const uint32_t v = m[tt];
            const uint32_t na = v & kMask, nb = (v >> NB) & kMask;
            uint32_t nc, rb;
            if (kSingle) { nc = (v >> (kSingle ? 2 * NB : 0)) & kMask; rb = K > 0 ? (v >> (kSingle ? 3 * NB : 0)) & ((1u << K) - 1u) : 0u; }
            else { nc = m2[kSingle ? 0 : tt] & kMask; rb = K > 0 ? (m2[kSingle ? 0 : tt] >> NB) & ((1u << K) - 1u) : 0u; }
            const key_t key = kbase + key_t(na) * ka + key_t(nb) * kb + key_t(nc) * kc;
            bool resolved = true;
            key_t d = kfl;
            if (K > 0 && rb) {
                d = ktop[K > 0 ? K - 1 : 0];            // lowest set plane = best rank present
#pragma unroll
                for (int k = K - 2; k >= 0; k--)
                    if (rb & (1u << k)) d = ktop[k];
            } else {
                if (floor_none) continue;               // no mutation possible at this offset
                resolved = floor_exact;
            }
            if (resolved) {
                if (KEY32) pres = max(pres, int32_t(key + d) * 32 + (31 - tt));
                else take(res, int64_t(key + d), int32_t(ln0 + tt * stride));
            } else {
                unresolved |= 1u << tt;
                if (KEY32) pub = max(pub, int32_t(key + d) * 32 + (31 - tt));
                else take(ub, int64_t(key + d), int32_t(ln0 + tt * stride));
            }
        }
        if (KEY32) {
            if (pres != INT32_MIN) take(res, int64_t(pres >> 5), int32_t(ln0 + (31 - (pres & 31)) * stride));
            if (pub != INT32_MIN) take(ub, int64_t(pub >> 5), int32_t(ln0 + (31 - (pub & 31)) * stride));
        }
        return unresolved;
    }
};

// -------------------------------------------------------------------------------------------------
// The same keys without leaving the bit-sliced domain (small integer weights, exact mode): the key planes
// are built with bit-sliced shift-and-add from the three vertical counters, the rank term is selected by
// the rank planes, and the best offset of the lane falls out of an MSB-first elimination over the planes.
// ~300 ALU ops per lane instead of ~900 for transpose + 32 scalar keys.
//   key' = key + bias  (bias makes every key' non-negative and < 2^planes; both come from the host)
// -------------------------------------------------------------------------------------------------
constexpr int sliced_planes(int nb) { return nb + 8; }     // key planes by counter width: 15 / 18 / 23

template <int NB, int K>
struct SlicedKeys {
    static constexpr int P = sliced_planes(NB);
    uint32_t acc[P];
    uint32_t rmask;      // offsets whose best rank the tracked planes determine
    uint32_t nokey;      // offsets with no possible mutation (only when some pair has no substitute)
    int64_t bias, kfl;

    template <int NUP>
    __device__ __forceinline__ void build(const DeviceTable& T, int len2, const VCounter<NUP>& A, const VCounter<NUP>& B,
                                          const VCounter<NUP>& C, const uint32_t (&racc)[K > 0 ? K : 1], int /*planes*/, int64_t bias_)
    {
        bias = bias_;
        const int ka = int(T.kcls[1] - T.kcls[0]), kb = int(T.kcls[2] - T.kcls[0]);
        const int kc = int(T.kcls[3] - T.kcls[1] - T.kcls[2] + T.kcls[0]);
        const int floor_rank = T.nranks - K;
        const bool floor_none = floor_rank <= 0;
        const bool floor_exact = floor_none || (floor_rank == 1 && !T.has_none);
        kfl = floor_none ? 0 : T.kdiff[floor_rank];
        int64_t ktop[K > 0 ? K : 1];
        int64_t dmin = floor_none ? INT64_MAX : kfl;
#pragma unroll
        for (int k = 0; k < K; k++) {
            ktop[k] = T.kdiff[(T.nranks - k) > 0 ? (T.nranks - k) : 0];
            dmin = ktop[k] < dmin ? ktop[k] : dmin;
        }
        const uint32_t c0 = uint32_t(bias + int64_t(len2) * T.kcls[0] + dmin);
#pragma unroll
        for (int j = 0; j < P; j++) acc[j] = ((c0 >> j) & 1u) ? 0xFFFFFFFFu : 0u;
        uint32_t x[NB];
#pragma unroll
        for (int k = 0; k < NB; k++) x[k] = A.plane(k);
        sliced_add_scaled<NB, P>(acc, x, ka);
#pragma unroll
        for (int k = 0; k < NB; k++) x[k] = B.plane(k);
        sliced_add_scaled<NB, P>(acc, x, kb);
#pragma unroll
        for (int k = 0; k < NB; k++) x[k] = C.plane(k);
        sliced_add_scaled<NB, P>(acc, x, kc);
        // rank term: one-hot selection masks, best tracked plane first
        uint32_t seen = 0, dpl[8];
#pragma unroll
        for (int j = 0; j < 8; j++) dpl[j] = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const uint32_t sel = racc[k] & ~seen;
            seen |= racc[k];
            const uint32_t dv = uint32_t(ktop[k] - dmin);
#pragma unroll
            for (int j = 0; j < 8; j++)
                if ((dv >> j) & 1u) dpl[j] |= sel;
        }
        if (!floor_none) {
            const uint32_t dv = uint32_t(kfl - dmin);
#pragma unroll
            for (int j = 0; j < 8; j++)
                if ((dv >> j) & 1u) dpl[j] |= ~seen;
        }
        sliced_add_scaled<8, P>(acc, dpl, 1);
        rmask = floor_exact ? 0xFFFFFFFFu : seen;
        nokey = floor_none ? ~seen : 0u;
    }

    // the same from constants the host resolved (SlicedPlan): nothing but bit-plane operations here
    template <int NUP>
    __device__ __forceinline__ void build(const SlicedPlan& S, const VCounter<NUP>& A, const VCounter<NUP>& B, const VCounter<NUP>& C,
                                          const uint32_t (&racc)[K > 0 ? K : 1])
    {
        bias = S.bias;
        kfl = S.kfl;
#pragma unroll
        for (int j = 0; j < P; j++) acc[j] = ((S.c0 >> j) & 1u) ? 0xFFFFFFFFu : 0u;
        uint32_t x[NB];
#pragma unroll
        for (int k = 0; k < NB; k++) x[k] = A.plane(k);
        sliced_add_scaled<NB, P>(acc, x, S.ka);
#pragma unroll
        for (int k = 0; k < NB; k++) x[k] = B.plane(k);
        sliced_add_scaled<NB, P>(acc, x, S.kb);
#pragma unroll
        for (int k = 0; k < NB; k++) x[k] = C.plane(k);
        sliced_add_scaled<NB, P>(acc, x, S.kc);
        uint32_t seen = 0, dpl[8];
#pragma unroll
        for (int j = 0; j < 8; j++) dpl[j] = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const uint32_t sel = racc[k] & ~seen;
            seen |= racc[k];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if ((S.dv[k] >> j) & 1u) dpl[j] |= sel;
        }
        if (!S.floor_none) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if ((S.dv_floor >> j) & 1u) dpl[j] |= ~seen;
        }
        sliced_add_scaled<8, P>(acc, dpl, 1);
        rmask = S.floor_exact ? 0xFFFFFFFFu : seen;
        nokey = S.floor_none ? ~seen : 0u;
    }

    __device__ __forceinline__ uint32_t scan(uint32_t mask, int64_t ln0, Cand& res, Cand& ub, int stride = 1) const
    {
        uint32_t v;
        int b;
        mask &= ~nokey;
        if (sliced_argmax<P>(acc, mask & rmask, v, b)) take(res, int64_t(v) - bias, int32_t(ln0 + b * stride));
        const uint32_t unresolved = mask & ~rmask;
        if (sliced_argmax<P>(acc, unresolved, v, b)) take(ub, int64_t(v) - bias, int32_t(ln0 + b * stride));
        return unresolved;
    }
};

// Exact order is required from the scan in exact mode: settle every unresolved offset whose bound could
// beat the warp's best resolved key by looking up its true best rank (one pass over the alignment, lanes
// striding i; the counts, hence the key without the difference term, are already exact).  Returns the
// warp's exact best.  All 32 lanes must call.
template <class Keys>
__device__ __forceinline__ Cand settle_unresolved(const DeviceTable& T, const BatchPtrs& P,
                                                  const Keys& keys, Cand mine, Cand ub, uint32_t umask,
                                                  int64_t ln0, int64_t qbeg, int len2, int stride = 1)
{
    Cand wbest = warp_best(mine);
    if (!__any_sync(0xFFFFFFFFu, umask != 0)) return wbest;
    uint32_t settled = 0;
    const int64_t kfloor = int64_t(keys.kfl);
    for (;;) {
        const bool could_win = ub.key != kKeyNone && !better(wbest.key, wbest.off, ub.key, ub.off);
        const uint32_t lanes = __ballot_sync(0xFFFFFFFFu, could_win);
        if (!lanes) break;
        const int L = __ffs(int(lanes)) - 1;
        const int64_t ukey = __shfl_sync(0xFFFFFFFFu, ub.key, L);
        const int32_t off = __shfl_sync(0xFFFFFFFFu, ub.off, L);
        uint32_t rmax = 0;
        for (int i = int(threadIdx.x & 31); i < len2; i += 32) {
            uint32_t c1 = symbol_of(P.seq1[off + i]), c2 = symbol_of(P.seq2s[qbeg + i]);
            if (c1 == 0xFFu || c2 == 0xFFu) { c1 = 0; c2 = 0; }
            rmax = max(rmax, uint32_t(__ldg(P.code_table + c2 * kRowPad + c1)) >> 2);
        }
        rmax = __reduce_max_sync(0xFFFFFFFFu, rmax);
        if (rmax) {
            const int64_t key = ukey - kfloor + T.kdiff[rmax];
            if (better(key, off, wbest.key, wbest.off)) { wbest.key = key; wbest.off = off; }
        }
        if (int(threadIdx.x & 31) == L) {
            // this lane's next-best unresolved offset
            settled |= 1u << int((int64_t(off) - ln0) / stride);
            Cand unused{ kKeyNone, 0x7FFFFFFF };
            ub = Cand{ kKeyNone, 0x7FFFFFFF };
            keys.scan(umask & ~settled, ln0, unused, ub, stride);
        }
    }
    return wbest;
}

// what the fused finish of k_scan does, by one warp: sign counts, first position carrying the best rank,
// replacement letter and score of the winning offset -> QueryRec
__device__ __forceinline__ void finish_query_warp(const DeviceTable& T, const BatchPtrs& P, int q, int64_t qbeg, int len2, Cand r)
{
    const int lane = threadIdx.x & 31;
    QueryRec out;
    out.score = T.is_max ? -INFINITY : INFINITY;
    out.offset = -1; out.char_offset = -1; out.ch = 0; out.rank = 0;
    out.counts[0] = out.counts[1] = out.counts[2] = out.counts[3] = 0;
    if (r.key == kKeyNone) {
        if (lane == 0) P.out[q] = out;
        return;
    }
    const uint8_t* a = P.seq1 + r.off;
    const uint8_t* b = P.seq2s + qbeg;
    int cnt[4] = { 0, 0, 0, 0 };
    unsigned long long pos = 0ull;        // (rank << 32) | ~i  -> max = best rank, then lowest i
    walk_alignment<8>(a, b, P.code_table, len2, lane, 32, cnt, pos);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, pos, d);
        pos = o > pos ? o : pos;
#pragma unroll
        for (int c = 0; c < 4; c++) cnt[c] += __shfl_xor_sync(0xFFFFFFFFu, cnt[c], d);
    }
    if (lane == 0) {
        const int rank = int(pos >> 32);
        const int i = int(~uint32_t(pos));
        uint32_t c1 = symbol_of(a[i]), c2 = symbol_of(b[i]);
        if (c1 == 0xFFu || c2 == 0xFFu) { c1 = 0; c2 = 0; }
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
        P.out[q] = out;
    }
}


// -------------------------------------------------------------------------------------------------
// Combine (slice mode): one thread per offset of a 256-offset tile adds the slices' partial counts, forms the key, and the
// block reduces to the tile's best.  Offsets that met no tracked rank plane are settled here when the order must be exact:
// those whose bound could beat the block's best walk the alignment for their true best rank (rare, and only for the few
// that matter).  Used by k_combine (psa_scan.cu, one tile per block) and k_single (psa_single.cu, tiles round the blocks
// after a grid barrier -- there the partial counts were written by other blocks of the SAME launch: ld.cg).
// -------------------------------------------------------------------------------------------------
constexpr int kCombineThreads = 256;

template <int K, bool PDL>
__device__ __forceinline__ void combine_tile(const DeviceTable& T, const BatchGeom& G, const BatchPtrs& P, const int slices, const int tile)
{
    __shared__ Cand s_part[kCombineThreads / 32];
    __shared__ int64_t s_top[kCombineThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int len2 = G.uniform_len2;
    const int64_t first = G.first, last = G.last;                   // slice mode always runs on an explicit range
    const int64_t rel = int64_t(tile) * kCombineThreads + tid;
    PSA_CHECK(tile < G.total_tiles && G.uniform_len2 > 0 && G.last >= 0);
    const int64_t n = tile_base(first) + rel;
    const bool valid = n >= first && n < last;

    int64_t key = kKeyNone;
    bool unresolved = false;
    const int floor_rank = T.nranks - K;
    const bool floor_none = floor_rank <= 0;
    const bool floor_exact = floor_none || (floor_rank == 1 && !T.has_none);
    const int64_t kfloor = floor_none ? 0 : T.kdiff[floor_rank];
    if (valid) {
        uint32_t na = 0, nb = 0, nc = 0, rb = 0;
        for (int sl = 0; sl < slices; sl++) {
            PSA_CHECK(rel < P.partial_stride);
            const uint2 v = __ldcg(P.partial + int64_t(sl) * P.partial_stride + rel);
            na += v.x & 0xFFFFu; nb += v.x >> 16; nc += v.y & 0xFFFFu; rb |= v.y >> 16;
        }
        const int64_t ka = T.kcls[1] - T.kcls[0], kb = T.kcls[2] - T.kcls[0];
        const int64_t kc = T.kcls[3] - T.kcls[1] - T.kcls[2] + T.kcls[0];
        key = int64_t(len2) * T.kcls[0] + int64_t(na) * ka + int64_t(nb) * kb + int64_t(nc) * kc;
        if (K > 0 && rb) {
            key += T.kdiff[T.nranks - (__ffs(int(rb)) - 1)];        // lowest set plane = best rank present
        } else if (floor_none) {
            key = kKeyNone;
        } else {
            key += kfloor;
            unresolved = !floor_exact;
        }
    }
    Cand mine{ unresolved ? kKeyNone : key, unresolved || key == kKeyNone ? 0x7FFFFFFF : int32_t(n) };
    Cand best = block_best<kCombineThreads>(mine, s_part);
    if (PDL) pdl_launch_dependents();
    if (T.exact) {
        // settle: any unresolved offset whose bound could beat the block's best looks up its true rank
        int again = __syncthreads_or(unresolved && !better(best.key, best.off, key, int32_t(n)));
        while (again) {
            if (unresolved && !better(best.key, best.off, key, int32_t(n))) {
                uint32_t rmax = 0;
                for (int i = 0; i < len2; i++) {
                    uint32_t c1 = symbol_of(P.seq1[n + i]), c2 = symbol_of(P.seq2s[i]);
                    if (c1 == 0xFFu || c2 == 0xFFu) { c1 = 0; c2 = 0; }
                    rmax = max(rmax, uint32_t(__ldg(P.code_table + c2 * kRowPad + c1)) >> 2);
                }
                key = rmax ? key - kfloor + T.kdiff[rmax] : kKeyNone;
                unresolved = false;
                mine = Cand{ key, key == kKeyNone ? 0x7FFFFFFF : int32_t(n) };
            }
            best = block_best<kCombineThreads>(mine, s_part);
            again = __syncthreads_or(unresolved && !better(best.key, best.off, key, int32_t(n)));
        }
    } else {
        // re-score mode: per 32-offset word an upper estimate of its keys for k_finish
        int64_t top = valid ? key : kKeyNone;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int64_t o = __shfl_xor_sync(0xFFFFFFFFu, top, d);
            top = o > top ? o : top;
        }
        if (lane == 0) {
            P.lane_keys[int64_t(tile) * (kCombineThreads / 32) + warp] = top;
            s_top[warp] = top;
        }
        __syncthreads();
    }
    if (tid == 0) {
        TileRec rec;
        rec.key = best.key; rec.offset = best.off;
        rec.ub_key = kKeyNone;
        if (!T.exact)
            for (int w = 0; w < kCombineThreads / 32; w++) rec.ub_key = s_top[w] > rec.ub_key ? s_top[w] : rec.ub_key;
        rec.ub_offset = 0x7FFFFFFF; rec.score = 0.0; rec.flags = 0; rec.pad = 0;
        P.tiles[tile] = rec;
    }
}

} // namespace
} // namespace psa
