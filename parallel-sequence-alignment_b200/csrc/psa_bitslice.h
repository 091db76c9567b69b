// psa_bitslice.h -- bit-sliced arithmetic used by the scan kernel (psa_scan.cu).
//
// One 32-bit word holds one bit for each of 32 consecutive offsets ("vertical" layout).  Counting how
// often a 1 appears per offset over many steps is done with carry-save adders (Harley-Seal): every
// input word costs one CSA (two LOP3) amortised, and each LOP3 serves 32 offsets at once.
// Everything here is plain integer code, compiled for both host (unit test) and device.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PSA_HD __host__ __device__ __forceinline__
#define PSA_UNROLL _Pragma("unroll")
#else
#define PSA_HD inline
#define PSA_UNROLL
#endif

namespace psa {

// sum/carry of three bit-vectors (per bit: a+b+c = sum + 2*carry)
PSA_HD void csa(uint32_t& sum, uint32_t& carry, uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t u = a ^ b;
    sum = u ^ c;
    carry = (a & b) | (u & c);
}

// Vertical counter: plane k holds bit k of 32 independent counts.
// low[0..4] are fed through a 32-step carry-save schedule, up[] receives one ripple add per 32 steps.
template <int NUP>
struct VCounter {
    uint32_t low[5];
    uint32_t up[NUP];
    PSA_HD void clear()
    {
PSA_UNROLL
        for (int k = 0; k < 5; k++) low[k] = 0;
PSA_UNROLL
        for (int k = 0; k < NUP; k++) up[k] = 0;
    }
    PSA_HD uint32_t plane(int k) const { return k < 5 ? low[k] : up[k - 5]; }
};

// Feed input word x as step s (0..31) of a 32-step group.  pend[] is scratch that lives only inside
// the group: level l holds an unpaired word of weight 2^l until its partner arrives.
// After step 31 every pend slot has been consumed and the carry out of the 16s plane has been
// rippled into up[].  With s a compile-time constant (fully unrolled caller) this is straight-line code.
template <int NUP>
PSA_HD void vc_feed(VCounter<NUP>& c, uint32_t (&pend)[5], uint32_t x, const int s)
{
    uint32_t carry = x;
PSA_UNROLL
    for (int l = 0; l < 5; l++) {
        if (((s >> l) & 1) == 0) { pend[l] = carry; return; }
        uint32_t sum, cy;
        csa(sum, cy, c.low[l], pend[l], carry);
        c.low[l] = sum;
        carry = cy;
    }
PSA_UNROLL
    for (int k = 0; k < NUP; k++) {
        const uint32_t t = c.up[k] & carry;
        c.up[k] ^= carry;
        carry = t;
    }
}

// One butterfly stage of the 32x32 bit transpose: rows k and k+J exchange J-wide column blocks.
template <int J>
PSA_HD void transpose_stage(uint32_t (&a)[32])
{
    constexpr uint32_t m = J == 16 ? 0x0000FFFFu : J == 8 ? 0x00FF00FFu : J == 4 ? 0x0F0F0F0Fu : J == 2 ? 0x33333333u : 0x55555555u;
PSA_UNROLL
    for (int k = 0; k < 32; k++) {
        if ((k & J) == 0) {
            const uint32_t t = ((a[k] >> J) ^ a[k + J]) & m;
            a[k + J] ^= t;
            a[k] ^= t << J;
        }
    }
}

// In-place 32x32 bit-matrix transpose: afterwards bit r of a[t] is what bit t of a[r] was.
PSA_HD void transpose32(uint32_t (&a)[32])
{
    transpose_stage<16>(a);
    transpose_stage<8>(a);
    transpose_stage<4>(a);
    transpose_stage<2>(a);
    transpose_stage<1>(a);
}


// ---- bit-sliced integer arithmetic on "vertical" numbers (plane j = bit j of 32 independent values) ----
constexpr int kSlicedMaxBits = 5;        // multipliers |k| < 2^5

// acc += k * x modulo 2^P, x unsigned with NX planes, |k| < 2^kSlicedMaxBits.
// Shift-and-add, one ripple-carry pass per set bit of |k| (two LOP3 per plane).  Negative k: two's complement,
// i.e. the operand is inverted (ones below the shift) and each term gets a carry-in of 1.  All plane indices
// are compile-time constants; the branches are on warp-uniform data.
template <int NX, int P>
PSA_HD void sliced_add_scaled(uint32_t (&acc)[P], const uint32_t (&x)[NX], const int k)
{
    if (k == 0) return;
    const uint32_t sm = k < 0 ? 0xFFFFFFFFu : 0u;
    const uint32_t mag = uint32_t(k < 0 ? -k : k);
    uint32_t xi[NX];
PSA_UNROLL
    for (int i = 0; i < NX; i++) xi[i] = x[i] ^ sm;
PSA_UNROLL
    for (int s = 0; s < kSlicedMaxBits; s++) {
        if ((mag >> s) & 1u) {
            uint32_t carry = sm;
PSA_UNROLL
            for (int j = s; j < P; j++) {
                const uint32_t xo = (j - s) < NX ? xi[(j - s) < NX ? (j - s) : 0] : sm;
                uint32_t sum, cy;
                csa(sum, cy, acc[j], xo, carry);
                acc[j] = sum;
                carry = cy;
            }
        }
    }
}

// largest value among the offsets in `alive` and the lowest offset holding it; returns false if alive == 0
template <int P>
PSA_HD bool sliced_argmax(const uint32_t (&acc)[P], uint32_t alive, uint32_t& value, int& bit)
{
    if (!alive) return false;
    uint32_t v = 0;
PSA_UNROLL
    for (int j = P - 1; j >= 0; j--) {
        const uint32_t t = alive & acc[j];
        const bool nz = t != 0;
        alive = nz ? t : alive;
        v = (v << 1) | (nz ? 1u : 0u);
    }
    value = v;
#if defined(__CUDA_ARCH__)
    bit = __ffs(int(alive)) - 1;
#else
    int b = 0;
    while (!((alive >> b) & 1u)) b++;
    bit = b;
#endif
    return true;
}

} // namespace psa
