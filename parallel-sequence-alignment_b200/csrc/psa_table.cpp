// psa_table.cpp -- host-side resolution of the per-pair table.
//
// Replaces, for the GPU path, three things the reference does per pair per offset:
//   * the 26x26 sign "hashtable" (fill_hash cpu_funcs.c:304-318, fill_hashtable_gpu
//     cuda_funcs.cu:267-278, lookup get_hashtable_sign cuda_funcs.cu:424-439),
//   * the substitution search (get_substitute .. get_substitute_by_sign_with_restrictions,
//     cuda_funcs.cu:310-421),
//   * the weight of a sign (get_weight cuda_funcs.cu:442-452).
// All are pure in (c1, c2, weights, goal); there are only 27x27 distinct inputs, so they are
// evaluated once here and shipped to the device as a 27x32 byte table + a handful of constants.
//
// It also decides how scores are ordered on the device.  The reference orders offsets by a double
// accumulated sequentially over i (cpu_funcs.c:278).  The device instead orders them by an int64
// fixed-point key computed from exact integer sign counts.  When every partial sum is exactly
// representable ("exact" mode: integer or dyadic weights, the common case) the two orders are
// identical.  Otherwise key_slack bounds how far the two can disagree, and every offset whose key
// is within that window of the best key is re-scored on the device in the reference's own
// summation order (finish_body in psa_finish.cuh), so the final order is again the reference's.
#include "psa_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace psa {

namespace {

// Conservative / semi-conservative groups (cpu_funcs.c:19-20, README "Definitions").
const char* const kConservative[] = { "NDEQ", "NEQK", "STA", "MILV", "QHRK", "NHQK", "FYW", "HY", "MILF" };
const char* const kSemiConservative[] = { "SAG", "ATV", "CSA", "SGND", "STPA", "STNK", "NEQHRK", "NDEQHK",
                                          "SNDEQK", "HFY", "FVLIM" };

struct SignMatrix {
    char s[kSymbols][kSymbols];
    SignMatrix()
    {
        bool cons[26][26] = {}, semi[26][26] = {};
        auto mark = [](const char* const* groups, size_t n, bool (*m)[26]) {
            for (size_t g = 0; g < n; g++)
                for (const char* p = groups[g]; *p; p++)
                    for (const char* q = groups[g]; *q; q++) m[*p - 'A'][*q - 'A'] = true;
        };
        mark(kConservative, sizeof(kConservative) / sizeof(*kConservative), cons);
        mark(kSemiConservative, sizeof(kSemiConservative) / sizeof(*kSemiConservative), semi);
        for (int a = 0; a < 26; a++)
            for (int b = 0; b < 26; b++)
                s[a][b] = a == b ? '*' : cons[a][b] ? ':' : semi[a][b] ? '.' : '_';
        for (int a = 0; a < kSymbols; a++) s[a][kGap] = s[kGap][a] = '_';
        s[kGap][kGap] = '*';
    }
};

const SignMatrix& signs()
{
    static const SignMatrix m;
    return m;
}

inline int class_of(char sign) { return sign == '*' ? 0 : sign == ':' ? 1 : sign == '.' ? 2 : 3; }

inline double signed_weight(int cls, const double* w) { return cls == 0 ? w[0] : -w[cls]; }

// First letter X (A..Z) with sign(c1, X) == want that is not conservative with the letter it
// replaces (the mutation rule, README "Mutation"); -1 if there is none.
int first_letter(int c1, char want, int replaced)
{
    const SignMatrix& m = signs();
    for (int x = 0; x < 26; x++)
        if (m.s[c1][x] == want && m.s[replaced][x] != ':') return x;
    return -1;
}

// Choice between two candidate (difference, letter) pairs: the first if it is at least as good
// and exists, else the second if it exists, else the first.
int prefer(bool is_max, double d1, int s1, double d2, int s2)
{
    bool first_ok = is_max ? d1 >= d2 : d1 <= d2;
    if (first_ok && s1 >= 0) return s1;
    return s2 >= 0 ? s2 : s1;
}

} // namespace

char sign_of(int c1, int c2) { return signs().s[c1][c2]; }

int symbol_index(char c)
{
    if (c >= 'A' && c <= 'Z') return c - 'A';
    return c == '-' ? kGap : -1;
}

char symbol_char(int idx) { return idx < 26 ? char('A' + idx) : '-'; }

int best_substitute(int c1, int c2, const double* w, bool is_max)
{
    char sg = sign_of(c1, c2);
    int to_colon = first_letter(c1, ':', c2);
    int to_dot = first_letter(c1, '.', c2);
    int to_space = first_letter(c1, '_', c2);
    if (is_max) {
        if (sg == '.' || sg == '_') return c1;       // an identical letter always wins
        double keep = sg == '*' ? -w[0] : w[1];      // minus the weight being given up
        return prefer(true, keep - w[2], to_dot, keep - w[3], to_space);
    }
    switch (sg) {
    case '*': return prefer(false, -w[0] - w[2], to_dot, -w[0] - w[3], to_space);
    case ':': return prefer(false, w[1] - w[2], to_dot, w[1] - w[3], to_space);
    case '.': {
        int s = prefer(false, w[2] - w[1], to_colon, w[2] - w[3], to_space);
        return s >= 0 ? s : c1;
    }
    default: {
        int s = prefer(false, w[3] - w[1], to_colon, w[3] - w[2], to_dot);
        return s >= 0 ? s : c1;
    }
    }
}

int build_tables(const double* w, int is_max, long long max_len2, psa_pair_table* pub, DeviceTable* dev)
{
    for (int k = 0; k < 4; k++)
        if (!std::isfinite(w[k])) return PSA_ERR_WEIGHTS;
    if (max_len2 < 1) max_len2 = 1;

    psa_pair_table t;
    std::memset(&t, 0, sizeof(t));
    std::vector<double> distinct;
    bool has_none = false;
    for (int c2 = 0; c2 < kSymbols; c2++)
        for (int c1 = 0; c1 < kSymbols; c1++) {
            char sg = sign_of(c1, c2);
            t.sign[c2][c1] = sg;
            int sub = best_substitute(c1, c2, w, is_max != 0);
            if (sub < 0) { has_none = true; continue; }
            t.substitute[c2][c1] = symbol_char(sub);
            // the difference that enters the score (cpu_funcs.c:285): w(sign(c1,sub)) - w(sign(c1,c2))
            double d = signed_weight(class_of(sign_of(c1, sub)), w) - signed_weight(class_of(sg), w);
            t.diff[c2][c1] = d;
            distinct.push_back(d);
        }
    std::sort(distinct.begin(), distinct.end());
    distinct.erase(std::unique(distinct.begin(), distinct.end()), distinct.end());
    if (!is_max) std::reverse(distinct.begin(), distinct.end());   // rank 1 = worst, nranks = best
    t.nranks = (int)distinct.size();
    if (t.nranks > kMaxRanks) return PSA_ERR_ARG;                  // cannot happen: <= 10 class pairs
    for (int c2 = 0; c2 < kSymbols; c2++)
        for (int c1 = 0; c1 < kSymbols; c1++) {
            if (!t.substitute[c2][c1]) continue;
            int r = int(std::find(distinct.begin(), distinct.end(), t.diff[c2][c1]) - distinct.begin());
            t.rank[c2][c1] = uint8_t(r + 1);
        }

    // ---- exactness analysis / fixed point scale --------------------------------------------
    double maxabs = 0;
    for (int k = 0; k < 4; k++) maxabs = std::max(maxabs, std::fabs(w[k]));
    const double terms = double(max_len2) + 2.0;   // len2 pair weights + the two weights inside the difference
    // magnitudes whose sums leave the double range (the reference just produces +-inf scores there) have no
    // fixed-point image: refuse them instead of deriving a scale from an infinite bound
    if (!std::isfinite(maxabs * terms) || maxabs * terms > std::ldexp(1.0, 1000)) return PSA_ERR_WEIGHTS;
    int frac = -1;
    for (int j = 0; j <= 60 && frac < 0; j++) {
        bool integral = true;
        for (int k = 0; k < 4; k++) {
            double s = std::ldexp(w[k], j);
            if (s != std::floor(s)) integral = false;
        }
        if (!integral) continue;
        // every partial sum is an integer multiple of 2^-j; exact in double iff below 2^53 units
        if (std::ldexp(maxabs, j) * terms <= 9007199254740992.0) frac = j;
        else break;                                  // larger j only makes it worse
    }
    int64_t slack = 0;
    if (frac >= 0) {
        t.exact = 1;
    } else {
        t.exact = 0;
        int e = 0;
        std::frexp(maxabs * terms, &e);              // maxabs*terms < 2^e
        frac = 60 - e;                               // |key| < 2^60
        // |key/2^frac - exact| <= terms * 2^-(frac+1)                    (rounded weights)
        // |reference double - exact| <= terms * 2^-53 * terms * maxabs * (1 + tiny)   (sequential adds)
        // two offsets can therefore swap order only if their keys differ by at most 2*(sum of both bounds).
        double fixed_err = terms * 0.5;
        double sum_err = std::ldexp(maxabs, frac - 53) * terms * terms * 1.0625;
        slack = (int64_t)std::ceil(2.0 * (fixed_err + sum_err)) + 4;
    }
    t.frac_bits = frac;
    t.key_slack = slack;
    if (pub) *pub = t;

    if (dev) {
        std::memset(dev, 0, sizeof(*dev));
        for (int c2 = 0; c2 < kSymbols; c2++)
            for (int c1 = 0; c1 < kSymbols; c1++) {
                dev->code[c2][c1] = uint8_t(class_of(t.sign[c2][c1]) | (t.rank[c2][c1] << 2));
                dev->sub[c2][c1] = (uint8_t)t.substitute[c2][c1];
            }
        const int64_t goal = is_max ? 1 : -1;
        int64_t fixed_w[4];
        for (int c = 0; c < 4; c++) {
            dev->wcls[c] = signed_weight(c, w);
            fixed_w[c] = (int64_t)std::llround(std::ldexp(dev->wcls[c], frac));
            dev->kcls[c] = goal * fixed_w[c];
        }
        dev->kdiff[0] = 0;
        dev->wdiff[0] = 0;
        for (int r = 0; r < t.nranks; r++) dev->wdiff[r + 1] = distinct[r];
        // fixed-point difference of a rank = fixed(after) - fixed(before) of any pair carrying it; in
        // exact mode all pairs of a rank agree, otherwise take the first (the slack covers the spread)
        std::vector<bool> seen(t.nranks + 1, false);
        for (int c2 = 0; c2 < kSymbols; c2++)
            for (int c1 = 0; c1 < kSymbols; c1++) {
                int r = t.rank[c2][c1];
                if (!r || seen[r]) continue;
                seen[r] = true;
                int before = class_of(t.sign[c2][c1]);
                int after = class_of(sign_of(c1, symbol_index(t.substitute[c2][c1])));
                dev->kdiff[r] = goal * (fixed_w[after] - fixed_w[before]);
            }
        dev->key_slack = slack;
        dev->nranks = t.nranks;
        dev->is_max = is_max ? 1 : 0;
        dev->exact = t.exact;
        dev->has_none = has_none ? 1 : 0;
        // Is "this pair's substitution has the best rank" decided by the pair's sign class alone?  (True for a
        // maximum: replacing by the Seq1 letter itself is best exactly on '.' or '_' pairs.)  Then the scan needs no
        // rank plane: it derives the bit from the two class planes it reads anyway.
        {
            int yes = 0, no = 0;
            for (int c2 = 0; c2 < kSymbols; c2++)
                for (int c1 = 0; c1 < kSymbols; c1++) {
                    const int cls = class_of(t.sign[c2][c1]);
                    (t.rank[c2][c1] == t.nranks ? yes : no) |= 1 << cls;
                }
            dev->top_rank_lut = (yes & no) == 0 ? yes : -1;
        }
        // Bit-plane columns for k_profile: per plane kind and Seq1 symbol, bit r = what row symbol r gets at a position
        // holding that Seq1 symbol (class bit 0, class bit 1, "rank == nranks - k").
        for (int c1 = 0; c1 < kRowPad; c1++)
            for (int kind = 0; kind < kPlaneKinds; kind++) dev->col[kind][c1] = 0;
        for (int c1 = 0; c1 < kSymbols; c1++)
            for (int r = 0; r < kSymbols; r++) {
                const uint32_t code = dev->code[r][c1];
                dev->col[0][c1] |= (code & 1u) << r;
                dev->col[1][c1] |= ((code >> 1) & 1u) << r;
                const int rank = int(code >> 2);
                for (int k = 0; k < kPlaneKinds - 2; k++)
                    dev->col[2 + k][c1] |= uint32_t(rank != 0 && rank == t.nranks - k) << r;
            }
    }
    return PSA_OK;
}

} // namespace psa

extern "C" int psa_build_pair_table(const double weights[4], int is_max, long long max_len2, psa_pair_table* out)
{
    if (!weights || !out) return PSA_ERR_ARG;
    return psa::build_tables(weights, is_max, max_len2, out, nullptr);
}
