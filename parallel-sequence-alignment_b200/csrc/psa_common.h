// psa_common.h -- structures shared by the host resolver and the sm_100a kernels.
#pragma once
#include <stdint.h>

namespace psa {

constexpr int kSymbols   = 27;   // 'A'..'Z', '-'
constexpr int kGap       = 26;
constexpr int kRowPad    = 32;   // table rows padded to 32 columns (one shared-memory bank line)
constexpr int kZeroRow   = 27;   // bit-plane row of zeros used to pad Seq2 to a multiple of 32 steps
constexpr int kPlaneRows = 28;
constexpr int kMaxRanks  = 15;   // rank field is 4 bits; 0 = "no substitute"
constexpr uint8_t kBadSymbol = 0xFF;
constexpr int kPlaneKinds = 6;   // bit-plane kinds k_profile can build: class bit 0, class bit 1, up to 4 rank planes

// Packed per-pair payload, code[c2][c1]:
//   bits 0-1  sign class   0 '*'  1 ':'  2 '.'  3 '_'
//   bits 2-5  rank of the best substitution difference (0 none, 1..nranks; higher = better for the goal)
// The reference evaluates get_hashtable_sign + get_substitute + get_weight for every pair of every
// offset (cpu_funcs.c:277-285, cuda_funcs.cu:183-193); all three are pure functions of (c1,c2) for a
// given (weights, goal), so they are resolved once on the host into this table.
struct DeviceTable {
    uint8_t code[kSymbols][kRowPad];
    uint8_t sub[kSymbols][kRowPad];   // ASCII replacement letter (0 = none)
    int64_t kcls[4];                  // goal-signed fixed-point pair weight per class (key units)
    int64_t kdiff[kMaxRanks + 1];     // goal-signed fixed-point substitution difference per rank
    double  wcls[4];                  // +W1, -W2, -W3, -W4 exactly as get_weight returns them
    double  wdiff[kMaxRanks + 1];     // the reference's double difference per rank
    int64_t key_slack;                // 0 in exact mode
    int32_t nranks;
    int32_t is_max;
    int32_t exact;
    int32_t has_none;                 // some pair has no substitute (never observed; handled anyway)
    uint32_t col[kPlaneKinds][kRowPad];   // k_profile's columns: [plane kind][Seq1 symbol], bit r = row symbol r
    int32_t top_rank_lut;             // >= 0: "pair carries the best rank" is a function of its sign class alone (bit c = class c);
                                      // -1: it is not (then the scan reads it from a rank bit plane)
};

// The constants of the bit-sliced key build (SlicedKeys, psa_scan_core.cuh) for one (table, query length, plane count),
// resolved on the host once per launch so that the kernel does no 64-bit arithmetic on table fields per pass:
//   key' = c0 + ka N(b0) + kb N(b1) + kc N(b0&b1) + dv[first tracked plane met, else the floor]     (all >= 0, < 2^planes)
struct SlicedPlan {
    int32_t  ka, kb, kc;       // multipliers of the three vertical counters, |k| < 32
    uint32_t c0;               // bias + len2 * k('*') + the smallest rank term
    uint32_t dv[4];            // rank term of tracked plane k, above the smallest one (< 256)
    uint32_t dv_floor;         // rank term of an offset that met no tracked plane
    int32_t  floor_none;       // nothing below the tracked planes but "no substitute at all"
    int32_t  floor_exact;      // an offset that met no tracked plane still has an exactly known rank
    int64_t  bias;             // key = key' - bias
    int64_t  kfl;              // key units of the floor rank's difference
};

constexpr int64_t kKeyNone = INT64_MIN;       // key of an offset with no possible mutation / no data

// One record per (query, tile of offsets), written by the scan / exact kernels.
struct TileRec {
    int64_t key;        // best key in the tile (kKeyNone if none); exact integer order in exact mode
    int64_t ub_key;     // scan engine, re-score mode: upper estimate of every key in the tile
    double  score;      // unused (kept for layout)
    int32_t offset;     // absolute offset of `key`
    int32_t ub_offset;
    int32_t flags;      // bit0: written by the exact kernel (key is a sortable reference double when !exact)
    int32_t pad;
};
constexpr int kTileExact = 1;

// One record per query, written by the finish kernel.  Same layout as the public psa_result (psa_b200.h;
// static_asserts in psa_engine.cu), so the device-to-host copy lands in the caller's array unchanged.
struct QueryRec {
    int32_t offset;       // -1: no mutation possible at any offset
    int32_t char_offset;
    int32_t ch;           // replacement letter in the low byte (psa_mutant::ch + its padding)
    int32_t rank;
    double  score;        // the reference's double: exact mode from the integer counts, else the re-scored sum
    int64_t counts[4];
};

// Launch geometry shared by all kernels of one batch.
struct BatchGeom {
    int64_t len1;
    int64_t first;          // single-query range mode: absolute first offset (else 0)
    int64_t last;           // single-query range mode: absolute last offset (exclusive), else -1 = all
    int32_t nq;
    int32_t tile;           // offsets per tile
    int32_t total_tiles;
    int32_t tiles_per_query;   // > 0 when every query has the same number of tiles (no search needed), else 0
    int32_t uniform_len2;      // > 0 when every query has this length: qoff / tile_start are implicit (not uploaded)
};

} // namespace psa
