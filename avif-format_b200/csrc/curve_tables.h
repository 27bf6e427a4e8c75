// curve_tables.h -- exact float -> code tables for the encode-side transfer curves.
//
// The headline conversion quantises code = trunc(clamp(OETF(x) * max)) with OETF = LinearToPQ (two powf) or
// LinearToSMPTE428 (one powf) -- ColorTransfer.cpp:69-92,119-127, WriteHeifImage.cpp:1093-1130.  Evaluating the
// glibc-exact powf for every sample costs ~110 issue slots per channel and caps the kernel at ~30 % of the HBM
// roofline.  But the OUTPUT is one of 2^depth codes, and as a function of the input float the code is a step
// function: outside narrow "fuzzy bands" around each step (where libm rounding noise makes the reference's own
// result non-monotone) the code is decided by comparing the input with a threshold.  So:
//
//   * a sweep kernel evaluates the EXACT curve (the same device functions the generic kernel uses) for every
//     non-negative finite float (2^31 - 2^23 inputs, a few milliseconds on B200) and records, per code c, the
//     smallest and largest input that produced it;
//   * the host turns that into thresholds first_k = min{x : code >= k}, band ends last_k = max{x : code < k} and
//     a two-level table: per binade (octave) a bucket size 2^S chosen so that no bucket meets two steps, per
//     bucket one 32-bit word {k-1, offset of the band start};
//   * the conversion kernel does two shared-memory look-ups and a handful of integer instructions per sample;
//     samples that fall inside a (conservatively widened) band -- about 2 % of typical data -- are handed to the
//     exact evaluation, compacted across the warp so the exact code runs at full lane occupancy;
//   * a verification kernel then re-sweeps every float and checks table == exact outside the bands; a table that
//     fails (it never has) is discarded and the generic kernel keeps serving that configuration.
//
// Nothing here approximates: every output is either decided by a threshold derived from the exact curve, or is
// the exact curve itself.
#ifndef AVIF_CURVE_TABLES_H
#define AVIF_CURVE_TABLES_H

#include <stdint.h>

#include <string>

#include <vector_types.h>

namespace avifgpu
{

enum CurveId
{
    kCurveLinearToPQ = 0,
    kCurveLinearToSMPTE428 = 1
};

constexpr uint32_t kBucketOffsetBits = 20;          // low bits of a bucket word: quantised band-start offset
constexpr uint32_t kBucketOffsetNone = 1u << 19;    // "no step in this bucket": every offset compares below it
constexpr uint32_t kOffsetResolutionBits = 19;      // offsets inside a bucket are kept to 19 bits

// Flat (single-level) variant, used when one bucket size separates the steps of every binade and the table
// still fits in shared memory (true for PQ): buckets of 2^flatShift floats covering bit patterns
// [flatLow << flatShift, (flatHigh + 1) << flatShift); inputs outside are clamped to the end buckets, which
// hold no step.  One 64-bit entry per bucket:
//     .x = bit pattern of the step's band start (first_k), 0 when the bucket meets no step
//     .y = kUpper << 20 | bandWidth   (bandWidth = number of in-band floats from first_k on, < 2^20)
// code = kUpper - (bits < .x); the sample is in band iff 0 <= bits - .x < bandWidth.
constexpr uint32_t kFlatMaxShift = 14;
constexpr uint32_t kFlatMaxBytes = 132 * 1024;

// Device-resident table (global memory; kernels stage it into shared memory).
struct CurveTableView
{
    const uint2* octaves;     // 256 entries: .x = first bucket index, .y = S | r << 8 | wq << 16
    const uint32_t* buckets;  // bucketCount words: (k-1) << 20 | offset
    int32_t bucketCount;
    const uint2* flat;        // flatCount entries, or nullptr when the flat variant does not apply
    int32_t flatCount;
    uint32_t flatShift;
    uint32_t flatLow;         // bucket number (bits >> flatShift) of flat[0]
    uint32_t flatHigh;        // bucket number of flat[flatCount - 1]
};

struct CurveTableStats
{
    double buildMilliseconds = 0.0;
    uint64_t sweptInputs = 0;
    uint64_t inBandInputs = 0;     // inputs the kernel sends to the exact path (two-level table)
    uint64_t flatInBandInputs = 0; // same for the flat variant (exact per-bucket band widths)
    int32_t flatBuckets = 0;       // 0 when the flat variant does not apply
    uint64_t verifyMismatches = 0; // must be 0
    int32_t steps = 0;             // thresholds found
    int32_t bands = 0;             // thresholds with a non-empty fuzzy band
    uint32_t widestBand = 0;       // in ulps
};

struct CurveTable
{
    int32_t curve = 0;
    int32_t param = 0; // PQ: peak nits
    int32_t depth = 0;
    bool valid = false;
    CurveTableView view{};
    CurveTableStats stats;
    std::string error;
    void* deviceOctaves = nullptr;
    void* deviceBuckets = nullptr;
    void* deviceFlat = nullptr;
};

// Builds (sweeps, assembles, uploads, verifies) the table on the current device.  Synchronous; uses `stream`.
// Returns true when the table is valid.  On failure table->error says why and the caller must not use it.
bool BuildCurveTable(int curve, int param, int depth, void* stream, CurveTable* table);
void FreeCurveTable(CurveTable* table);

} // namespace avifgpu

#endif
