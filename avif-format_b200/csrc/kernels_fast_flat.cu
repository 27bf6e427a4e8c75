// kernels_fast_flat.cu -- the float RGB -> planar YCbCr encode kernel for curves whose step table has the flat
// form (curve_tables.h): BASELINE config 2, 7680x4320 RGB32f -> 12-bit PQ 4:2:0.
//
//   * one CTA per SM, kFlatWarps warps; the step table (compact one-word entries + first_k, 67 KB at 12 bits; or the
//     128 KB 64-bit flat table / the two-level table where the compact form does not apply) sits in shared memory next to
//     one 3 KB staging buffer per warp.  The compact table's image is brought in by the copy engine (table_staging.cuh).
//     Measured on the 8K PQ frame (Gpx/s, round 1): 16 warps 238 | 20: 251 | 24: 273 | 28: 278 | 30: 267 (30 leaves 64
//     registers); the kernel is bound by instruction issue and shared-memory wavefronts, so resident warps matter, and
//     the copy engine keeps the per-warp register cost of a fetch at zero;
//   * a warp converts tiles of 2 rows x 128 pixels, walking straight down one tile column (FlatSchedule).  The two
//     1536-byte row segments of a tile are fetched by the copy engine (cp.async.bulk, completion on the warp's own
//     mbarrier) straight into the staging buffer, so the fetch of tile i+1 costs the warp two instructions and no
//     registers and overlaps the matrix and the stores of tile i; the lanes then read their 48 bytes per row with three
//     conflict-free LDS.128;
//   * float -> code: one table look-up per sample (LookupCurveCompact: a 32-bit gather, 12 instructions); samples the
//     look-up flags as possibly inside a fuzzy band (~1.5 %) are resolved by first_k and one bit of the L2-resident band
//     bitmap, up to two samples of a lane at a time; +inf / NaN take the exact glibc-identical evaluation;
//   * forward matrix, quantisation, chroma down-filter and stores: StoreTile (kernels_fast_common.cuh, packed FP32).
#include "kernels_fast_common.cuh"
#include "table_staging.cuh"
#include "../../include/avifgpu.h"

namespace avifgpu
{

using namespace avifpix;
using namespace fastenc;
using namespace staging;
using avifmath::LibmTables;

namespace
{

#ifndef AVIF_FLAT_WARPS
#define AVIF_FLAT_WARPS 28
#endif
constexpr int kFlatWarps = AVIF_FLAT_WARPS;
constexpr int kFlatThreads = kFlatWarps * 32;
constexpr int kRowSegmentBytes = kTilePixels * 12;      // one tile row of RGB32f
constexpr int kStageBytesPerWarp = 2 * kRowSegmentBytes; // both rows, linear
constexpr int kRowSegmentWords = kRowSegmentBytes / 4;
constexpr int kSharedBarriers = 256;                     // kFlatWarps x 8 bytes, padded; the last slot is the table's barrier
constexpr int kTableBarrierSlot = kSharedBarriers / 8 - 1;
static_assert(kFlatWarps <= kTableBarrierSlot, "the warps' barriers and the table's share kSharedBarriers");
constexpr int kSharedLimit = 227 * 1024;
__host__ __device__ constexpr int FlatFixedBytes() { return kSharedLibm + kSharedBarriers + kFlatWarps * kStageBytesPerWarp; }

// How the persistent warps share the tiles, worked out once on the host (the grid is known at launch).  The image is cut
// into `items` = tilesX columns x `segments` runs of consecutive tile rows (run lengths differ by at most one), one item
// per warp when the grid has at least tilesX warps -- so a warp walks straight down one tile column and its per-tile
// bookkeeping is four pointer increments by launch constants: no column wrap, no division, no per-tile schedule state.
struct FlatSchedule
{
    int32_t tilesX, tileRows, warpCount;
    int32_t segments, items;          // items = tilesX * segments
    int32_t segmentRows, longSegments; // segment s holds segmentRows (+ 1 if s < longSegments) tile rows
    int32_t lastColumnBytes;          // row-segment bytes of the last tile column (width need not be a multiple of 128)
    int32_t unpairedTileRow;          // tile row whose second image row does not exist (odd row count), or -1
    // Run lengths differ by one tile row.  With the warps of a CTA on adjacent items a CTA is all long runs or all short
    // ones, and the launch ends when the long CTAs do (config 2: 45 of 148 CTAs run 32 tile rounds, the rest 31, a 3 us
    // tail).  scatter = 1 deals the items round-robin over the CTAs instead (item = warp * CTAs + CTA): every CTA gets
    // the same mix, so the extra round is run by 8-9 warps per SM rather than by 28 on a third of the SMs.
    int32_t scatter;
};

// One lane of the (converged) warp.
__device__ __forceinline__ bool ElectOne()
{
    uint32_t elected;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(elected));
    return elected != 0;
}

// TWO_LEVEL = 0: the flat table (64-bit entries) + band bitmap.  TWO_LEVEL = 1: the per-binade two-level table (curves whose
// steps are too dense for one bucket size, e.g. 12-bit SMPTE 428); its rare in-band samples take the exact evaluation in
// place.  TWO_LEVEL = 2: the flat table in its compact one-word form (curve_tables.h "Compact entries") + the first_k
// array + band bitmap: a 32-bit gather costs ~3.5 shared-memory wavefronts where the 64-bit one costs ~5.2, and that pipe
// is what bounds the 64-bit variant.  TWO_LEVEL = 3: the same with flatShift = 14 known at compile time (12-bit PQ).
// INTERLEAVED = 1: the reference's own output layout (heif_channel_interleaved RGB, WriteHeifImage.cpp:1098-1130) -- the
// codes are stored as they are, 3 x uint16 per pixel into plane Y's buffer, no matrix (XS = YS = 0 then).
template <int CURVE, int XS, int YS, int TWO_LEVEL, int INTERLEAVED>
__global__ void __launch_bounds__(kFlatThreads, 1) EncodeRgbF32FlatKernel(const FastEncodeParams p, const FlatSchedule schedule)
{
    constexpr bool kCompact = TWO_LEVEL >= 2;
    constexpr int kCompactShift = TWO_LEVEL == 3 ? 14 : 0;
    extern __shared__ __align__(128) uint8_t sharedBytes[];
    uint64_t* libmStorage = reinterpret_cast<uint64_t*>(sharedBytes);
    uint64_t* barriers = reinterpret_cast<uint64_t*>(sharedBytes + kSharedLibm);
    uint8_t* stageAll = sharedBytes + kSharedLibm + kSharedBarriers;
    uint2* flatEntries = reinterpret_cast<uint2*>(sharedBytes + FlatFixedBytes());
    uint32_t* compactEntries = reinterpret_cast<uint32_t*>(sharedBytes + FlatFixedBytes());                                  // TWO_LEVEL == 2 ...
    const uint32_t* firstBits = compactEntries + ((p.table.flatCount + 3) & ~3);                                             // ... then first_k per code
    uint2* octaves = reinterpret_cast<uint2*>(sharedBytes + FlatFixedBytes());            // TWO_LEVEL: 256 entries ...
    uint32_t* bucketWords = reinterpret_cast<uint32_t*>(sharedBytes + FlatFixedBytes() + 2048); // ... then the bucket words

    const int lane = threadIdx.x & 31;
    // Read through a shuffle so the compiler knows the warp index (and everything derived from it: tile coordinates,
    // copy addresses) is warp-uniform and keeps it in the uniform datapath, which the bulk-copy instruction needs.
    const int warpInBlock = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const uint32_t barrier = SharedAddress(barriers + warpInBlock);
    uint32_t* stage = reinterpret_cast<uint32_t*>(stageAll + warpInBlock * kStageBytesPerWarp);
    const uint32_t stageAddress = SharedAddress(stage);
    const uint32_t* myStage = stage + lane * 12; // row 0; row 1 is kRowSegmentWords further

    const int tilesX = schedule.tilesX;
    const int warpCount = schedule.warpCount;
    const int firstItem = schedule.scatter ? warpInBlock * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x) * kFlatWarps + warpInBlock;
    constexpr int kChromaRowsPerTile = YS ? 1 : 2;
    constexpr int kChromaTileBytes = XS ? kTilePixels : 2 * kTilePixels;

    // An item's tile column and its run of tile rows [rowBegin, rowEnd).
    auto itemRows = [&](int item, int& column, int& rowBegin, int& rowEnd)
    {
        const int segment = item / tilesX;
        column = item - segment * tilesX;
        rowBegin = segment * schedule.segmentRows + min(segment, schedule.longSegments);
        rowEnd = rowBegin + schedule.segmentRows + (segment < schedule.longSegments ? 1 : 0);
    };
    // One elected lane asks the copy engine for a tile's row segments (warp-uniform arguments).
    auto fetchTile = [&](int row, uint32_t bytes, int64_t offset)
    {
        const bool second = row != schedule.unpairedTileRow;
        const uint8_t* source = p.rows + offset;
        BarrierExpect(barrier, second ? 2u * bytes : bytes);
        BulkCopyToShared(stageAddress, source, bytes, barrier);
        if (second)
        {
            BulkCopyToShared(stageAddress + kRowSegmentBytes, source + p.rowStride, bytes, barrier);
        }
    };
    auto columnBytes = [&](int column) { return column == tilesX - 1 ? static_cast<uint32_t>(schedule.lastColumnBytes) : static_cast<uint32_t>(kRowSegmentBytes); };
    auto sourceOffsetOf = [&](int row, int column) { return static_cast<int64_t>(row) * 2 * p.rowStride + static_cast<int64_t>(column) * kRowSegmentBytes; };

    if (kCompact && threadIdx.x == 0)
    {
        BeginTableImageCopy(p.table, compactEntries, barriers + kTableBarrierSlot); // table_staging.cuh
    }
    if (ElectOne())
    {
        BarrierInit(barrier, 1);
        BarrierInitFence();
        if (firstItem < schedule.items)
        {
            int column, rowBegin, rowEnd;
            itemRows(firstItem, column, rowBegin, rowEnd);
            if (rowBegin < rowEnd)
            {
                fetchTile(rowBegin, columnBytes(column), sourceOffsetOf(rowBegin, column)); // in flight while the table is staged
            }
        }
    }

    const LibmTables t = avifmath::StageLibmTables(libmStorage, threadIdx.x, blockDim.x);
    if (kCompact)
    {
        // in flight (above)
    }
    else if (TWO_LEVEL)
    {
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
        {
            octaves[i] = p.table.octaves[i];
        }
        for (int i = threadIdx.x; i < p.table.bucketCount; i += blockDim.x)
        {
            bucketWords[i] = p.table.buckets[i];
        }
    }
    else
    {
        // 128 KB from L2: 128-bit copies, eight in flight per thread (the table is 16-byte aligned, count is even-padded)
        const uint4* source = reinterpret_cast<const uint4*>(p.table.flat);
        uint4* target = reinterpret_cast<uint4*>(flatEntries);
        const int pairs = (p.table.flatCount + 1) / 2;
#pragma unroll 8
        for (int i = threadIdx.x; i < pairs; i += blockDim.x)
        {
            target[i] = __ldg(source + i);
        }
    }
    __syncthreads(); // the libm tables, the barriers' initialisation
    if (kCompact)
    {
        WaitTableImage(barriers + kTableBarrierSlot);
    }

    const uint32_t flatShift = p.table.flatShift;
    const int32_t negativeLow = -static_cast<int32_t>(p.table.flatLow);
    const int32_t span = static_cast<int32_t>(p.table.flatHigh - p.table.flatLow);
    const uint32_t bandStrideLog2 = p.table.bandStrideLog2;
    const uint32_t* __restrict__ bandBits = p.table.bandBits;
    const uint32_t compactTopShift = 32u - flatShift;
    const uint32_t compactCodeMask = p.table.compactCodeMask;
    const uint32_t compactMagic = p.table.compactMagic;
    uint32_t parity = 0;

#pragma unroll 1
    for (int item = firstItem; item < schedule.items; item += warpCount)
    {
        int column, rowBegin, rowEnd;
        itemRows(item, column, rowBegin, rowEnd);
        if (item != firstItem && rowBegin < rowEnd && ElectOne())
        {
            fetchTile(rowBegin, columnBytes(column), sourceOffsetOf(rowBegin, column)); // only grids smaller than tilesX warps get here
        }
        const uint32_t tileBytes = columnBytes(column);
        const bool laneActive = column * kTilePixels + lane * 4 < p.width;
        int64_t sourceOffset = sourceOffsetOf(rowBegin, column);
        uint8_t* yPointer = p.planeY + static_cast<int64_t>(rowBegin) * 2 * p.strideY +
                            (INTERLEAVED ? static_cast<int64_t>(column) * (6 * kTilePixels) + lane * 24 : static_cast<int64_t>(column) * (2 * kTilePixels) + lane * 8);
        uint8_t* cbPointer = p.planeCb + static_cast<int64_t>(rowBegin) * kChromaRowsPerTile * p.strideCb + static_cast<int64_t>(column) * kChromaTileBytes + lane * (XS ? 4 : 8);
        uint8_t* crPointer = p.planeCr + static_cast<int64_t>(rowBegin) * kChromaRowsPerTile * p.strideCr + static_cast<int64_t>(column) * kChromaTileBytes + lane * (XS ? 4 : 8);

#pragma unroll 1
    for (int tileRow = rowBegin; tileRow < rowEnd; ++tileRow)
    {
        const bool secondRow = tileRow != schedule.unpairedTileRow;
        sourceOffset += 2 * p.rowStride; // the next tile of this column

        BarrierWait(barrier, parity);
        parity ^= 1u;

        // ---- float -> code through the exact step table ------------------------------------------------------------
        float codeF[kValuesPerLane]; // the codes, as the floats the forward matrix consumes
        uint32_t bandMask = 0;
        int32_t largest = 0; // max over the samples as signed integers: > 0x7f7fffff <=> a +inf / NaN is among them
#pragma unroll
        for (int q = 0; q < 6; ++q)
        {
            const uint4 w = *reinterpret_cast<const uint4*>(myStage + (q / 3) * kRowSegmentWords + (q % 3) * 4);
            const uint32_t bits[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
            for (int e = 0; e < 4; ++e)
            {
                const int j = 4 * q + e;
                bool inBand;
                if (kCompact)
                {
                    uint32_t entry;
                    codeF[j] = LookupCurveCompact<kCompactShift>(bits[e], compactEntries, flatShift, negativeLow, span, compactTopShift, compactCodeMask,
                                                                 compactMagic, inBand, entry);
                }
                else if (TWO_LEVEL)
                {
                    codeF[j] = CodeToFloat(LookupCurveCode(bits[e], octaves, bucketWords, inBand)); // reports +inf / NaN in band itself
                }
                else
                {
                    codeF[j] = LookupCurveFlat(bits[e], flatEntries, flatShift, negativeLow, span, inBand);
                }
                asm("{ .reg .pred p; setp.ne.u32 p, %1, 0; @p or.b32 %0, %0, %2; }" : "+r"(bandMask) : "r"(static_cast<uint32_t>(inBand)), "r"(1u << j));
            }
            if (TWO_LEVEL != 1)
            {
                largest = max(largest, __vimax3_s32(static_cast<int32_t>(w.x), static_cast<int32_t>(w.y), static_cast<int32_t>(w.z)));
                largest = max(largest, static_cast<int32_t>(w.w));
            }
        }

        // ---- in-band samples: one bit of the band bitmap each, two loads in flight per lane ---------------------------
        auto stagedBits = [&](int j) { return myStage[j + (j >= 12 ? kRowSegmentWords - 12 : 0)]; };
        uint32_t lowerMask = 0; // samples whose exact code is one below the table's
        if (kCompact)
        {
            // flagged samples: which step, how far above its first_k, one bit of the band bitmap -- two at a time
            uint32_t pending = bandMask;
            while (pending != 0)
            {
                uint32_t word[2] = { 0xffffffffu, 0xffffffffu }, index[2] = { 0u, 0u }, sampleBit[2] = { 0u, 0u };
#pragma unroll
                for (int u = 0; u < 2; ++u)
                {
                    if (pending != 0)
                    {
                        const int j = __ffs(static_cast<int>(pending)) - 1;
                        pending &= pending - 1;
                        const uint32_t bits = stagedBits(j);
                        const int32_t bucket = __viaddmin_s32_relu(static_cast<int32_t>(bits) >> flatShift, negativeLow, span);
                        const uint32_t entry = compactEntries[bucket];
                        const uint32_t k = ((entry & compactCodeMask) >> kCompactLenBits) + ((entry >> compactTopShift) != 0 ? 1u : 0u);
                        const uint32_t distance = bits - firstBits[k];
                        if (k != 0 && distance < (1u << bandStrideLog2)) // else flagged by the superset test only: the table's code stands
                        {
                            index[u] = (k << bandStrideLog2) + distance;
                            word[u] = __ldg(bandBits + (index[u] >> 5));
                            sampleBit[u] = 1u << j;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
                {
                    if (((word[u] >> (index[u] & 31u)) & 1u) == 0)
                    {
                        lowerMask |= sampleBit[u];
                    }
                }
            }
        }
        else if (!TWO_LEVEL)
        {
            uint32_t pending = bandMask;
            while (pending != 0)
            {
                const int j0 = __ffs(static_cast<int>(pending)) - 1;
                pending &= pending - 1;
                const uint32_t bits0 = stagedBits(j0);
                bool inBand;
                uint2 entry0;
                LookupCurveFlat(bits0, flatEntries, flatShift, negativeLow, span, inBand, entry0);
                const uint32_t index0 = BandBitIndex(bits0, entry0, bandStrideLog2);
                const uint32_t word0 = __ldg(bandBits + (index0 >> 5));
                uint32_t word1 = 0xffffffffu, index1 = 0, sample1 = 0;
                if (pending != 0)
                {
                    const int j1 = __ffs(static_cast<int>(pending)) - 1;
                    pending &= pending - 1;
                    const uint32_t bits1 = stagedBits(j1);
                    uint2 entry1;
                    LookupCurveFlat(bits1, flatEntries, flatShift, negativeLow, span, inBand, entry1);
                    index1 = BandBitIndex(bits1, entry1, bandStrideLog2);
                    word1 = __ldg(bandBits + (index1 >> 5));
                    sample1 = 1u << j1;
                }
                lowerMask |= (((word0 >> (index0 & 31u)) & 1u) ^ 1u) << j0;
                if (((word1 >> (index1 & 31u)) & 1u) == 0)
                {
                    lowerMask |= sample1;
                }
            }
        }

        // ---- +inf / NaN (never in real frames): the exact evaluation, lane by lane ------------------------------------
        if (__any_sync(0xffffffffu, TWO_LEVEL == 1 ? bandMask != 0 : largest > 0x7f7fffff))
        {
            for (int j = 0; j < kValuesPerLane; ++j)
            {
                const uint32_t bits = stagedBits(j);
                if (TWO_LEVEL == 1 ? ((bandMask >> j) & 1u) != 0 : static_cast<int32_t>(bits) > 0x7f7fffff)
                {
                    const float exact = CodeToFloat(ExactCurveCode<CURVE>(__uint_as_float(bits), p.pqMultiplier, p.maxCodeFloat, t));
#pragma unroll
                    for (int slot = 0; slot < kValuesPerLane; ++slot)
                    {
                        if (slot == j)
                        {
                            codeF[slot] = exact;
                        }
                    }
                }
            }
        }

        // The staging buffer is free: fetch the next tile while this one goes through the matrix and the stores.
        __syncwarp();
        if (tileRow + 1 < rowEnd && ElectOne())
        {
            fetchTile(tileRow + 1, tileBytes, sourceOffset);
        }

#pragma unroll
        for (int j = 0; j < kValuesPerLane; ++j)
        {
            if (lowerMask & (1u << j))
            {
                codeF[j] -= 1.0f;
            }
        }

        if (laneActive)
        {
            if (INTERLEAVED)
            {
                // 4 pixels x 3 codes per row = 24 bytes: three 64-bit stores
#pragma unroll
                for (int r = 0; r < 2; ++r)
                {
                    if (r == 1 && !secondRow) break;
                    uint32_t words[6];
#pragma unroll
                    for (int w = 0; w < 6; ++w)
                    {
                        words[w] = __float2uint_rz(codeF[12 * r + 2 * w]) | (__float2uint_rz(codeF[12 * r + 2 * w + 1]) << 16);
                    }
                    uint2* target = reinterpret_cast<uint2*>(yPointer + r * p.strideY);
                    __stcs(target, make_uint2(words[0], words[1]));
                    __stcs(target + 1, make_uint2(words[2], words[3]));
                    __stcs(target + 2, make_uint2(words[4], words[5]));
                }
            }
            else
            {
                StoreTile<XS, YS>(p, codeF, yPointer, cbPointer, crPointer, secondRow);
            }
        }
        yPointer += 2 * p.strideY;
        cbPointer += kChromaRowsPerTile * p.strideCb;
        crPointer += kChromaRowsPerTile * p.strideCr;
    }
    }
}

// tableKind: 0 flat (64-bit entries), 1 two-level, 2 compact (32-bit entries + first_k per code)
inline size_t TableSharedBytes(const FastEncodeParams& fp, int tableKind)
{
    if (tableKind == 1) return 2048 + static_cast<size_t>(fp.table.bucketCount) * sizeof(uint32_t);
    if (tableKind >= 2) return fp.table.compactImageBytes;
    return static_cast<size_t>((fp.table.flatCount + 1) / 2) * sizeof(uint4);
}

template <int CURVE, int XS, int YS, int TWO_LEVEL, int INTERLEAVED = 0>
cudaError_t LaunchFlatKernel(const FastEncodeParams& fp, int smCount, cudaStream_t stream)
{
    const size_t shared = static_cast<size_t>(FlatFixedBytes()) + TableSharedBytes(fp, TWO_LEVEL);
    static std::atomic<uint64_t> configuredDevices{ 0 }; // per instantiation
    {
        const cudaError_t e = AllowDynamicShared(EncodeRgbF32FlatKernel<CURVE, XS, YS, TWO_LEVEL, INTERLEAVED>, kSharedLimit, configuredDevices);
        if (e != cudaSuccess)
        {
            return e;
        }
    }
    const long long tiles = static_cast<long long>((fp.width + kTilePixels - 1) / kTilePixels) * ((fp.rowCount + 1) / 2);
    if (tiles > 0x7fffffffll || shared > static_cast<size_t>(kSharedLimit))
    {
        return cudaErrorInvalidValue;
    }
    long long blocks = (tiles + kFlatWarps - 1) / kFlatWarps;
    if (blocks > smCount)
    {
        blocks = smCount;
    }
    FlatSchedule schedule{};
    schedule.tilesX = (fp.width + kTilePixels - 1) / kTilePixels;
    schedule.tileRows = (fp.rowCount + 1) / 2;
    schedule.warpCount = static_cast<int32_t>(blocks) * kFlatWarps;
    schedule.segments = schedule.warpCount / schedule.tilesX;
    if (schedule.segments < 1) schedule.segments = 1;
    if (schedule.segments > schedule.tileRows) schedule.segments = schedule.tileRows;
    schedule.items = schedule.tilesX * schedule.segments;
    schedule.segmentRows = schedule.tileRows / schedule.segments;
    schedule.longSegments = schedule.tileRows % schedule.segments;
    schedule.lastColumnBytes = (fp.width - (schedule.tilesX - 1) * kTilePixels) * 12;
    schedule.unpairedTileRow = (fp.rowCount & 1) ? fp.rowCount / 2 : -1;
    static const bool scatterOff = []() { const char* v = getenv("AVIFGPU_FLAT_SCATTER"); return v != nullptr && v[0] == '0'; }(); // A/B measurements
    schedule.scatter = scatterOff ? 0 : 1;
    EncodeRgbF32FlatKernel<CURVE, XS, YS, TWO_LEVEL, INTERLEAVED><<<static_cast<unsigned>(blocks), kFlatThreads, shared, stream>>>(fp, schedule);
    return cudaGetLastError();
}

template <int CURVE, int TWO_LEVEL>
cudaError_t DispatchFlatChroma(const FastEncodeParams& fp, int xs, int ys, int smCount, cudaStream_t stream)
{
    if (xs == 1 && ys == 1) return LaunchFlatKernel<CURVE, 1, 1, TWO_LEVEL>(fp, smCount, stream);
    if (xs == 1) return LaunchFlatKernel<CURVE, 1, 0, TWO_LEVEL>(fp, smCount, stream);
    return LaunchFlatKernel<CURVE, 0, 0, TWO_LEVEL>(fp, smCount, stream);
}

} // namespace

static bool FlatTableFits(const FastEncodeParams& fp)
{
    return fp.table.flat != nullptr && fp.table.bandBits != nullptr &&
           static_cast<size_t>(FlatFixedBytes()) + TableSharedBytes(fp, 0) <= static_cast<size_t>(kSharedLimit);
}

static bool CompactTableFits(const FastEncodeParams& fp)
{
    return fp.table.compact != nullptr && fp.table.firstBits != nullptr && fp.table.bandBits != nullptr &&
           static_cast<size_t>(FlatFixedBytes()) + TableSharedBytes(fp, 2) <= static_cast<size_t>(kSharedLimit);
}

static bool TwoLevelTableFits(const FastEncodeParams& fp)
{
    return fp.table.buckets != nullptr && fp.table.octaves != nullptr &&
           static_cast<size_t>(FlatFixedBytes()) + TableSharedBytes(fp, 1) <= static_cast<size_t>(kSharedLimit);
}

// True when the copy-engine kernel can serve this table: the flat form with its bitmap, else the two-level form, in
// shared memory next to the staging buffers.
bool FlatEncodeApplies(const FastEncodeParams& fp)
{
    return CompactTableFits(fp) || FlatTableFits(fp) || TwoLevelTableFits(fp);
}

// The reference's interleaved RGB layout through the same kernel (fp.planeY / strideY = the interleaved buffer).
cudaError_t LaunchFastEncodeFlatInterleaved(const FastEncodeParams& fp, int curve, int smCount, cudaStream_t stream)
{
    if (CompactTableFits(fp))
    {
        if (curve == kCurveLinearToPQ)
        {
            return fp.table.flatShift == 14 ? LaunchFlatKernel<kCurveLinearToPQ, 0, 0, 3, 1>(fp, smCount, stream) : LaunchFlatKernel<kCurveLinearToPQ, 0, 0, 2, 1>(fp, smCount, stream);
        }
        return LaunchFlatKernel<kCurveLinearToSMPTE428, 0, 0, 2, 1>(fp, smCount, stream);
    }
    if (FlatTableFits(fp))
    {
        if (curve == kCurveLinearToPQ) return LaunchFlatKernel<kCurveLinearToPQ, 0, 0, 0, 1>(fp, smCount, stream);
        return LaunchFlatKernel<kCurveLinearToSMPTE428, 0, 0, 0, 1>(fp, smCount, stream);
    }
    if (curve == kCurveLinearToPQ) return LaunchFlatKernel<kCurveLinearToPQ, 0, 0, 1, 1>(fp, smCount, stream);
    return LaunchFlatKernel<kCurveLinearToSMPTE428, 0, 0, 1, 1>(fp, smCount, stream);
}

cudaError_t LaunchFastEncodeFlat(const FastEncodeParams& fp, int curve, int xs, int ys, int smCount, cudaStream_t stream)
{
    if (CompactTableFits(fp) && !fp.preferWideEntries)
    {
        if (curve == kCurveLinearToPQ)
        {
            return fp.table.flatShift == 14 ? DispatchFlatChroma<kCurveLinearToPQ, 3>(fp, xs, ys, smCount, stream) : DispatchFlatChroma<kCurveLinearToPQ, 2>(fp, xs, ys, smCount, stream);
        }
        return DispatchFlatChroma<kCurveLinearToSMPTE428, 2>(fp, xs, ys, smCount, stream);
    }
    if (FlatTableFits(fp))
    {
        if (curve == kCurveLinearToPQ) return DispatchFlatChroma<kCurveLinearToPQ, 0>(fp, xs, ys, smCount, stream);
        return DispatchFlatChroma<kCurveLinearToSMPTE428, 0>(fp, xs, ys, smCount, stream);
    }
    if (curve == kCurveLinearToPQ) return DispatchFlatChroma<kCurveLinearToPQ, 1>(fp, xs, ys, smCount, stream);
    return DispatchFlatChroma<kCurveLinearToSMPTE428, 1>(fp, xs, ys, smCount, stream);
}

} // namespace avifgpu
