// kernels_fast_rgba.cu -- float RGBA hosts -> planar YCbCr + alpha plane through the step table (compact one-word entries
// where the table has them, else the 64-bit flat entries; curve_tables.h) and the band bitmap
// (CreateHeifImageRGBThirtyTwoBit's alpha branch, WriteHeifImage.cpp:1039-1077, fused with the libheif stage).
//
// Same warp tile as kernels_fast_flat.cu (2 rows x 128 pixels, a lane owns 4 adjacent pixels in both rows), but 16 bytes
// per pixel do not lay out conflict-free in a linear staging buffer, so the loads are per-lane 128-bit loads (one pixel
// each, issued one tile ahead) and only the 24 colour samples of a lane -- after the clamp / premultiplication, the
// values the curve actually sees -- are parked in shared memory for the band-bitmap probes.  One CTA of 16 warps per SM.
#include "kernels_fast_common.cuh"
#include "table_staging.cuh"
#include "../../include/avifgpu.h"

namespace avifgpu
{

using namespace avifpix;
using namespace fastenc;
using avifmath::LibmTables;

namespace
{

#ifndef AVIF_RGBA_WARPS
#define AVIF_RGBA_WARPS 16
#endif
constexpr int kRgbaWarps = AVIF_RGBA_WARPS;
constexpr int kRgbaThreads = kRgbaWarps * 32;
constexpr int kLaneStrideWords = 28; // 24 colour samples + padding: 16-byte aligned, conflict-free for STS.128
constexpr int kStagePerWarp = 32 * kLaneStrideWords * 4;
constexpr int kSharedLimit = 227 * 1024;
constexpr int kTableBarrierBytes = 16; // the table image's mbarrier, padded
__host__ __device__ constexpr int RgbaFixedBytes() { return kSharedLibm + kTableBarrierBytes + kRgbaWarps * kStagePerWarp; }

// COMPACT = 1: compact table + first_k array in shared memory (kernels_fast_flat.cu has the commentary).
template <int CURVE, int XS, int YS, int COMPACT>
__global__ void __launch_bounds__(kRgbaThreads, 1) EncodeRgbaF32FlatKernel(const FastEncodeParams p)
{
    extern __shared__ __align__(16) uint8_t sharedBytes[];
    uint64_t* libmStorage = reinterpret_cast<uint64_t*>(sharedBytes);
    uint64_t* tableBarrier = reinterpret_cast<uint64_t*>(sharedBytes + kSharedLibm);
    uint32_t* stageAll = reinterpret_cast<uint32_t*>(sharedBytes + kSharedLibm + kTableBarrierBytes);
    uint2* flatEntries = reinterpret_cast<uint2*>(sharedBytes + RgbaFixedBytes());
    uint32_t* compactEntries = reinterpret_cast<uint32_t*>(sharedBytes + RgbaFixedBytes());
    const uint32_t* firstBits = compactEntries + ((p.table.flatCount + 3) & ~3);

    if (COMPACT && threadIdx.x == 0)
    {
        staging::BeginTableImageCopy(p.table, compactEntries, tableBarrier); // table_staging.cuh
    }
    const LibmTables t = avifmath::StageLibmTables(libmStorage, threadIdx.x, blockDim.x);
    if (!COMPACT)
    {
        const uint4* source = reinterpret_cast<const uint4*>(p.table.flat);
        uint4* target = reinterpret_cast<uint4*>(flatEntries);
        const int pairs = (p.table.flatCount + 1) / 2;
#pragma unroll 8
        for (int i = threadIdx.x; i < pairs; i += blockDim.x)
        {
            target[i] = __ldg(source + i);
        }
    }
    __syncthreads(); // the libm tables, the table barrier's initialisation

    const int lane = threadIdx.x & 31;
    const int warpInBlock = threadIdx.x >> 5;
    uint32_t* myStage = stageAll + warpInBlock * (32 * kLaneStrideWords) + lane * kLaneStrideWords;
    const uint32_t flatShift = p.table.flatShift;
    const int32_t negativeLow = -static_cast<int32_t>(p.table.flatLow);
    const int32_t span = static_cast<int32_t>(p.table.flatHigh - p.table.flatLow);
    const uint32_t bandStrideLog2 = p.table.bandStrideLog2;
    const uint32_t* __restrict__ bandBits = p.table.bandBits;
    const uint32_t compactTopShift = 32u - flatShift;
    const uint32_t compactCodeMask = p.table.compactCodeMask;
    const uint32_t compactMagic = p.table.compactMagic;

    const int tilesX = (p.width + kTilePixels - 1) / kTilePixels;
    const int tileRows = (p.rowCount + 1) / 2;
    const int tileCount = tilesX * tileRows;
    const int warpCount = static_cast<int>(gridDim.x) * kRgbaWarps;
    const int firstTile = static_cast<int>(blockIdx.x) * kRgbaWarps + warpInBlock;
    const int stepRows = warpCount / tilesX;
    const int stepX = warpCount - stepRows * tilesX;
    int tileRow = firstTile / tilesX;
    int tileX = firstTile - tileRow * tilesX;

    uint4 raw[8]; // pixel i of row r = raw[4 * r + i] = { R, G, B, A }
    auto loadTile = [&](int row, int column, bool valid)
    {
        const int x = column * kTilePixels + lane * 4;
        const int y = row * 2;
        const bool active = valid && x < p.width;
        const bool second = active && (y + 1) < p.rowCount;
        const uint8_t* r0 = p.rows + static_cast<int64_t>(y) * p.rowStride + static_cast<int64_t>(x) * 16;
        const uint8_t* r1 = r0 + p.rowStride;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            raw[q] = active ? __ldg(reinterpret_cast<const uint4*>(r0 + 16 * q)) : zero;
            raw[4 + q] = second ? __ldg(reinterpret_cast<const uint4*>(r1 + 16 * q)) : zero;
        }
    };
    loadTile(tileRow, tileX, firstTile < tileCount); // in flight while the table image lands
    if (COMPACT)
    {
        staging::WaitTableImage(tableBarrier);
    }

#pragma unroll 1
    for (int tile = firstTile; tile < tileCount; tile += warpCount)
    {
        const int x0 = tileX * kTilePixels + lane * 4;
        const int y0 = tileRow * 2;
        const bool laneActive = x0 < p.width;
        const bool secondRow = (y0 + 1) < p.rowCount;
        int nextRow = tileRow + stepRows;
        int nextX = tileX + stepX;
        if (nextX >= tilesX)
        {
            nextX -= tilesX;
            ++nextRow;
        }

        // ---- alpha, clamp / premultiplication (WriteHeifImage.cpp:1043-1077), then the colour samples the curve sees ----
        uint32_t alphaCode[8];
        uint32_t colourBits[kValuesPerLane];
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            const float alpha = ClampF(__uint_as_float(raw[k].w), 0.0f, 1.0f);
            float colour[3] = { __uint_as_float(raw[k].x), __uint_as_float(raw[k].y), __uint_as_float(raw[k].z) };
            if (p.premultiply && alpha < 1.0f)
            {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                {
                    colour[c] = (alpha == 0) ? 0.0f : PremultiplyColor(ClampF(colour[c], 0.0f, 1.0f), alpha, 1.0f);
                }
            }
            alphaCode[k] = FloatToCode(alpha, p.maxCodeFloat);
#pragma unroll
            for (int c = 0; c < 3; ++c)
            {
                colourBits[3 * k + c] = __float_as_uint(colour[c]);
            }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q)
        {
            *reinterpret_cast<uint4*>(myStage + 4 * q) = make_uint4(colourBits[4 * q], colourBits[4 * q + 1], colourBits[4 * q + 2], colourBits[4 * q + 3]);
        }

        // ---- float -> code through the exact step table ------------------------------------------------------------
        float codeF[kValuesPerLane];
        uint32_t bandMask = 0;
        int32_t largest = 0;
#pragma unroll
        for (int j = 0; j < kValuesPerLane; ++j)
        {
            bool inBand;
            if (COMPACT)
            {
                uint32_t entry;
                codeF[j] = LookupCurveCompact<0>(colourBits[j], compactEntries, flatShift, negativeLow, span, compactTopShift, compactCodeMask, compactMagic, inBand, entry);
            }
            else
            {
                codeF[j] = LookupCurveFlat(colourBits[j], flatEntries, flatShift, negativeLow, span, inBand);
            }
            asm("{ .reg .pred q; setp.ne.u32 q, %1, 0; @q or.b32 %0, %0, %2; }" : "+r"(bandMask) : "r"(static_cast<uint32_t>(inBand)), "r"(1u << j));
            largest = max(largest, static_cast<int32_t>(colourBits[j]));
        }
        loadTile(nextRow, nextX, tile + warpCount < tileCount);

        // ---- in-band samples: one bit of the band bitmap each ---------------------------------------------------------
        uint32_t lowerMask = 0;
        {
            uint32_t pending = bandMask;
            while (pending != 0)
            {
                const int j = __ffs(static_cast<int>(pending)) - 1;
                pending &= pending - 1;
                const uint32_t bits = myStage[j];
                if (COMPACT)
                {
                    const int32_t bucket = __viaddmin_s32_relu(static_cast<int32_t>(bits) >> flatShift, negativeLow, span);
                    const uint32_t entry = compactEntries[bucket];
                    const uint32_t k = ((entry & compactCodeMask) >> kCompactLenBits) + ((entry >> compactTopShift) != 0 ? 1u : 0u);
                    const uint32_t distance = bits - firstBits[k];
                    if (k != 0 && distance < (1u << bandStrideLog2)) // else flagged by the superset test only
                    {
                        const uint32_t bitIndex = (k << bandStrideLog2) + distance;
                        const uint32_t word = __ldg(bandBits + (bitIndex >> 5));
                        lowerMask |= (((word >> (bitIndex & 31u)) & 1u) ^ 1u) << j;
                    }
                    continue;
                }
                bool inBand;
                uint2 entry;
                LookupCurveFlat(bits, flatEntries, flatShift, negativeLow, span, inBand, entry);
                const uint32_t bitIndex = BandBitIndex(bits, entry, bandStrideLog2);
                const uint32_t word = __ldg(bandBits + (bitIndex >> 5));
                lowerMask |= (((word >> (bitIndex & 31u)) & 1u) ^ 1u) << j;
            }
        }
        // ---- +inf / NaN: the exact evaluation, lane by lane -------------------------------------------------------------
        if (__any_sync(0xffffffffu, largest > 0x7f7fffff))
        {
            for (int j = 0; j < kValuesPerLane; ++j)
            {
                const uint32_t bits = myStage[j];
                if (static_cast<int32_t>(bits) > 0x7f7fffff)
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
            const int64_t chromaRow = YS ? tileRow : y0;
            const int64_t chromaColumn = static_cast<int64_t>(XS ? (x0 >> 1) : x0) * 2;
            StoreTile<XS, YS>(p, codeF, p.planeY + static_cast<int64_t>(y0) * p.strideY + static_cast<int64_t>(x0) * 2,
                              p.planeCb + chromaRow * p.strideCb + chromaColumn, p.planeCr + chromaRow * p.strideCr + chromaColumn, secondRow);
            uint8_t* alphaRow = p.planeA + static_cast<int64_t>(y0) * p.strideA + static_cast<int64_t>(x0) * 2;
            __stcs(reinterpret_cast<uint2*>(alphaRow), make_uint2(alphaCode[0] | (alphaCode[1] << 16), alphaCode[2] | (alphaCode[3] << 16)));
            if (secondRow)
            {
                __stcs(reinterpret_cast<uint2*>(alphaRow + p.strideA), make_uint2(alphaCode[4] | (alphaCode[5] << 16), alphaCode[6] | (alphaCode[7] << 16)));
            }
        }
        tileRow = nextRow;
        tileX = nextX;
    }
}

size_t RgbaTableBytes(const FastEncodeParams& fp, bool compact)
{
    return compact ? static_cast<size_t>(fp.table.compactImageBytes) : static_cast<size_t>((fp.table.flatCount + 1) / 2) * sizeof(uint4);
}

template <int CURVE, int XS, int YS, int COMPACT>
cudaError_t LaunchRgbaKernel(const FastEncodeParams& fp, int smCount, cudaStream_t stream)
{
    const size_t shared = static_cast<size_t>(RgbaFixedBytes()) + RgbaTableBytes(fp, COMPACT != 0);
    static std::atomic<uint64_t> configuredDevices{ 0 }; // per instantiation
    {
        const cudaError_t e = AllowDynamicShared(EncodeRgbaF32FlatKernel<CURVE, XS, YS, COMPACT>, kSharedLimit, configuredDevices);
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
    long long blocks = (tiles + kRgbaWarps - 1) / kRgbaWarps;
    if (blocks > smCount)
    {
        blocks = smCount;
    }
    EncodeRgbaF32FlatKernel<CURVE, XS, YS, COMPACT><<<static_cast<unsigned>(blocks), kRgbaThreads, shared, stream>>>(fp);
    return cudaGetLastError();
}

template <int CURVE, int COMPACT>
cudaError_t DispatchRgbaChroma(const FastEncodeParams& fp, int xs, int ys, int smCount, cudaStream_t stream)
{
    if (xs == 1 && ys == 1) return LaunchRgbaKernel<CURVE, 1, 1, COMPACT>(fp, smCount, stream);
    if (xs == 1) return LaunchRgbaKernel<CURVE, 1, 0, COMPACT>(fp, smCount, stream);
    return LaunchRgbaKernel<CURVE, 0, 0, COMPACT>(fp, smCount, stream);
}

} // namespace

bool RgbaEncodeApplies(const FastEncodeParams& fp)
{
    return fp.planeA != nullptr && fp.table.flat != nullptr && fp.table.bandBits != nullptr &&
           static_cast<size_t>(RgbaFixedBytes()) + static_cast<size_t>((fp.table.flatCount + 1) / 2) * sizeof(uint4) <= static_cast<size_t>(kSharedLimit);
}

cudaError_t LaunchFastEncodeRgba(const FastEncodeParams& fp, int curve, int xs, int ys, int smCount, cudaStream_t stream)
{
    if (fp.table.compact != nullptr && fp.table.firstBits != nullptr && !fp.preferWideEntries)
    {
        if (curve == kCurveLinearToPQ) return DispatchRgbaChroma<kCurveLinearToPQ, 1>(fp, xs, ys, smCount, stream);
        return DispatchRgbaChroma<kCurveLinearToSMPTE428, 1>(fp, xs, ys, smCount, stream);
    }
    if (curve == kCurveLinearToPQ) return DispatchRgbaChroma<kCurveLinearToPQ, 0>(fp, xs, ys, smCount, stream);
    return DispatchRgbaChroma<kCurveLinearToSMPTE428, 0>(fp, xs, ys, smCount, stream);
}

} // namespace avifgpu
