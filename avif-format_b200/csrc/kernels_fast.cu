// kernels_fast.cu -- tuned kernels for the layouts BASELINE.json measures.  Anything they do not cover falls back
// to kernels_generic.cu (LaunchEncodeFast / LaunchDecodeFast return 0 = "not applicable").  The arithmetic is the
// same pixel_math.cuh; what changes is the work decomposition and where the transcendental work goes.
//
// Encode, float RGB -> planar YCbCr: dispatch, plus the kernel for the cases kernels_fast_flat.cu does not take
// (no transfer curve, or a curve whose step table only has the two-level form, e.g. SMPTE 428):
//   * one warp converts a tile of 2 rows x 128 pixels; a lane owns 4 adjacent pixels in both rows = two 2x2
//     chroma sites, so the 4:2:0 box filter needs no cross-lane traffic at all;
//   * loads: 3 x LDG.128 per row per lane (48 contiguous bytes), issued one tile ahead; stores: Y 8 bytes per row
//     per lane, Cb / Cr 4 bytes per lane;
//   * float -> code goes through the two-level step table of curve_tables.h (two shared-memory look-ups per
//     sample); samples inside a fuzzy band are queued per warp and evaluated with the exact glibc-identical powf
//     at full lane occupancy (warp-level compaction), then patched back;
//   * persistent grid: 2 CTAs of 8 warps per SM, warps stride over the tiles.
#include "kernels_fast_common.cuh"
#include "../../include/avifgpu.h"

#include <cuda_runtime.h>


namespace avifgpu
{

using namespace avifpix;
using namespace fastenc;
using avifmath::LibmTables;

namespace
{

constexpr int kLaneStrideWords = 28; // staging stride per lane (16-byte aligned, conflict-free for STS.128)

// Shared-memory carve-up (bytes).
constexpr int kSharedOctaves = 2048;
constexpr int kSharedStagePerWarp = 32 * kLaneStrideWords * 4;   // sample bits, later the exact codes
constexpr int kQueueCapacity = 256;                              // in-band samples per exact-path round (uint16 slots)
constexpr int kSharedQueuePerWarp = kQueueCapacity * 2;
__host__ __device__ constexpr int SharedFixedBytes(int warps) { return kSharedLibm + kSharedOctaves + warps * (kSharedStagePerWarp + kSharedQueuePerWarp); }

struct FastConfig
{
    static constexpr int threads = 256;
    static constexpr int warps = threads / 32;
    static constexpr int blocksPerSm = 2;
    static constexpr int sharedLimit = 112 * 1024;
};

template <int CURVE, int XS, int YS>
__global__ void __launch_bounds__(FastConfig::threads, FastConfig::blocksPerSm) EncodeRgbF32PlanarKernel(const FastEncodeParams p)
{
    constexpr int kFastWarps = FastConfig::warps;
    constexpr int kSharedFixed = SharedFixedBytes(kFastWarps);
    extern __shared__ __align__(16) uint8_t sharedBytes[];
    uint64_t* libmStorage = reinterpret_cast<uint64_t*>(sharedBytes);
    uint2* octaves = reinterpret_cast<uint2*>(sharedBytes + kSharedLibm);
    uint32_t* stageAll = reinterpret_cast<uint32_t*>(sharedBytes + kSharedLibm + kSharedOctaves);
    uint16_t* queueAll = reinterpret_cast<uint16_t*>(sharedBytes + kSharedLibm + kSharedOctaves + kFastWarps * kSharedStagePerWarp);
    uint32_t* tableWords = reinterpret_cast<uint32_t*>(sharedBytes + kSharedFixed);

    const LibmTables t = avifmath::StageLibmTables(libmStorage, threadIdx.x, blockDim.x);
    if (CURVE != kCurveClip)
    {
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
        {
            octaves[i] = p.table.octaves[i];
        }
        for (int i = threadIdx.x; i < p.table.bucketCount; i += blockDim.x)
        {
            tableWords[i] = p.table.buckets[i];
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warpInBlock = threadIdx.x >> 5;
    uint32_t* stage = stageAll + warpInBlock * (32 * kLaneStrideWords);
    uint32_t* myStage = stage + lane * kLaneStrideWords;
    uint16_t* queue = queueAll + warpInBlock * kQueueCapacity;

    const int tilesX = (p.width + kTilePixels - 1) / kTilePixels;
    const int tileRows = (p.rowCount + 1) / 2;
    const int tileCount = tilesX * tileRows;
    const int warpCount = static_cast<int>(gridDim.x) * kFastWarps;

    // Tile coordinates advance incrementally (an integer division per tile costs ~24 instructions).
    const int firstTile = static_cast<int>(blockIdx.x) * kFastWarps + warpInBlock;
    const int stepRows = warpCount / tilesX;
    const int stepX = warpCount - stepRows * tilesX;
    int tileRow = firstTile / tilesX;
    int tileX = firstTile - tileRow * tilesX;
    // Software pipeline: the six 128-bit loads of tile i+1 are issued as soon as the look-ups of tile i have
    // consumed the registers, so they are in flight during the exact path, the matrix and the stores of tile i.
    uint4 raw[6];
    auto loadTile = [&](int row, int column, bool valid)
    {
        const int x = column * kTilePixels + lane * 4;
        const int y = row * 2;
        const bool active = valid && x < p.width;
        const bool second = active && (y + 1) < p.rowCount;
        const uint8_t* r0 = p.rows + static_cast<int64_t>(y) * p.rowStride + static_cast<int64_t>(x) * 12;
        const uint8_t* r1 = r0 + p.rowStride;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < 3; ++q)
        {
            raw[q] = active ? __ldg(reinterpret_cast<const uint4*>(r0 + 16 * q)) : zero;
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
        {
            raw[3 + q] = second ? __ldg(reinterpret_cast<const uint4*>(r1 + 16 * q)) : zero;
        }
    };
    loadTile(tileRow, tileX, firstTile < tileCount);
    auto prefetchNext = [&](int tile)
    {
        int nextRow = tileRow + stepRows;
        int nextX = tileX + stepX;
        if (nextX >= tilesX)
        {
            nextX -= tilesX;
            ++nextRow;
        }
        loadTile(nextRow, nextX, tile + warpCount < tileCount);
    };

    for (int tile = firstTile; tile < tileCount; tile += warpCount, tileRow += stepRows, tileX += stepX)
    {
        if (tileX >= tilesX)
        {
            tileX -= tilesX;
            ++tileRow;
        }
        const int x0 = tileX * kTilePixels + lane * 4;
        const int y0 = tileRow * 2;
        const bool laneActive = x0 < p.width;
        const bool secondRow = (y0 + 1) < p.rowCount;

        float codeF[kValuesPerLane]; // the codes, as the floats the forward matrix consumes

        if (CURVE == kCurveClip)
        {
#pragma unroll
            for (int q = 0; q < 6; ++q)
            {
                codeF[4 * q + 0] = CodeToFloat(FloatToCode(__uint_as_float(raw[q].x), p.maxCodeFloat));
                codeF[4 * q + 1] = CodeToFloat(FloatToCode(__uint_as_float(raw[q].y), p.maxCodeFloat));
                codeF[4 * q + 2] = CodeToFloat(FloatToCode(__uint_as_float(raw[q].z), p.maxCodeFloat));
                codeF[4 * q + 3] = CodeToFloat(FloatToCode(__uint_as_float(raw[q].w), p.maxCodeFloat));
            }
            prefetchNext(tile);
        }
        else
        {
            // ---- two-level table; in-band samples go to the exact evaluation, compacted across the warp ---------
            uint32_t code[kValuesPerLane];
#pragma unroll
            for (int q = 0; q < 6; ++q)
            {
                *reinterpret_cast<uint4*>(myStage + 4 * q) = raw[q];
            }
            uint32_t bandMask = 0;
#pragma unroll
            for (int j = 0; j < kValuesPerLane; ++j)
            {
                const uint4 w = raw[j >> 2];
                const uint32_t bits = (j & 3) == 0 ? w.x : (j & 3) == 1 ? w.y : (j & 3) == 2 ? w.z : w.w;
                bool inBand;
                code[j] = LookupCurveCode(bits, octaves, tableWords, inBand);
                asm("{ .reg .pred p; setp.ne.u32 p, %1, 0; @p or.b32 %0, %0, %2; }" : "+r"(bandMask) : "r"(static_cast<uint32_t>(inBand)), "r"(1u << j));
            }
            prefetchNext(tile);
            // Warp-level compaction: exclusive prefix sum of the per-lane counts gives every lane its queue range.
            const int mine = __popc(bandMask);
            int inclusive = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const int up = __shfl_up_sync(0xffffffffu, inclusive, d);
                inclusive += (lane >= d) ? up : 0;
            }
            const int total = __shfl_sync(0xffffffffu, inclusive, 31);
            if (total > 0)
            {
                // Rounds of at most kQueueCapacity samples (one round unless a whole tile sits inside fuzzy bands).
                for (int roundStart = 0; roundStart < total; roundStart += kQueueCapacity)
                {
                    int position = inclusive - mine - roundStart;
                    uint32_t pending = bandMask;
                    while (pending != 0)
                    {
                        const int j = __ffs(static_cast<int>(pending)) - 1;
                        pending &= pending - 1;
                        if (position >= 0 && position < kQueueCapacity)
                        {
                            queue[position] = static_cast<uint16_t>(lane * kLaneStrideWords + j);
                        }
                        ++position;
                    }
                    __syncwarp();
                    const int count = min(total - roundStart, kQueueCapacity);
#pragma unroll 1
                    for (int q = lane; q < count; q += 32)
                    {
                        const uint32_t slot = queue[q];
                        stage[slot] = ExactCurveCode<CURVE == kCurveClip ? kCurveLinearToPQ : CURVE>(__uint_as_float(stage[slot]), p.pqMultiplier,
                                                                                                  p.maxCodeFloat, t);
                    }
                    __syncwarp();
                }
                if (bandMask != 0)
                {
#pragma unroll
                    for (int j = 0; j < kValuesPerLane; ++j)
                    {
                        if (bandMask & (1u << j))
                        {
                            code[j] = myStage[j];
                        }
                    }
                }
                __syncwarp(); // the staging area is rewritten by the next tile
            }
#pragma unroll
            for (int j = 0; j < kValuesPerLane; ++j)
            {
                codeF[j] = CodeToFloat(code[j]);
            }
        }

        if (!laneActive)
        {
            continue;
        }

        {
            const int64_t chromaRow = YS ? tileRow : y0;
            const int64_t chromaColumn = static_cast<int64_t>(XS ? (x0 >> 1) : x0) * 2;
            StoreTile<XS, YS>(p, codeF, p.planeY + static_cast<int64_t>(y0) * p.strideY + static_cast<int64_t>(x0) * 2,
                              p.planeCb + chromaRow * p.strideCb + chromaColumn, p.planeCr + chromaRow * p.strideCr + chromaColumn, secondRow);
        }
    }
}

template <int CURVE, int XS, int YS>
size_t FastEncodeSharedBytes(const FastEncodeParams& fp)
{
    const size_t tableBytes = CURVE == kCurveClip ? 0 : static_cast<size_t>(fp.table.bucketCount) * sizeof(uint32_t);
    return static_cast<size_t>(SharedFixedBytes(FastConfig::warps)) + tableBytes;
}

template <int CURVE, int XS, int YS>
cudaError_t LaunchFastEncodeKernel(const FastEncodeParams& fp, int smCount, cudaStream_t stream)
{
    using Config = FastConfig;
    const size_t shared = FastEncodeSharedBytes<CURVE, XS, YS>(fp);
    static std::atomic<uint64_t> configuredDevices{ 0 }; // per instantiation
    {
        const cudaError_t e = AllowDynamicShared(EncodeRgbF32PlanarKernel<CURVE, XS, YS>, Config::sharedLimit, configuredDevices);
        if (e != cudaSuccess)
        {
            return e;
        }
    }
    if (shared > static_cast<size_t>(Config::sharedLimit))
    {
        return cudaErrorInvalidValue;
    }
    const long long tiles = static_cast<long long>((fp.width + kTilePixels - 1) / kTilePixels) * ((fp.rowCount + 1) / 2);
    if (tiles > 0x7fffffffll)
    {
        return cudaErrorInvalidValue;
    }
    long long blocks = (tiles + Config::warps - 1) / Config::warps;
    const long long resident = static_cast<long long>(smCount) * Config::blocksPerSm;
    if (blocks > resident)
    {
        blocks = resident;
    }
    EncodeRgbF32PlanarKernel<CURVE, XS, YS><<<static_cast<unsigned>(blocks), Config::threads, shared, stream>>>(fp);
    return cudaGetLastError();
}

template <int CURVE>
cudaError_t DispatchChroma(const FastEncodeParams& fp, int xs, int ys, int smCount, cudaStream_t stream)
{
    if (xs == 1 && ys == 1) return LaunchFastEncodeKernel<CURVE, 1, 1>(fp, smCount, stream);
    if (xs == 1) return LaunchFastEncodeKernel<CURVE, 1, 0>(fp, smCount, stream);
    return LaunchFastEncodeKernel<CURVE, 0, 0>(fp, smCount, stream);
}

template <int CURVE>
cudaError_t DispatchTable(const FastEncodeParams& fp, int xs, int ys, int smCount, cudaStream_t stream)
{
    if (FlatEncodeApplies(fp))
    {
        return LaunchFastEncodeFlat(fp, CURVE, xs, ys, smCount, stream);
    }
    return DispatchChroma<CURVE>(fp, xs, ys, smCount, stream);
}

bool Aligned(const void* p, int64_t stride, int alignment)
{
    return (reinterpret_cast<uintptr_t>(p) % alignment) == 0 && (stride % alignment) == 0;
}

} // namespace

int LaunchEncodeGeneric(const EncodeParams& params, int hostDepth, void* stream);

// Returns the number of kernels launched, 0 if this configuration is not covered, or a negative status.
int LaunchEncodeFast(const EncodeParams& p, int hostDepth, void* streamHandle)
{
    cudaStream_t stream = static_cast<cudaStream_t>(streamHandle);
    const bool rgba = p.channels == 4 && p.hasAlpha;
    if (hostDepth != 32 || !p.planar || (p.channels != 3 && !rgba) || (p.channels == 3 && p.hasAlpha) || p.imageDepth <= 8)
    {
        return 0;
    }
    int curve;
    if (p.transfer == AVIFGPU_TRANSFER_PQ) curve = kCurveLinearToPQ;
    else if (p.transfer == AVIFGPU_TRANSFER_SMPTE428) curve = kCurveLinearToSMPTE428;
    else if (p.transfer == AVIFGPU_TRANSFER_CLIP) curve = kCurveClip;
    else return 0; // HLG save path: the generic kernel (with the context's step table when it has one)
    if (curve != kCurveClip && (p.curveTable == nullptr || p.curveTable->buckets == nullptr))
    {
        return 0; // no verified table for this curve: the generic exact kernel serves it
    }
    if (!ForwardMatrixStaysInRange(p.matrix, p.chromaOffset, static_cast<int>(p.maxCode)))
    {
        return 0; // an exotic matrix: the generic kernel clamps
    }
    if (!Aligned(p.rows, p.rowStride, 16) || !Aligned(p.plane[0], p.planeStride[0], 8) ||
        !Aligned(p.plane[1], p.planeStride[1], p.xs ? 4 : 8) || !Aligned(p.plane[2], p.planeStride[2], p.xs ? 4 : 8) ||
        (rgba && !Aligned(p.plane[3], p.planeStride[3], 8)))
    {
        return 0;
    }
    if (rgba && (curve == kCurveClip || p.curveTable->flat == nullptr || p.curveTable->bandBits == nullptr))
    {
        return 0; // the RGBA kernel is built on the flat table + band bitmap; everything else with alpha: generic
    }
    const int width4 = p.width & ~3;
    const int evenRows = p.ys ? (p.rowCount & ~1) : p.rowCount;
    if (width4 < 4 || evenRows < 1)
    {
        return 0;
    }

    FastEncodeParams fp{};
    fp.rows = static_cast<const uint8_t*>(p.rows);
    fp.rowStride = p.rowStride;
    fp.planeY = static_cast<uint8_t*>(p.plane[0]);
    fp.strideY = p.planeStride[0];
    fp.planeCb = static_cast<uint8_t*>(p.plane[1]);
    fp.strideCb = p.planeStride[1];
    fp.planeCr = static_cast<uint8_t*>(p.plane[2]);
    fp.strideCr = p.planeStride[2];
    fp.width = width4;
    fp.rowCount = evenRows;
    fp.pqMultiplier = p.pqMultiplier;
    fp.maxCodeFloat = p.maxCodeFloat;
    fp.maxCode = static_cast<int32_t>(p.maxCode);
    fp.matrix = p.matrix;
    fp.chromaOffset = p.chromaOffset;
    fp.topLeft = p.topLeft;
    if (curve != kCurveClip)
    {
        fp.table = *p.curveTable;
    }
    if (rgba)
    {
        fp.planeA = static_cast<uint8_t*>(p.plane[3]);
        fp.strideA = p.planeStride[3];
        fp.premultiply = p.premultiply;
    }

    const int smCount = p.smCount > 0 ? p.smCount : 148;
    cudaError_t e;
    if (rgba)
    {
        if (!RgbaEncodeApplies(fp))
        {
            return 0;
        }
        e = LaunchFastEncodeRgba(fp, curve, p.xs, p.ys, smCount, stream);
    }
    else if (curve == kCurveLinearToPQ) e = DispatchTable<kCurveLinearToPQ>(fp, p.xs, p.ys, smCount, stream);
    else if (curve == kCurveLinearToSMPTE428) e = DispatchTable<kCurveLinearToSMPTE428>(fp, p.xs, p.ys, smCount, stream);
    else e = DispatchChroma<kCurveClip>(fp, p.xs, p.ys, smCount, stream);
    if (e != cudaSuccess)
    {
        return AVIFGPU_ERR_CUDA;
    }
    int launched = 1;

    // Edges the tile kernel does not cover go through the generic kernel as sub-rectangles: the right strip
    // (width % 4 columns, all rows) and, for vertically sub-sampled chroma, an odd last row.
    if (width4 < p.width)
    {
        EncodeParams strip = p;
        strip.rows = static_cast<const uint8_t*>(p.rows) + static_cast<int64_t>(width4) * (rgba ? 16 : 12);
        strip.width = p.width - width4;
        strip.plane[0] = static_cast<uint8_t*>(p.plane[0]) + static_cast<int64_t>(width4) * 2;
        strip.plane[1] = static_cast<uint8_t*>(p.plane[1]) + static_cast<int64_t>(width4 >> p.xs) * 2;
        strip.plane[2] = static_cast<uint8_t*>(p.plane[2]) + static_cast<int64_t>(width4 >> p.xs) * 2;
        if (rgba) strip.plane[3] = static_cast<uint8_t*>(p.plane[3]) + static_cast<int64_t>(width4) * 2;
        const int n = LaunchEncodeGeneric(strip, hostDepth, streamHandle);
        if (n < 0) return n;
        launched += n;
    }
    if (evenRows < p.rowCount)
    {
        EncodeParams strip = p;
        strip.rows = static_cast<const uint8_t*>(p.rows) + static_cast<int64_t>(evenRows) * p.rowStride;
        strip.rowCount = p.rowCount - evenRows;
        strip.width = width4;
        strip.plane[0] = static_cast<uint8_t*>(p.plane[0]) + static_cast<int64_t>(evenRows) * p.planeStride[0];
        strip.plane[1] = static_cast<uint8_t*>(p.plane[1]) + static_cast<int64_t>(evenRows >> p.ys) * p.planeStride[1];
        strip.plane[2] = static_cast<uint8_t*>(p.plane[2]) + static_cast<int64_t>(evenRows >> p.ys) * p.planeStride[2];
        if (rgba) strip.plane[3] = static_cast<uint8_t*>(p.plane[3]) + static_cast<int64_t>(evenRows) * p.planeStride[3];
        const int n = LaunchEncodeGeneric(strip, hostDepth, streamHandle);
        if (n < 0) return n;
        launched += n;
    }
    return launched;
}

} // namespace avifgpu
