// kernels_fast.cu -- tuned kernels for the layouts BASELINE.json measures.  Anything they do not cover falls back
// to kernels_generic.cu (LaunchEncodeFast / LaunchDecodeFast return 0 = "not applicable").  The arithmetic is the
// same pixel_math.cuh; what changes is the work decomposition and where the transcendental work goes.
//
// Encode, float RGB -> planar YCbCr (BASELINE config 2: 7680x4320 RGB32f -> 12-bit PQ 4:2:0)
// -----------------------------------------------------------------------------------------------
//   * one warp converts a tile of 2 rows x 128 pixels; a lane owns 4 adjacent pixels in both rows = two 2x2
//     chroma sites, so the 4:2:0 box filter needs no cross-lane traffic at all;
//   * loads: 3 x LDG.128 per row per lane (48 contiguous bytes), a warp reads 1536 contiguous bytes per row;
//     stores: Y 8 bytes per row per lane (256 contiguous bytes per warp), Cb / Cr 4 bytes per lane;
//   * float -> code goes through the exact step tables of curve_tables.h (two shared-memory look-ups per
//     sample); samples inside a fuzzy band are queued per warp and evaluated with the exact glibc-identical powf
//     at full lane occupancy (warp-level compaction), then patched back;
//   * persistent grid: 2 CTAs of 8 warps per SM, warps stride over the tiles.
#include "kernel_params.h"
#include "curve_lookup.cuh"
#include "../../include/avifgpu.h"

#include <cuda_runtime.h>

namespace avifgpu
{

using namespace avifpix;
using avifmath::LibmTables;

namespace
{

constexpr int kFastThreads = 256;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kTilePixels = 128;    // per row
constexpr int kValuesPerLane = 24;  // 2 rows x 4 pixels x 3 channels
constexpr int kQueueCapacity = 128; // entries per warp between flushes
constexpr int kCurveClip = 2;       // no transfer curve: code = trunc(clamp(v * max))

struct FastEncodeParams
{
    const uint8_t* rows;
    int64_t rowStride;
    uint8_t* planeY;
    int64_t strideY;
    uint8_t* planeCb;
    int64_t strideCb;
    uint8_t* planeCr;
    int64_t strideCr;
    int32_t width;    // multiple of 4
    int32_t rowCount; // even when the chroma is vertically sub-sampled
    float pqMultiplier;
    float maxCodeFloat;
    int32_t maxCode;
    ForwardMatrix matrix;
    float chromaOffset;
    int32_t topLeft;
    CurveTableView table;
};

struct QueueEntry
{
    uint32_t bits;
    uint32_t slot;
};

__device__ __forceinline__ float4 LoadRow4(const uint8_t* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int CURVE, int XS, int YS>
__global__ void __launch_bounds__(kFastThreads, 2) EncodeRgbF32PlanarKernel(const FastEncodeParams p)
{
    extern __shared__ __align__(16) uint8_t sharedBytes[];
    // layout: libm tables (768 B) | octaves (2048 B) | per-warp queues | per-warp results | buckets
    uint64_t* libmStorage = reinterpret_cast<uint64_t*>(sharedBytes);
    uint2* octaves = reinterpret_cast<uint2*>(sharedBytes + 768);
    QueueEntry* queues = reinterpret_cast<QueueEntry*>(sharedBytes + 768 + 2048);
    uint16_t* results = reinterpret_cast<uint16_t*>(sharedBytes + 768 + 2048 + kFastWarps * kQueueCapacity * sizeof(QueueEntry));
    uint32_t* buckets = reinterpret_cast<uint32_t*>(sharedBytes + 768 + 2048 + kFastWarps * kQueueCapacity * sizeof(QueueEntry) +
                                                    kFastWarps * 32 * kValuesPerLane * sizeof(uint16_t));

    const LibmTables t = avifmath::StageLibmTables(libmStorage, threadIdx.x, blockDim.x);
    if (CURVE != kCurveClip)
    {
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
        {
            octaves[i] = p.table.octaves[i];
        }
        for (int i = threadIdx.x; i < p.table.bucketCount; i += blockDim.x)
        {
            buckets[i] = p.table.buckets[i];
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warpInBlock = threadIdx.x >> 5;
    QueueEntry* queue = queues + warpInBlock * kQueueCapacity;
    uint16_t* result = results + warpInBlock * 32 * kValuesPerLane;

    const int tilesX = (p.width + kTilePixels - 1) / kTilePixels;
    const int tileRows = (p.rowCount + 1) / 2;
    const long long tileCount = static_cast<long long>(tilesX) * tileRows;
    const long long warpCount = static_cast<long long>(gridDim.x) * kFastWarps;

    for (long long tile = static_cast<long long>(blockIdx.x) * kFastWarps + warpInBlock; tile < tileCount; tile += warpCount)
    {
        const int tileRow = static_cast<int>(tile / tilesX);
        const int tileX = static_cast<int>(tile - static_cast<long long>(tileRow) * tilesX);
        const int x0 = tileX * kTilePixels + lane * 4;
        const int y0 = tileRow * 2;
        const bool laneActive = x0 < p.width;
        const bool secondRow = (y0 + 1) < p.rowCount;

        // ---- load 2 rows x 4 pixels x RGB ------------------------------------------------------------------
        float v[kValuesPerLane];
        if (laneActive)
        {
            const uint8_t* r0 = p.rows + static_cast<int64_t>(y0) * p.rowStride + static_cast<int64_t>(x0) * 12;
            const float4 a0 = LoadRow4(r0), a1 = LoadRow4(r0 + 16), a2 = LoadRow4(r0 + 32);
            v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
            v[8] = a2.x; v[9] = a2.y; v[10] = a2.z; v[11] = a2.w;
            if (secondRow)
            {
                const uint8_t* r1 = r0 + p.rowStride;
                const float4 b0 = LoadRow4(r1), b1 = LoadRow4(r1 + 16), b2 = LoadRow4(r1 + 32);
                v[12] = b0.x; v[13] = b0.y; v[14] = b0.z; v[15] = b0.w; v[16] = b1.x; v[17] = b1.y; v[18] = b1.z; v[19] = b1.w;
                v[20] = b2.x; v[21] = b2.y; v[22] = b2.z; v[23] = b2.w;
            }
            else
            {
#pragma unroll
                for (int j = 12; j < 24; ++j) v[j] = 0.0f;
            }
        }
        else
        {
#pragma unroll
            for (int j = 0; j < 24; ++j) v[j] = 0.0f;
        }

        // ---- float -> code: table, with in-band samples queued for the exact path ---------------------------
        uint32_t code[kValuesPerLane];
        uint32_t bandMask = 0;
        int queued = 0; // warp-uniform
#pragma unroll
        for (int j = 0; j < kValuesPerLane; ++j)
        {
            if (CURVE == kCurveClip)
            {
                code[j] = FloatToCode(v[j], p.maxCodeFloat);
            }
            else
            {
                const uint32_t bits = __float_as_uint(v[j]);
                bool inBand;
                code[j] = LookupCurveCode(bits, octaves, buckets, inBand);
                const uint32_t ballot = __ballot_sync(0xffffffffu, inBand);
                if (ballot != 0)
                {
                    const int count = __popc(ballot);
                    if (queued + count > kQueueCapacity)
                    {
                        // flush: evaluate what is queued with all lanes busy
                        __syncwarp();
                        for (int q = lane; q < queued; q += 32)
                        {
                            const QueueEntry entry = queue[q];
                            result[entry.slot] = static_cast<uint16_t>(
                                ExactCurveCode<CURVE == kCurveClip ? kCurveLinearToPQ : CURVE>(__uint_as_float(entry.bits), p.pqMultiplier, p.maxCodeFloat, t));
                        }
                        __syncwarp();
                        queued = 0;
                    }
                    if (inBand)
                    {
                        const int position = queued + __popc(ballot & ((1u << lane) - 1u));
                        QueueEntry entry;
                        entry.bits = bits;
                        entry.slot = static_cast<uint32_t>(lane * kValuesPerLane + j);
                        queue[position] = entry;
                        bandMask |= 1u << j;
                    }
                    queued += count;
                }
            }
        }
        if (CURVE != kCurveClip)
        {
            if (queued > 0)
            {
                __syncwarp();
                for (int q = lane; q < queued; q += 32)
                {
                    const QueueEntry entry = queue[q];
                    result[entry.slot] = static_cast<uint16_t>(
                        ExactCurveCode<CURVE == kCurveClip ? kCurveLinearToPQ : CURVE>(__uint_as_float(entry.bits), p.pqMultiplier, p.maxCodeFloat, t));
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
                        code[j] = result[lane * kValuesPerLane + j];
                    }
                }
            }
            __syncwarp(); // results are rewritten by the next tile
        }

        if (!laneActive)
        {
            continue;
        }

        // ---- forward matrix, luma quantisation, chroma down-filter ------------------------------------------
        float cb[2][4], cr[2][4];
        uint32_t yCode[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
        {
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                const int j = r * 12 + i * 3;
                float yf;
                ForwardPixel(p.matrix, code[j], code[j + 1], code[j + 2], yf, cb[r][i], cr[r][i]);
                yCode[r][i] = QuantiseLuma(yf, p.maxCode);
            }
        }
        {
            uint8_t* yRow = p.planeY + static_cast<int64_t>(y0) * p.strideY + static_cast<int64_t>(x0) * 2;
            const uint2 packed0 = make_uint2(yCode[0][0] | (yCode[0][1] << 16), yCode[0][2] | (yCode[0][3] << 16));
            __stcs(reinterpret_cast<uint2*>(yRow), packed0);
            if (secondRow)
            {
                const uint2 packed1 = make_uint2(yCode[1][0] | (yCode[1][1] << 16), yCode[1][2] | (yCode[1][3] << 16));
                __stcs(reinterpret_cast<uint2*>(yRow + p.strideY), packed1);
            }
        }
        if (XS == 1 && YS == 1)
        {
            uint32_t cbCode[2], crCode[2];
#pragma unroll
            for (int s = 0; s < 2; ++s)
            {
                float cbv, crv;
                if (p.topLeft)
                {
                    cbv = cb[0][2 * s];
                    crv = cr[0][2 * s];
                }
                else
                {
                    cbv = ((cb[0][2 * s] + cb[0][2 * s + 1]) + (cb[1][2 * s] + cb[1][2 * s + 1])) * 0.25f;
                    crv = ((cr[0][2 * s] + cr[0][2 * s + 1]) + (cr[1][2 * s] + cr[1][2 * s + 1])) * 0.25f;
                }
                cbCode[s] = QuantiseChroma(cbv, p.chromaOffset, p.maxCode);
                crCode[s] = QuantiseChroma(crv, p.chromaOffset, p.maxCode);
            }
            const int64_t chromaOffsetBytes = static_cast<int64_t>(tileRow) * p.strideCb + static_cast<int64_t>(x0 >> 1) * 2;
            __stcs(reinterpret_cast<uint32_t*>(p.planeCb + chromaOffsetBytes), cbCode[0] | (cbCode[1] << 16));
            __stcs(reinterpret_cast<uint32_t*>(p.planeCr + static_cast<int64_t>(tileRow) * p.strideCr + static_cast<int64_t>(x0 >> 1) * 2),
                   crCode[0] | (crCode[1] << 16));
        }
        else if (XS == 1)
        {
#pragma unroll
            for (int r = 0; r < 2; ++r)
            {
                if (r == 1 && !secondRow) break;
                uint32_t cbCode[2], crCode[2];
#pragma unroll
                for (int s = 0; s < 2; ++s)
                {
                    const float cbv = p.topLeft ? cb[r][2 * s] : (cb[r][2 * s] + cb[r][2 * s + 1]) * 0.5f;
                    const float crv = p.topLeft ? cr[r][2 * s] : (cr[r][2 * s] + cr[r][2 * s + 1]) * 0.5f;
                    cbCode[s] = QuantiseChroma(cbv, p.chromaOffset, p.maxCode);
                    crCode[s] = QuantiseChroma(crv, p.chromaOffset, p.maxCode);
                }
                __stcs(reinterpret_cast<uint32_t*>(p.planeCb + static_cast<int64_t>(y0 + r) * p.strideCb + static_cast<int64_t>(x0 >> 1) * 2),
                       cbCode[0] | (cbCode[1] << 16));
                __stcs(reinterpret_cast<uint32_t*>(p.planeCr + static_cast<int64_t>(y0 + r) * p.strideCr + static_cast<int64_t>(x0 >> 1) * 2),
                       crCode[0] | (crCode[1] << 16));
            }
        }
        else
        {
#pragma unroll
            for (int r = 0; r < 2; ++r)
            {
                if (r == 1 && !secondRow) break;
                uint32_t cbCode[4], crCode[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    cbCode[i] = QuantiseChroma(cb[r][i], p.chromaOffset, p.maxCode);
                    crCode[i] = QuantiseChroma(cr[r][i], p.chromaOffset, p.maxCode);
                }
                __stcs(reinterpret_cast<uint2*>(p.planeCb + static_cast<int64_t>(y0 + r) * p.strideCb + static_cast<int64_t>(x0) * 2),
                       make_uint2(cbCode[0] | (cbCode[1] << 16), cbCode[2] | (cbCode[3] << 16)));
                __stcs(reinterpret_cast<uint2*>(p.planeCr + static_cast<int64_t>(y0 + r) * p.strideCr + static_cast<int64_t>(x0) * 2),
                       make_uint2(crCode[0] | (crCode[1] << 16), crCode[2] | (crCode[3] << 16)));
            }
        }
    }
}

size_t FastEncodeSharedBytes(int bucketCount)
{
    return 768 + 2048 + kFastWarps * kQueueCapacity * sizeof(QueueEntry) + kFastWarps * 32 * kValuesPerLane * sizeof(uint16_t) +
           static_cast<size_t>(bucketCount) * sizeof(uint32_t);
}

template <int CURVE, int XS, int YS>
cudaError_t LaunchFastEncodeKernel(const FastEncodeParams& fp, int smCount, cudaStream_t stream)
{
    const size_t shared = FastEncodeSharedBytes(CURVE == kCurveClip ? 0 : fp.table.bucketCount);
    static bool configured = false; // per instantiation
    if (!configured)
    {
        const cudaError_t e = cudaFuncSetAttribute(EncodeRgbF32PlanarKernel<CURVE, XS, YS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        if (e != cudaSuccess)
        {
            return e;
        }
        configured = true;
    }
    if (shared > 110 * 1024)
    {
        return cudaErrorInvalidValue;
    }
    const long long tiles = static_cast<long long>((fp.width + kTilePixels - 1) / kTilePixels) * ((fp.rowCount + 1) / 2);
    long long blocks = (tiles + kFastWarps - 1) / kFastWarps;
    const long long resident = static_cast<long long>(smCount) * 2;
    if (blocks > resident)
    {
        blocks = resident;
    }
    EncodeRgbF32PlanarKernel<CURVE, XS, YS><<<static_cast<unsigned>(blocks), kFastThreads, shared, stream>>>(fp);
    return cudaGetLastError();
}

template <int CURVE>
cudaError_t DispatchChroma(const FastEncodeParams& fp, int xs, int ys, int smCount, cudaStream_t stream)
{
    if (xs == 1 && ys == 1) return LaunchFastEncodeKernel<CURVE, 1, 1>(fp, smCount, stream);
    if (xs == 1) return LaunchFastEncodeKernel<CURVE, 1, 0>(fp, smCount, stream);
    return LaunchFastEncodeKernel<CURVE, 0, 0>(fp, smCount, stream);
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
    if (hostDepth != 32 || !p.planar || p.channels != 3 || p.hasAlpha || p.imageDepth <= 8 || p.matrix.identity)
    {
        return 0;
    }
    int curve;
    if (p.transfer == AVIFGPU_TRANSFER_PQ) curve = kCurveLinearToPQ;
    else if (p.transfer == AVIFGPU_TRANSFER_SMPTE428) curve = kCurveLinearToSMPTE428;
    else curve = kCurveClip;
    if (curve != kCurveClip && (p.curveTable == nullptr || p.curveTable->buckets == nullptr))
    {
        return 0; // no verified table for this curve: the generic exact kernel serves it
    }
    if (!Aligned(p.rows, p.rowStride, 16) || !Aligned(p.plane[0], p.planeStride[0], 8) ||
        !Aligned(p.plane[1], p.planeStride[1], p.xs ? 4 : 8) || !Aligned(p.plane[2], p.planeStride[2], p.xs ? 4 : 8))
    {
        return 0;
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

    const int smCount = p.smCount > 0 ? p.smCount : 148;
    cudaError_t e;
    if (curve == kCurveLinearToPQ) e = DispatchChroma<kCurveLinearToPQ>(fp, p.xs, p.ys, smCount, stream);
    else if (curve == kCurveLinearToSMPTE428) e = DispatchChroma<kCurveLinearToSMPTE428>(fp, p.xs, p.ys, smCount, stream);
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
        strip.rows = static_cast<const uint8_t*>(p.rows) + static_cast<int64_t>(width4) * 12;
        strip.width = p.width - width4;
        strip.plane[0] = static_cast<uint8_t*>(p.plane[0]) + static_cast<int64_t>(width4) * 2;
        strip.plane[1] = static_cast<uint8_t*>(p.plane[1]) + static_cast<int64_t>(width4 >> p.xs) * 2;
        strip.plane[2] = static_cast<uint8_t*>(p.plane[2]) + static_cast<int64_t>(width4 >> p.xs) * 2;
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
        const int n = LaunchEncodeGeneric(strip, hostDepth, streamHandle);
        if (n < 0) return n;
        launched += n;
    }
    return launched;
}

int LaunchDecodeFast(const DecodeParams&, void*) { return 0; }

} // namespace avifgpu
