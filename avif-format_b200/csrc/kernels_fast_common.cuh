// kernels_fast_common.cuh -- pieces shared by the tuned float-RGB(A) encode kernels (kernels_fast.cu: no curve;
// kernels_fast_flat.cu: step tables + band bitmap + bulk-copy staging; kernels_fast_rgba.cu: the same look-up for RGBA).
#ifndef AVIF_KERNELS_FAST_COMMON_CUH
#define AVIF_KERNELS_FAST_COMMON_CUH

#include "kernel_params.h"
#include "curve_lookup.cuh"

#include <cuda_runtime.h>

namespace avifgpu
{
namespace fastenc
{

using namespace avifpix;

constexpr int kTilePixels = 128;     // per row
constexpr int kValuesPerLane = 24;   // 2 rows x 4 pixels x 3 channels
constexpr int kCurveClip = 2;        // no transfer curve: code = trunc(clamp(v * max))
constexpr int kSharedLibm = 768;

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
    uint8_t* planeA;  // RGBA hosts (kernels_fast_rgba.cu)
    int64_t strideA;
    int32_t premultiply;
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

// The caller (LaunchEncodeFast) has checked ForwardMatrixStaysInRange(), so the quantisers need no upper clamp.
// A lane's 2 rows x 4 pixels of R'G'B' codes (as floats, row-major, interleaved) -> Y / Cb / Cr codes in the planes:
// forward matrix, luma quantisation, chroma down-filter (the lane owns whole chroma sites, no cross-lane traffic).
// yRow / cbRow / crRow point at the lane's first sample of the tile's first row in each plane.
template <int XS, int YS>
__device__ __forceinline__ void StoreTile(const FastEncodeParams& p, const float (&codeF)[kValuesPerLane], uint8_t* yRow, uint8_t* cbRow, uint8_t* crRow,
                                          bool secondRow)
{
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
            ForwardPixelFloat(p.matrix, codeF[j], codeF[j + 1], codeF[j + 2], yf, cb[r][i], cr[r][i]);
            yCode[r][i] = QuantiseLumaInRange(yf);
        }
    }
    {
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
            cbCode[s] = QuantiseChromaInRange(cbv, p.chromaOffset);
            crCode[s] = QuantiseChromaInRange(crv, p.chromaOffset);
        }
        __stcs(reinterpret_cast<uint32_t*>(cbRow), cbCode[0] | (cbCode[1] << 16));
        __stcs(reinterpret_cast<uint32_t*>(crRow), crCode[0] | (crCode[1] << 16));
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
                cbCode[s] = QuantiseChromaInRange(cbv, p.chromaOffset);
                crCode[s] = QuantiseChromaInRange(crv, p.chromaOffset);
            }
            __stcs(reinterpret_cast<uint32_t*>(cbRow + r * p.strideCb), cbCode[0] | (cbCode[1] << 16));
            __stcs(reinterpret_cast<uint32_t*>(crRow + r * p.strideCr), crCode[0] | (crCode[1] << 16));
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
                cbCode[i] = QuantiseChromaInRange(cb[r][i], p.chromaOffset);
                crCode[i] = QuantiseChromaInRange(cr[r][i], p.chromaOffset);
            }
            __stcs(reinterpret_cast<uint2*>(cbRow + r * p.strideCb), make_uint2(cbCode[0] | (cbCode[1] << 16), cbCode[2] | (cbCode[3] << 16)));
            __stcs(reinterpret_cast<uint2*>(crRow + r * p.strideCr), make_uint2(crCode[0] | (crCode[1] << 16), crCode[2] | (crCode[3] << 16)));
        }
    }
}

} // namespace fastenc

// kernels_fast_rgba.cu
bool RgbaEncodeApplies(const fastenc::FastEncodeParams& fp);
cudaError_t LaunchFastEncodeRgba(const fastenc::FastEncodeParams& fp, int curve, int xs, int ys, int smCount, cudaStream_t stream);

// kernels_fast_flat.cu
bool FlatEncodeApplies(const fastenc::FastEncodeParams& fp);
cudaError_t LaunchFastEncodeFlat(const fastenc::FastEncodeParams& fp, int curve, int xs, int ys, int smCount, cudaStream_t stream);

} // namespace avifgpu

#endif
