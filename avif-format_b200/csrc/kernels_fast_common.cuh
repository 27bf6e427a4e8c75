// kernels_fast_common.cuh -- pieces shared by the tuned float-RGB(A) encode kernels (kernels_fast.cu: no curve;
// kernels_fast_flat.cu: step tables + band bitmap + bulk-copy staging; kernels_fast_rgba.cu: the same look-up for RGBA).
#ifndef AVIF_KERNELS_FAST_COMMON_CUH
#define AVIF_KERNELS_FAST_COMMON_CUH

#include "kernel_params.h"
#include "curve_lookup.cuh"
#include "packed_f32x2.cuh"

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
    int32_t preferWideEntries; // AVIFGPU_WIDE_TABLE_ENTRIES=1 in the environment: the 64-bit flat table even where the compact one applies (A/B measurements)
    CurveTableView table;
};

// The caller (LaunchEncodeFast) has checked ForwardMatrixStaysInRange(): luma needs no upper clamp, chroma only the
// H.273 clip of 2^depth (a saturated red / blue) to 2^depth - 1, done on two packed codes at once.
// A lane's 2 rows x 4 pixels of R'G'B' codes (as floats, row-major, interleaved) -> Y / Cb / Cr codes in the planes:
// forward matrix, luma quantisation, chroma down-filter (the lane owns whole chroma sites, no cross-lane traffic).
// yRow / cbRow / crRow point at the lane's first sample of the tile's first row in each plane.
//
// The float arithmetic runs two pixels per instruction (packed_f32x2.cuh): pixels (0, 2) and (1, 3) of a row share a
// register pair, so the two chroma sites of a 4:2:0 / 4:2:2 lane are the two halves of one packed value.  The operation
// sequence per pixel is pixel_math.cuh's ForwardPixelFloat, rounding for rounding; see packed_f32x2.cuh for which
// operations may be packed (a product's sum is always a scalar add).
template <int XS, int YS>
__device__ __forceinline__ void StoreTile(const FastEncodeParams& p, const float (&codeF)[kValuesPerLane], uint8_t* yRow, uint8_t* cbRow, uint8_t* crRow,
                                          bool secondRow)
{
    using namespace avifx2;
    const F32x2 half2 = Splat(0.5f);
    const F32x2 offset2 = Splat(p.chromaOffset);
    const uint32_t maxPair = static_cast<uint32_t>(p.maxCode) * 0x00010001u;
    // cb2[r][h] / cr2[r][h]: float chroma of pixels (h, h + 2) of row r
    F32x2 cb2[2][2], cr2[2][2];
    uint32_t yWord[2][2]; // two packed 16-bit codes each: pixels (0, 1) and (2, 3)
#pragma unroll
    for (int r = 0; r < 2; ++r)
    {
        uint32_t yLow[2], yHigh[2];
#pragma unroll
        for (int h = 0; h < 2; ++h)
        {
            const int j0 = r * 12 + h * 3;
            const int j1 = j0 + 6;
            if (p.matrix.identity)
            {
                // lossless GBR: Y = G, Cb = B, Cr = R
                cb2[r][h] = Pack(codeF[j0 + 2], codeF[j1 + 2]);
                cr2[r][h] = Pack(codeF[j0], codeF[j1]);
                yLow[h] = __float2uint_rz(codeF[j0 + 1] + 0.5f);
                yHigh[h] = __float2uint_rz(codeF[j1 + 1] + 0.5f);
                continue;
            }
            const F32x2 red = Pack(codeF[j0], codeF[j1]);
            const F32x2 green = Pack(codeF[j0 + 1], codeF[j1 + 1]);
            const F32x2 blue = Pack(codeF[j0 + 2], codeF[j1 + 2]);
            float r0, r1, g0, g1, b0, b1;
            Unpack(Mul2(red, Splat(p.matrix.kr)), r0, r1);
            Unpack(Mul2(green, Splat(p.matrix.kg)), g0, g1);
            Unpack(Mul2(blue, Splat(p.matrix.kb)), b0, b1);
            const F32x2 luma = Pack(__fadd_rn(__fadd_rn(r0, g0), b0), __fadd_rn(__fadd_rn(r1, g1), b1)); // (kr R + kg G) + kb B
            cb2[r][h] = Mul2(Sub2(blue, luma), Splat(p.matrix.cbScale));
            cr2[r][h] = Mul2(Sub2(red, luma), Splat(p.matrix.crScale));
            float q0, q1;
            Unpack(Add2(luma, half2), q0, q1);
            yLow[h] = __float2uint_rz(q0);
            yHigh[h] = __float2uint_rz(q1);
        }
        yWord[r][0] = yLow[0] | (yLow[1] << 16);   // pixels 0, 1
        yWord[r][1] = yHigh[0] | (yHigh[1] << 16); // pixels 2, 3
    }
    __stcs(reinterpret_cast<uint2*>(yRow), make_uint2(yWord[0][0], yWord[0][1]));
    if (secondRow)
    {
        __stcs(reinterpret_cast<uint2*>(yRow + p.strideY), make_uint2(yWord[1][0], yWord[1][1]));
    }

    // (chroma + offset) + 0.5 -> code, both halves; `biased` is a scalar-add result or an exact scaling, never a product
    const auto quantisePair = [&](F32x2 biased) -> uint32_t
    {
        float q0, q1;
        Unpack(Add2(biased, half2), q0, q1);
        return __vminu2(__float2uint_rz(q0) | (__float2uint_rz(q1) << 16), maxPair);
    };
    // the two halves of a product pair added to those of another, as scalars (see the header's rule)
    const auto addHalves = [](F32x2 a, F32x2 b) -> F32x2
    {
        float a0, a1, b0, b1;
        Unpack(a, a0, a1);
        Unpack(b, b0, b1);
        return Pack(__fadd_rn(a0, b0), __fadd_rn(a1, b1));
    };
    const auto addOffset = [&](F32x2 product) -> F32x2
    {
        float c0, c1;
        Unpack(product, c0, c1);
        return Pack(__fadd_rn(c0, p.chromaOffset), __fadd_rn(c1, p.chromaOffset));
    };
    if (XS == 1 && YS == 1)
    {
        uint32_t cbWord, crWord;
        if (p.topLeft)
        {
            cbWord = quantisePair(addOffset(cb2[0][0]));
            crWord = quantisePair(addOffset(cr2[0][0]));
        }
        else
        {
            // ((c00 + c01) + (c10 + c11)) * 0.25f + offset; the scaling by 2^-2 is exact, so the fused form is the same number
            const F32x2 quarter2 = Splat(0.25f);
            const F32x2 cbSum = Add2(addHalves(cb2[0][0], cb2[0][1]), addHalves(cb2[1][0], cb2[1][1]));
            const F32x2 crSum = Add2(addHalves(cr2[0][0], cr2[0][1]), addHalves(cr2[1][0], cr2[1][1]));
            cbWord = quantisePair(Fma2(cbSum, quarter2, offset2));
            crWord = quantisePair(Fma2(crSum, quarter2, offset2));
        }
        __stcs(reinterpret_cast<uint32_t*>(cbRow), cbWord);
        __stcs(reinterpret_cast<uint32_t*>(crRow), crWord);
    }
    else if (XS == 1)
    {
#pragma unroll
        for (int r = 0; r < 2; ++r)
        {
            if (r == 1 && !secondRow) break;
            uint32_t cbWord, crWord;
            if (p.topLeft)
            {
                cbWord = quantisePair(addOffset(cb2[r][0]));
                crWord = quantisePair(addOffset(cr2[r][0]));
            }
            else
            {
                cbWord = quantisePair(Fma2(addHalves(cb2[r][0], cb2[r][1]), half2, offset2)); // (c0 + c1) * 0.5f + offset, exact scaling
                crWord = quantisePair(Fma2(addHalves(cr2[r][0], cr2[r][1]), half2, offset2));
            }
            __stcs(reinterpret_cast<uint32_t*>(cbRow + r * p.strideCb), cbWord);
            __stcs(reinterpret_cast<uint32_t*>(crRow + r * p.strideCr), crWord);
        }
    }
    else
    {
#pragma unroll
        for (int r = 0; r < 2; ++r)
        {
            if (r == 1 && !secondRow) break;
            // pairs hold pixels (0, 2) and (1, 3); the stores want (0, 1) and (2, 3)
            const uint32_t cbEven = quantisePair(addOffset(cb2[r][0])), cbOdd = quantisePair(addOffset(cb2[r][1]));
            const uint32_t crEven = quantisePair(addOffset(cr2[r][0])), crOdd = quantisePair(addOffset(cr2[r][1]));
            __stcs(reinterpret_cast<uint2*>(cbRow + r * p.strideCb), make_uint2(__byte_perm(cbEven, cbOdd, 0x5410), __byte_perm(cbEven, cbOdd, 0x7632)));
            __stcs(reinterpret_cast<uint2*>(crRow + r * p.strideCr), make_uint2(__byte_perm(crEven, crOdd, 0x5410), __byte_perm(crEven, crOdd, 0x7632)));
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
cudaError_t LaunchFastEncodeFlatInterleaved(const fastenc::FastEncodeParams& fp, int curve, int smCount, cudaStream_t stream);

} // namespace avifgpu

#endif
