// pixel_math.cuh -- the per-pixel arithmetic of the avif-format colour path, as __host__ __device__ functions.
//
// Each function states the reference lines whose arithmetic it reproduces (paths relative to the reference's
// src/common/).  The float expressions are written in the reference's association and must be compiled
// without contraction (nvcc -fmad=false; g++ -ffp-contract=off for the host-side checks): integer outputs are
// trunc(0.5f + c * max) or trunc(clamp(c * max)) of a float chain, so a single fused multiply-add flips codes.
// Division and square root are IEEE (nvcc default -prec-div=true -prec-sqrt=true), denormals are kept
// (-ftz=false).  libm calls go through device_math.cuh (glibc-identical powf / expf / logf).
#ifndef AVIF_PIXEL_MATH_CUH
#define AVIF_PIXEL_MATH_CUH

#include "device_math.cuh"

namespace avifpix
{

using avifmath::LibmTables;

// std::min(a, b) and std::clamp(v, lo, hi) exactly as the reference uses them (NaN behaviour included).
AVIF_HD float MinF(float a, float b) { return (b < a) ? b : a; }
AVIF_HD float MaxF(float a, float b) { return (a < b) ? b : a; }
AVIF_HD float ClampF(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

// ---- transfer functions: ColorTransfer.cpp ----------------------------------------------------------------

struct PqConstants
{
    // ColorTransfer.cpp:73-77 / 100-104 (constexpr float arithmetic, evaluated here by the host compiler in
    // binary32 exactly as MSVC / gcc evaluate the reference's constexpr initialisers).
    static constexpr float m1 = 2610.0f / 16384.0f;
    static constexpr float m2 = 2523.0f / 4096.0f * 128.0f;
    static constexpr float c1 = 3424.0f / 4096.0f;
    static constexpr float c2 = 2413.0f / 4096.0f * 32.0f;
    static constexpr float c3 = 2392.0f / 4096.0f * 32.0f;
    static constexpr float inv_m2 = 1.0f / m2;
    static constexpr float inv_m1 = 1.0f / m1;
};

// ColorTransfer.cpp:69-92.  luminanceMultiplier = imageMaxLuminanceLevel / 10000.0f (float division, hoisted:
// it does not depend on the pixel).
AVIF_HD float LinearToPQ(float value, float luminanceMultiplier, const LibmTables& t)
{
    if (value < 0.0f)
    {
        return 0.0f;
    }
    const float x = avifmath::Powf(value * luminanceMultiplier, PqConstants::m1, t);
    const float pq = avifmath::Powf((PqConstants::c1 + PqConstants::c2 * x) / (1.0f + PqConstants::c3 * x), PqConstants::m2, t);
    return pq;
}

// ColorTransfer.cpp:94-117.  luminanceMultiplier = 10000.0f / imageMaxLuminanceLevel.
AVIF_HD float PQToLinear(float value, float luminanceMultiplier, const LibmTables& t)
{
    if (value < 0.0f)
    {
        return 0.0f;
    }
    // 1 / m2 = 0.0127: no exponent screens (PowfModerateExponent); a -0.0 or NaN base takes the full function inside it.
    // The second base can be negative (value above ~2: the denominator changes sign) and its exponent is 6.3: the full powf.
    const float x = avifmath::PowfModerateExponent(value, PqConstants::inv_m2, t);
    const float normalizedLinear =
        avifmath::Powf(MaxF(x - PqConstants::c1, 0.0f) / (PqConstants::c2 - PqConstants::c3 * x), PqConstants::inv_m1, t);
    return normalizedLinear * luminanceMultiplier;
}

// ColorTransfer.cpp:119-127
AVIF_HD float LinearToSMPTE428(float value, const LibmTables& t)
{
    if (value < 0.0f)
    {
        return 0.0f;
    }
    return avifmath::Powf(value * 48.0f / 52.37f, 1.0f / 2.6f, t);
}

// ColorTransfer.cpp:129-139
AVIF_HD float SMPTE428ToLinear(float value, const LibmTables& t)
{
    if (value < 0.0f)
    {
        return 0.0f;
    }
    return avifmath::Powf(value, 2.6f, t) * (52.37f / 48.0f);
}

// ColorTransfer.cpp:141-164 (no caller in the reference; provided for the HLG-save extension)
AVIF_HD float LinearToHLG(float value, const LibmTables& t)
{
    constexpr float a = 0.17883277f;
    constexpr float b = 0.28466892f;
    constexpr float c = 0.55991073f;
    if (value < 0.0f)
    {
        return 0.0f;
    }
    if (value > (1.0f / 12.0f))
    {
        value = a * avifmath::Logf(value * 12.0f - b, t) + c;
    }
    else
    {
        value = sqrtf(value * 3.0f);
    }
    return value;
}

// ColorTransfer.cpp:166-190
AVIF_HD float HLGToLinear(float value, const LibmTables& t)
{
    constexpr float a = 0.17883277f;
    constexpr float b = 0.28466892f;
    constexpr float c = 0.55991073f;
    if (value < 0.0f)
    {
        return 0.0f;
    }
    if (value > 0.5f)
    {
        value = (avifmath::Expf((value - c) / a, t) + b) / 12.0f;
    }
    else
    {
        value = (value * value) * (1.0f / 3.0f);
    }
    return value;
}

#if defined(__CUDACC__)
// x / d for a constant d through its rounded reciprocal and one residual correction (Markstein): three
// full-rate instructions instead of the ~14 of an IEEE division.  The result equals the correctly rounded
// quotient for almost every (x, d); this library only uses it where a verification kernel has compared it with
// the IEEE division for EVERY numerator the call site can produce (kernels_fast_decode.cu, VerifyHlgDivisions),
// so "almost" never enters: a site that failed verification keeps the true division.
__device__ __forceinline__ float DivideByConstant(float x, float d, float reciprocal)
{
    const float q = __fmul_rn(x, reciprocal);
    const float r = __fmaf_rn(-q, d, x);
    return __fmaf_rn(r, reciprocal, q);
}

// HLGToLinear (ColorTransfer.cpp:166-190) for value in [0, 1] (the decoders clamp before calling it), with the two
// constant divisions -- (value - c) / a over the 2^23 possible numerators and (e + b) / 12 over [1, 16) --
// replaced by DivideByConstant when `verifiedDivisions` (a compile-time choice) is set.  The exp argument is within (-0.34, 2.47), so the
// overflow / underflow screening of expf cannot trigger and is skipped.
template <bool verifiedDivisions>
__device__ __forceinline__ float HLGToLinearUnit(float value, const LibmTables& t)
{
    constexpr float a = 0.17883277f;
    constexpr float b = 0.28466892f;
    constexpr float c = 0.55991073f;
    // Both branches are evaluated and one is selected: on typical data half the lanes of a warp take each side, so
    // a real branch would execute both anyway, plus the divergence bookkeeping.  (For value <= 0.5 the exponential's
    // argument is in [-3.14, -0.33]: harmless, and its result is discarded.)
    const float numerator = value - c;
    const float argument = verifiedDivisions ? DivideByConstant(numerator, a, 1.0f / a) : numerator / a;
    const float e = avifmath::ExpfNoScreen(argument, t) + b;
    const float high = verifiedDivisions ? DivideByConstant(e, 12.0f, 1.0f / 12.0f) : e / 12.0f;
    const float low = (value * value) * (1.0f / 3.0f);
    return value > 0.5f ? high : low;
}
#endif

// ColorTransfer.cpp:192-205.  gammaMinusOne = displayGamma - 1.0f (float subtraction, hoisted).
// kLumaNotNegative: the caller knows r, g, b >= +0 (the decoders clamp to [0, 1] before the inverse OETF) and the
// luma coefficients are positive, so the luma's sign bit is clear and powf needs no negative-base handling.
template <bool kLumaNotNegative = false>
AVIF_HD void ApplyHLGOOTF(float& r, float& g, float& b, float lumaR, float lumaG, float lumaB, float gammaMinusOne,
                          float nominalPeakBrightness, const LibmTables& t)
{
    const float luma = (r * lumaR) + (g * lumaG) + (b * lumaB);
    const float factor = nominalPeakBrightness * avifmath::PowfImpl<kLumaNotNegative>(luma, gammaMinusOne, t);
    r *= factor;
    g *= factor;
    b *= factor;
}

// ColorTransfer.cpp:207-220 (the reference never calls it; kept at parity as a primitive, and the seam an HLG save path
// would use).
AVIF_HD void ApplyInverseHLGOOTF(float& r, float& g, float& b, float lumaR, float lumaG, float lumaB, float displayGamma,
                                 float nominalPeakBrightness, const LibmTables& t)
{
    const float luma = (r * lumaR) + (g * lumaG) + (b * lumaB);
    const float factor = avifmath::Powf(luma / nominalPeakBrightness, (displayGamma - 1.0f) / displayGamma, t) / nominalPeakBrightness;
    r *= factor;
    g *= factor;
    b *= factor;
}

// ---- alpha: PremultipliedAlpha.cpp ------------------------------------------------------------------------

// PremultipliedAlpha.cpp:49-52
AVIF_HD float PremultiplyColor(float color, float alpha, float maxValue) { return color * alpha / maxValue; }

// PremultipliedAlpha.cpp:54-70 (uint8 and uint16 overloads share this body; maxValue 255 or 2^depth-1)
AVIF_HD uint32_t PremultiplyCode(uint32_t color, uint32_t alpha, float maxValueFloat)
{
    const float value = PremultiplyColor(static_cast<float>(color), static_cast<float>(alpha), maxValueFloat);
    return static_cast<uint32_t>(MinF(roundf(value), maxValueFloat));
}

// PremultipliedAlpha.cpp:72-75
AVIF_HD float UnpremultiplyColor(float color, float alpha, float maxValue) { return MinF(color * maxValue / alpha, maxValue); }

// PremultipliedAlpha.cpp:77-93
AVIF_HD uint32_t UnpremultiplyCode(uint32_t color, uint32_t alpha, float maxValueFloat)
{
    const float value = UnpremultiplyColor(static_cast<float>(color), static_cast<float>(alpha), maxValueFloat);
    return static_cast<uint32_t>(MinF(roundf(value), maxValueFloat));
}

// The caller-side guard every integer premultiply site wraps around PremultiplyColor
// (WriteHeifImage.cpp:238-251, 404-417, 700-718, 760-778, 877-895, 947-965).
AVIF_HD uint32_t PremultiplyCodeGuarded(uint32_t color, uint32_t alpha, uint32_t maxValue)
{
    if (alpha < maxValue)
    {
        if (alpha == 0)
        {
            return 0;
        }
        return PremultiplyCode(color, alpha, static_cast<float>(maxValue));
    }
    return color;
}

// ---- encode side: WriteHeifImage.cpp ----------------------------------------------------------------------

// WriteHeifImage.cpp:87-166: lut[i] = clamp((int)((i / fromMax) * toMax + 0.5f), 0, toMax) evaluated directly
// (fromMax = 255.0f or 32768.0f).  For 16-bit hosts a sample above 32768 would index past the reference's
// table (undefined there); the same formula is DEFINED to apply, the clamp then yields toMax.
AVIF_HD uint32_t DepthLutEntry(uint32_t i, float fromMax, uint32_t toMax)
{
    int value = static_cast<int>(((static_cast<float>(i) / fromMax) * static_cast<float>(toMax)) + 0.5f);
    if (value < 0)
    {
        value = 0;
    }
    else if (value > static_cast<int>(toMax))
    {
        value = static_cast<int>(toMax);
    }
    return static_cast<uint32_t>(value);
}

// WriteHeifImage.cpp:590, 618, 1093-1096, 1128-1130: static_cast<uint16_t>(std::clamp(v * max, 0.0f, max)).
// NaN survives std::clamp and the cast is undefined in the reference (0 on x86); DEFINED here as 0.
AVIF_HD uint32_t FloatToCode(float v, float maxValue)
{
    const float scaled = ClampF(v * maxValue, 0.0f, maxValue);
    if (scaled != scaled)
    {
        return 0;
    }
    return static_cast<uint32_t>(scaled);
}

// Forward matrix, this project's definition of the stage the reference leaves to libheif (DESIGN.md "Forward
// matrix"): H.273's full-range equations on integer codes (the algebraic inverse of YuvDecode.cpp:555-557 up to the
// chroma zero): Y = (kr R + kg G) + kb B, Cb = (B - Y) * 0.5f/(1-kb), Cr = (R - Y) * 0.5f/(1-kr); Ycode = Clip(Round(Y)),
// Ccode = Clip(Round(C + 2^(depth-1))) with Round(v) = (int)(v + 0.5f).
struct ForwardMatrix
{
    float kr, kg, kb;
    float cbScale; // 0.5f / (1-kb)
    float crScale; // 0.5f / (1-kr)
    int identity;    // matrix_coefficients == 0 (GBR)
};

// (float)code.  The conversion instruction issues on the quarter-rate pipe, which this path otherwise leaves idle;
// the integer pipe is the busy one, so the conversion is cheaper here than the usual 2^23 bit trick.
AVIF_HD float CodeToFloat(uint32_t code) { return static_cast<float>(code); }

AVIF_HD void ForwardPixelFloat(const ForwardMatrix& m, float r, float g, float b, float& y, float& cb, float& cr)
{
    if (m.identity)
    {
        y = g;
        cb = b;
        cr = r;
        return;
    }
    y = ((m.kr * r) + (m.kg * g)) + (m.kb * b);
    cb = (b - y) * m.cbScale;
    cr = (r - y) * m.crScale;
}

AVIF_HD void ForwardPixel(const ForwardMatrix& m, uint32_t rc, uint32_t gc, uint32_t bc, float& y, float& cb, float& cr)
{
    ForwardPixelFloat(m, CodeToFloat(rc), CodeToFloat(gc), CodeToFloat(bc), y, cb, cr);
}

// clamp((int)(v), 0, maxCode).  On the device the float -> unsigned conversion saturates negatives (and NaN) to
// 0 -- the same value the signed conversion + lower clamp gives -- so one min() finishes the job.
AVIF_HD uint32_t TruncateToCode(float v, int maxCode)
{
#if defined(__CUDA_ARCH__)
    return min(__float2uint_rz(v), static_cast<uint32_t>(maxCode));
#else
    int i = static_cast<int>(v);
    i = (i < 0) ? 0 : ((i > maxCode) ? maxCode : i);
    return static_cast<uint32_t>(i);
#endif
}

AVIF_HD uint32_t QuantiseLuma(float y, int maxCode) { return TruncateToCode(y + 0.5f, maxCode); }

// The luma quantiser without the upper clamp, for callers that have checked ForwardMatrixStaysInRange(): the
// float -> unsigned conversion already sends negatives (and NaN) to 0, and the matrix cannot reach maxCode + 1.
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t QuantiseLumaInRange(float y) { return __float2uint_rz(y + 0.5f); }
#endif

// True when, for R'G'B' codes in [0, maxCode], Y + 0.5 stays below maxCode + 1 with a margin (0.25) far above the float
// rounding of the five-operation matrix (< 0.01 at 16 bits), and C + chromaOffset + 0.5 stays below maxCode + 2: with
// the H.273 offset 2^(depth-1) a saturated red / blue lands exactly on 2^depth = maxCode + 1, which the tuned kernels
// clip with one packed min; nothing may reach further.
inline bool ForwardMatrixStaysInRange(const ForwardMatrix& m, float chromaOffset, int maxCode)
{
    if (m.identity)
    {
        return true; // Y = G, Cb = B, Cr = R: codes pass through
    }
    const double max = static_cast<double>(maxCode);
    const double lumaTop = (static_cast<double>(m.kr) + m.kg + m.kb) * max + 0.5;
    const double cbTop = max * (1.0 - m.kb) * m.cbScale + chromaOffset + 0.5;
    const double crTop = max * (1.0 - m.kr) * m.crScale + chromaOffset + 0.5;
    return m.kr >= 0 && m.kg >= 0 && m.kb >= 0 && m.kb < 1 && m.kr < 1 && lumaTop < max + 0.75 && cbTop < max + 1.75 && crTop < max + 1.75;
}

// chromaOffset = 2^(depth-1) as a float (H.273 full range), or 0 for the identity matrix.
AVIF_HD uint32_t QuantiseChroma(float c, float chromaOffset, int maxCode) { return TruncateToCode((c + chromaOffset) + 0.5f, maxCode); }

// ---- decode side: YuvLookupTables.cpp / YuvDecode.cpp -----------------------------------------------------

// YuvLookupTables.cpp:64-66 with the reference's int arithmetic (the 16-bit case overflows int there; both the
// oracle (-fwrapv) and this code use two's-complement wrap-around so they agree).
AVIF_HD int LimitedToFull(int v, int lo, int hi, int full)
{
    const uint32_t product = static_cast<uint32_t>(v - lo) * static_cast<uint32_t>(full) + static_cast<uint32_t>((hi - lo) / 2);
    v = static_cast<int32_t>(product) / (hi - lo);
    return (v > full) ? full : ((v < 0) ? 0 : v);
}

struct RangeParams
{
    int fullRange;
    int yLo, yHi;   // limited-range luma foot / head   (YuvLookupTables.cpp:72-83)
    int uvLo, uvHi; // limited-range chroma foot / head (YuvLookupTables.cpp:93-104)
    int maxChannel; // (1 << depth) - 1
    float maxChannelFloat;
    int identityMatrix;
};

// unormFloatTableY[i], YuvLookupTables.cpp:157-171
AVIF_HD float UnormToFloatY(uint32_t code, const RangeParams& p)
{
    int v = static_cast<int>(code);
    if (!p.fullRange)
    {
        v = LimitedToFull(v, p.yLo, p.yHi, p.maxChannel);
    }
    return static_cast<float>(v) / p.maxChannelFloat;
}

// unormFloatTableUV[i], YuvLookupTables.cpp:173-184
AVIF_HD float UnormToFloatUV(uint32_t code, const RangeParams& p)
{
    if (p.identityMatrix)
    {
        return UnormToFloatY(code, p);
    }
    int v = static_cast<int>(code);
    if (!p.fullRange)
    {
        v = LimitedToFull(v, p.uvLo, p.uvHi, p.maxChannel);
    }
    return static_cast<float>(v) / p.maxChannelFloat - 0.5f;
}

// unormFloatTableAlpha[i], YuvLookupTables.cpp:186-189; also BuildUnormToFloatLookupTable, ReadHeifImage.cpp:402-415
AVIF_HD float UnormToFloatPlain(uint32_t code, float maxChannelFloat) { return static_cast<float>(code) / maxChannelFloat; }

struct InverseMatrix
{
    float kr, kg, kb;
};

// YuvDecode.cpp:306-312 (the same three lines appear in all six colour row decoders).
AVIF_HD void YuvToRgb(const InverseMatrix& m, float Y, float Cb, float Cr, float& R, float& G, float& B)
{
    R = Y + (2 * (1 - m.kr)) * Cr;
    B = Y + (2 * (1 - m.kb)) * Cb;
    G = Y - ((2 * ((m.kr * (1 - m.kr) * Cr) + (m.kb * (1 - m.kb) * Cb))) / m.kg);
    R = ClampF(R, 0.0f, 1.0f);
    G = ClampF(G, 0.0f, 1.0f);
    B = ClampF(B, 0.0f, 1.0f);
}

} // namespace avifpix

#endif // AVIF_PIXEL_MATH_CUH
