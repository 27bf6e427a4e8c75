/*
 * oracle/avif_oracle.c -- TEST INFRASTRUCTURE ONLY (see avif_oracle.h for the parity status).
 *
 * CPU restatement of the avif-format colour-conversion path on plain buffers.  Every function cites the
 * reference lines it follows (paths relative to the reference's src/common/).  Floating point is IEEE
 * binary32 evaluated op by op in the reference's association; build with -ffp-contract=off (oracle/Makefile).
 * libm (powf / expf / logf / sqrtf / roundf) is whatever the host provides -- the reference's numerics are
 * defined by its libm too; avif_oracle_libm_version() is recorded with every result.
 */
#include "avif_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_error[256];

static int fail(int code, const char* message)
{
    snprintf(g_error, sizeof(g_error), "%s", message);
    return code;
}

const char* avif_oracle_last_error(void) { return g_error; }

const char* avif_oracle_libm_version(void)
{
#if defined(__GLIBC__)
    static char text[64];
    snprintf(text, sizeof(text), "glibc %d.%d", __GLIBC__, __GLIBC_MINOR__);
    return text;
#else
    return "unknown libm";
#endif
}

/* ---------------------------------------------------------------------------------------------------- */
/* Transfer functions: ColorTransfer.cpp                                                                  */
/* ---------------------------------------------------------------------------------------------------- */

/* ColorTransfer.cpp:27-29 */
static const float kPqMaxLuminance = 10000.0f;

/* ColorTransfer.cpp:69-92 */
static float linear_to_pq(float value, float image_max_luminance)
{
    const float m1 = 2610.0f / 16384.0f;
    const float m2 = 2523.0f / 4096.0f * 128.0f;
    const float c1 = 3424.0f / 4096.0f;
    const float c2 = 2413.0f / 4096.0f * 32.0f;
    const float c3 = 2392.0f / 4096.0f * 32.0f;

    if (value < 0.0f)
    {
        return 0.0f;
    }
    const float luminance_multiplier = image_max_luminance / kPqMaxLuminance;
    const float x = powf(value * luminance_multiplier, m1);
    const float pq = powf((c1 + c2 * x) / (1.0f + c3 * x), m2);
    return pq;
}

/* ColorTransfer.cpp:94-117 */
static float pq_to_linear(float value, float image_max_luminance)
{
    const float m1 = 2610.0f / 16384.0f;
    const float m2 = 2523.0f / 4096.0f * 128.0f;
    const float c1 = 3424.0f / 4096.0f;
    const float c2 = 2413.0f / 4096.0f * 32.0f;
    const float c3 = 2392.0f / 4096.0f * 32.0f;

    if (value < 0.0f)
    {
        return 0.0f;
    }
    const float x = powf(value, 1.0f / m2);
    const float difference = x - c1;
    /* std::max(x - c1, 0.0f): returns the first argument unless it is less than the second. */
    const float numerator = (difference < 0.0f) ? 0.0f : difference;
    const float normalized_linear = powf(numerator / (c2 - c3 * x), 1.0f / m1);
    const float luminance_multiplier = kPqMaxLuminance / image_max_luminance;
    return normalized_linear * luminance_multiplier;
}

/* ColorTransfer.cpp:119-127 */
static float linear_to_smpte428(float value)
{
    if (value < 0.0f)
    {
        return 0.0f;
    }
    return powf(value * 48.0f / 52.37f, 1.0f / 2.6f);
}

/* ColorTransfer.cpp:129-139 */
static float smpte428_to_linear(float value)
{
    if (value < 0.0f)
    {
        return 0.0f;
    }
    return powf(value, 2.6f) * (52.37f / 48.0f);
}

/* ColorTransfer.cpp:141-164 (no caller in the reference; kept for the HLG-save extension) */
static float linear_to_hlg(float value)
{
    const float a = 0.17883277f;
    const float b = 0.28466892f;
    const float c = 0.55991073f;
    if (value < 0.0f)
    {
        return 0.0f;
    }
    if (value > (1.0f / 12.0f))
    {
        value = a * logf(value * 12.0f - b) + c;
    }
    else
    {
        value = sqrtf(value * 3.0f);
    }
    return value;
}

/* ColorTransfer.cpp:166-190 */
static float hlg_to_linear(float value)
{
    const float a = 0.17883277f;
    const float b = 0.28466892f;
    const float c = 0.55991073f;
    if (value < 0.0f)
    {
        return 0.0f;
    }
    if (value > 0.5f)
    {
        value = (expf((value - c) / a) + b) / 12.0f;
    }
    else
    {
        value = (value * value) * (1.0f / 3.0f);
    }
    return value;
}

/* ColorTransfer.cpp:192-205 */
static void apply_hlg_ootf(float* rgb, const float luma_coefficients[3], float display_gamma, float peak)
{
    const float luma = (rgb[0] * luma_coefficients[0]) + (rgb[1] * luma_coefficients[1]) + (rgb[2] * luma_coefficients[2]);
    const float factor = peak * powf(luma, display_gamma - 1.0f);
    rgb[0] *= factor;
    rgb[1] *= factor;
    rgb[2] *= factor;
}

/* ColorTransfer.cpp:207-220 (no caller in the reference; kept at parity as a primitive) */
static void apply_inverse_hlg_ootf(float* rgb, const float luma_coefficients[3], float display_gamma, float peak)
{
    const float luma = (rgb[0] * luma_coefficients[0]) + (rgb[1] * luma_coefficients[1]) + (rgb[2] * luma_coefficients[2]);
    const float factor = powf(luma / peak, (display_gamma - 1.0f) / display_gamma) / peak;
    rgb[0] *= factor;
    rgb[1] *= factor;
    rgb[2] *= factor;
}

/* ColorTransfer.cpp:31-45 */
int avif_oracle_get_hlg_luma_coefficients(int32_t primaries, float* out)
{
    switch (primaries)
    {
    case 1: /* BT.709 */
        out[0] = 0.2126f; out[1] = 0.7152f; out[2] = 0.0722f;
        return AVIFGPU_OK;
    case 5: /* BT.470 System B/G */
    case 6: /* BT.601 */
        out[0] = 0.299f; out[1] = 0.587f; out[2] = 0.114f;
        return AVIFGPU_OK;
    case 9: /* BT.2020 / BT.2100 */
        out[0] = 0.2627f; out[1] = 0.6780f; out[2] = 0.0593f;
        return AVIFGPU_OK;
    default:
        return fail(AVIFGPU_ERR_UNSUPPORTED, "Unsupported color primaries for the HLG Luma Coefficients ");
    }
}

int avif_oracle_transfer_f32(int32_t function, float param, const float* in, float* out, size_t n)
{
    size_t i;
    switch (function)
    {
    case AVIFGPU_FN_LINEAR_TO_PQ: for (i = 0; i < n; ++i) out[i] = linear_to_pq(in[i], param); break;
    case AVIFGPU_FN_PQ_TO_LINEAR: for (i = 0; i < n; ++i) out[i] = pq_to_linear(in[i], param); break;
    case AVIFGPU_FN_LINEAR_TO_SMPTE428: for (i = 0; i < n; ++i) out[i] = linear_to_smpte428(in[i]); break;
    case AVIFGPU_FN_SMPTE428_TO_LINEAR: for (i = 0; i < n; ++i) out[i] = smpte428_to_linear(in[i]); break;
    case AVIFGPU_FN_HLG_TO_LINEAR: for (i = 0; i < n; ++i) out[i] = hlg_to_linear(in[i]); break;
    case AVIFGPU_FN_LINEAR_TO_HLG: for (i = 0; i < n; ++i) out[i] = linear_to_hlg(in[i]); break;
    case AVIFGPU_FN_POWF: for (i = 0; i < n; ++i) out[i] = powf(in[i], param); break;
    case AVIFGPU_FN_EXPF: for (i = 0; i < n; ++i) out[i] = expf(in[i]); break;
    case AVIFGPU_FN_LOGF: for (i = 0; i < n; ++i) out[i] = logf(in[i]); break;
    default: return fail(AVIFGPU_ERR_BAD_PARAM, "unknown function");
    }
    return AVIFGPU_OK;
}

int avif_oracle_hlg_ootf(float* rgb, size_t pixels, int32_t primaries, float display_gamma, float peak)
{
    float luma[3];
    size_t i;
    const int status = avif_oracle_get_hlg_luma_coefficients(primaries, luma);
    if (status != AVIFGPU_OK)
    {
        return status;
    }
    for (i = 0; i < pixels; ++i)
    {
        apply_hlg_ootf(rgb + 3 * i, luma, display_gamma, peak);
    }
    return AVIFGPU_OK;
}

int avif_oracle_hlg_inverse_ootf(float* rgb, size_t pixels, int32_t primaries, float display_gamma, float peak)
{
    float luma[3];
    size_t i;
    const int status = avif_oracle_get_hlg_luma_coefficients(primaries, luma);
    if (status != AVIFGPU_OK)
    {
        return status;
    }
    for (i = 0; i < pixels; ++i)
    {
        apply_inverse_hlg_ootf(rgb + 3 * i, luma, display_gamma, peak);
    }
    return AVIFGPU_OK;
}

/* ---------------------------------------------------------------------------------------------------- */
/* Alpha: PremultipliedAlpha.cpp                                                                          */
/* ---------------------------------------------------------------------------------------------------- */

static float min_f(float a, float b) { return (b < a) ? b : a; } /* std::min(a, b) */

/* std::clamp(v, lo, hi): (v < lo) ? lo : (hi < v) ? hi : v  -- NaN passes through. */
static float clamp_f(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

/* PremultipliedAlpha.cpp:49-52 */
float avif_oracle_premultiply_f32(float color, float alpha, float max_value) { return color * alpha / max_value; }

/* PremultipliedAlpha.cpp:54-61 */
uint8_t avif_oracle_premultiply_u8(uint8_t color, uint8_t alpha)
{
    const float value = avif_oracle_premultiply_f32((float)color, (float)alpha, 255.0f);
    return (uint8_t)min_f(roundf(value), 255.0f);
}

/* PremultipliedAlpha.cpp:63-70 */
uint16_t avif_oracle_premultiply_u16(uint16_t color, uint16_t alpha, uint16_t max_value)
{
    const float max_value_float = (float)max_value;
    const float value = avif_oracle_premultiply_f32((float)color, (float)alpha, max_value_float);
    return (uint16_t)min_f(roundf(value), max_value_float);
}

/* PremultipliedAlpha.cpp:72-75 */
float avif_oracle_unpremultiply_f32(float color, float alpha, float max_value)
{
    return min_f(color * max_value / alpha, max_value);
}

/* PremultipliedAlpha.cpp:77-84 */
uint8_t avif_oracle_unpremultiply_u8(uint8_t color, uint8_t alpha)
{
    const float value = avif_oracle_unpremultiply_f32((float)color, (float)alpha, 255.0f);
    return (uint8_t)min_f(roundf(value), 255.0f);
}

/* PremultipliedAlpha.cpp:86-93 */
uint16_t avif_oracle_unpremultiply_u16(uint16_t color, uint16_t alpha, uint16_t max_value)
{
    const float max_value_float = (float)max_value;
    const float value = avif_oracle_unpremultiply_f32((float)color, (float)alpha, max_value_float);
    return (uint16_t)min_f(roundf(value), max_value_float);
}

void avif_oracle_premultiply_table_u16(uint16_t max_value, int unpremultiply, uint16_t* out)
{
    const int count = (int)max_value + 1;
    int c, a;
    for (c = 0; c < count; ++c)
    {
        for (a = 0; a < count; ++a)
        {
            uint16_t v;
            if (unpremultiply)
            {
                v = (a == 0) ? 0 : avif_oracle_unpremultiply_u16((uint16_t)c, (uint16_t)a, max_value);
            }
            else
            {
                v = avif_oracle_premultiply_u16((uint16_t)c, (uint16_t)a, max_value);
            }
            out[(size_t)c * (size_t)count + (size_t)a] = v;
        }
    }
}

void avif_oracle_premultiply_table_u8(int unpremultiply, uint8_t* out)
{
    int c, a;
    for (c = 0; c < 256; ++c)
    {
        for (a = 0; a < 256; ++a)
        {
            uint8_t v;
            if (unpremultiply)
            {
                v = (a == 0) ? 0 : avif_oracle_unpremultiply_u8((uint8_t)c, (uint8_t)a);
            }
            else
            {
                v = avif_oracle_premultiply_u8((uint8_t)c, (uint8_t)a);
            }
            out[c * 256 + a] = v;
        }
    }
}

/* ---------------------------------------------------------------------------------------------------- */
/* Matrix coefficients: YUVCoefficiants.cpp                                                               */
/* ---------------------------------------------------------------------------------------------------- */

/* YUVCoefficiants.cpp:56-68: rX rY gX gY bX bY wX wY by H.273 colour-primaries code point. */
static void colour_primaries(int32_t code, float out[8])
{
    static const struct { int32_t code; float v[8]; } table[] = {
        { 1, { 0.64f, 0.33f, 0.3f, 0.6f, 0.15f, 0.06f, 0.3127f, 0.329f } },
        { 4, { 0.67f, 0.33f, 0.21f, 0.71f, 0.14f, 0.08f, 0.310f, 0.316f } },
        { 5, { 0.64f, 0.33f, 0.29f, 0.60f, 0.15f, 0.06f, 0.3127f, 0.3290f } },
        { 6, { 0.630f, 0.340f, 0.310f, 0.595f, 0.155f, 0.070f, 0.3127f, 0.3290f } },
        { 7, { 0.630f, 0.340f, 0.310f, 0.595f, 0.155f, 0.070f, 0.3127f, 0.3290f } },
        { 8, { 0.681f, 0.319f, 0.243f, 0.692f, 0.145f, 0.049f, 0.310f, 0.316f } },
        { 9, { 0.708f, 0.292f, 0.170f, 0.797f, 0.131f, 0.046f, 0.3127f, 0.3290f } },
        { 10, { 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.3333f, 0.3333f } },
        { 11, { 0.680f, 0.320f, 0.265f, 0.690f, 0.150f, 0.060f, 0.314f, 0.351f } },
        { 12, { 0.680f, 0.320f, 0.265f, 0.690f, 0.150f, 0.060f, 0.3127f, 0.3290f } },
        { 22, { 0.630f, 0.340f, 0.295f, 0.605f, 0.155f, 0.077f, 0.3127f, 0.3290f } },
    };
    size_t i;
    for (i = 0; i < sizeof(table) / sizeof(table[0]); ++i)
    {
        if (table[i].code == code)
        {
            memcpy(out, table[i].v, sizeof(table[i].v));
            return;
        }
    }
    /* YUVCoefficiants.cpp:81-82: unknown primaries fall back to the first row. */
    memcpy(out, table[0].v, sizeof(table[0].v));
}

/* YUVCoefficiants.cpp:108-152 */
static int coefficients_from_cicp(const avifgpu_nclx* cicp, float coeffs[3])
{
    /* YUVCoefficiants.cpp:94-106: matrix code point -> kr, kb */
    static const struct { int32_t code; float kr; float kb; } matrices[] = {
        { 1, 0.2126f, 0.0722f }, { 4, 0.30f, 0.11f }, { 5, 0.299f, 0.114f },
        { 6, 0.299f, 0.114f },   { 7, 0.212f, 0.087f }, { 9, 0.2627f, 0.0593f },
    };
    size_t i;
    if (cicp == NULL || !cicp->present)
    {
        return 0;
    }
    if (cicp->matrix_coefficients == 12) /* chromaticity-derived non-constant luminance */
    {
        float p[8];
        colour_primaries(cicp->color_primaries, p);
        {
            const float rX = p[0], rY = p[1], gX = p[2], gY = p[3], bX = p[4], bY = p[5], wX = p[6], wY = p[7];
            const float rZ = 1.0f - (rX + rY);
            const float gZ = 1.0f - (gX + gY);
            const float bZ = 1.0f - (bX + bY);
            const float wZ = 1.0f - (wX + wY);
            const float kr = (rY * (wX * (gY * bZ - bY * gZ) + wY * (bX * gZ - gX * bZ) + wZ * (gX * bY - bX * gY))) /
                             (wY * (rX * (gY * bZ - bY * gZ) + gX * (bY * rZ - rY * bZ) + bX * (rY * gZ - gY * rZ)));
            const float kb = (bY * (wX * (rY * gZ - gY * rZ) + wY * (gX * rZ - rX * gZ) + wZ * (rX * gY - gX * rY))) /
                             (wY * (rX * (gY * bZ - bY * gZ) + gX * (bY * rZ - rY * bZ) + bX * (rY * gZ - gY * rZ)));
            coeffs[0] = kr;
            coeffs[2] = kb;
            coeffs[1] = 1.0f - coeffs[0] - coeffs[2];
            return 1;
        }
    }
    for (i = 0; i < sizeof(matrices) / sizeof(matrices[0]); ++i)
    {
        if (matrices[i].code == cicp->matrix_coefficients)
        {
            coeffs[0] = matrices[i].kr;
            coeffs[2] = matrices[i].kb;
            coeffs[1] = 1.0f - coeffs[0] - coeffs[2];
            return 1;
        }
    }
    return 0;
}

/* YUVCoefficiants.cpp:154-188 */
int avif_oracle_get_yuv_coefficients(const avifgpu_nclx* nclx, float* out)
{
    float kr = 0.299f;
    float kb = 0.114f;
    float kg = 1.0f - kr - kb;
    float coeffs[3];
    if (coefficients_from_cicp(nclx, coeffs))
    {
        kr = coeffs[0];
        kg = coeffs[1];
        kb = coeffs[2];
    }
    out[0] = kr;
    out[1] = kg;
    out[2] = kb;
    return AVIFGPU_OK;
}

/* ---------------------------------------------------------------------------------------------------- */
/* unorm -> float tables: YuvLookupTables.cpp                                                             */
/* ---------------------------------------------------------------------------------------------------- */

/* YuvLookupTables.cpp:64-66: ((v - lo) * full + (hi - lo) / 2) / (hi - lo), clamped to [0, full]; C integer
 * division truncates toward zero exactly as in the reference. */
static int limited_to_full(int v, int lo, int hi, int full)
{
    v = (((v - lo) * full) + ((hi - lo) / 2)) / (hi - lo);
    return (v > full) ? full : ((v < 0) ? 0 : v);
}

/* YuvLookupTables.cpp:69-88 */
static int limited_to_full_y(int depth, int v)
{
    switch (depth)
    {
    case 8: return limited_to_full(v, 16, 235, 255);
    case 10: return limited_to_full(v, 64, 940, 1023);
    case 12: return limited_to_full(v, 256, 3760, 4095);
    default: return limited_to_full(v, 1024, 60160, 65535);
    }
}

/* YuvLookupTables.cpp:90-109 */
static int limited_to_full_uv(int depth, int v)
{
    switch (depth)
    {
    case 8: return limited_to_full(v, 16, 240, 255);
    case 10: return limited_to_full(v, 64, 960, 1023);
    case 12: return limited_to_full(v, 256, 3840, 4095);
    default: return limited_to_full(v, 1024, 61440, 65535);
    }
}

typedef struct yuv_tables
{
    float* y;
    float* uv;
    float* alpha;
    int max_channel;
} yuv_tables;

/* YuvLookupTables.cpp:115-192 */
static int build_tables(const avifgpu_nclx* nclx, int bit_depth, int monochrome, int has_alpha, float* y, float* uv,
                        float* alpha)
{
    int i;
    if (bit_depth != 8 && bit_depth != 10 && bit_depth != 12 && bit_depth != 16)
    {
        return fail(AVIFGPU_ERR_UNSUPPORTED, "The image has an unsupported bit depth, must be 8, 10, 12 or 16.");
    }
    {
        const int has_nclx = nclx != NULL && nclx->present;
        const int full_range = has_nclx ? (nclx->full_range_flag != 0) : 1;
        const int matrix = has_nclx ? nclx->matrix_coefficients : 6;
        const int count = 1 << bit_depth;
        const int is_color = !monochrome;
        const int is_identity = is_color && matrix == 0;
        const float max_float = (float)(count - 1);
        for (i = 0; i < count; ++i)
        {
            int unorm_y = i;
            int unorm_uv = i;
            if (!full_range)
            {
                unorm_y = limited_to_full_y(bit_depth, unorm_y);
                if (is_color)
                {
                    unorm_uv = limited_to_full_uv(bit_depth, unorm_uv);
                }
            }
            if (y) y[i] = (float)unorm_y / max_float;
            if (is_color && uv)
            {
                if (is_identity)
                {
                    uv[i] = (float)unorm_y / max_float;
                }
                else
                {
                    uv[i] = (float)unorm_uv / max_float - 0.5f;
                }
            }
            if (has_alpha && alpha)
            {
                alpha[i] = (float)i / max_float;
            }
        }
    }
    return AVIFGPU_OK;
}

int avif_oracle_build_yuv_tables(const avifgpu_nclx* nclx, int32_t bit_depth, int32_t monochrome, float* out_y,
                                 float* out_uv, float* out_alpha)
{
    return build_tables(nclx, bit_depth, monochrome, out_alpha != NULL, out_y, out_uv, out_alpha);
}

/* ---------------------------------------------------------------------------------------------------- */
/* Encode: WriteHeifImage.cpp                                                                             */
/* ---------------------------------------------------------------------------------------------------- */

/* WriteHeifImage.cpp:87-166: (int)((i / from_max) * to_max + 0.5f) clamped to [0, to_max]. */
static int depth_lut_entry(int i, float from_max, int to_max)
{
    int value = (int)((((float)i / from_max) * (float)to_max) + 0.5f);
    if (value < 0)
    {
        value = 0;
    }
    else if (value > to_max)
    {
        value = to_max;
    }
    return value;
}

int avif_oracle_build_depth_lut(int32_t host_depth, int32_t image_bit_depth, uint16_t* out)
{
    const int to_max = (1 << image_bit_depth) - 1;
    int i;
    if (host_depth == 8)
    {
        for (i = 0; i < 256; ++i) out[i] = (uint16_t)depth_lut_entry(i, 255.0f, to_max);
        return AVIFGPU_OK;
    }
    if (host_depth == 16)
    {
        for (i = 0; i < 32769; ++i) out[i] = (uint16_t)depth_lut_entry(i, 32768.0f, to_max);
        return AVIFGPU_OK;
    }
    return fail(AVIFGPU_ERR_BAD_PARAM, "depth lut: host depth must be 8 or 16");
}

/* WriteHeifImage.cpp:590,618,1093-1096,1128-1130: static_cast<uint16_t>(std::clamp(v * max, 0, max)).
 * A NaN survives std::clamp and the cast is then undefined; x86 cvttss2si yields 0 in the low 16 bits and this
 * project DEFINES NaN -> code 0. */
static uint16_t float_to_code(float v, float max_value)
{
    const float scaled = clamp_f(v * max_value, 0.0f, max_value);
    if (scaled != scaled)
    {
        return 0;
    }
    return (uint16_t)scaled;
}

/* Host sample -> image code for integer hosts.  8-bit host to 8-bit image copies (WriteHeifImage.cpp:263-331,
 * 730-803); otherwise the look-up tables of WriteHeifImage.cpp:87-166.  A 16-bit host sample above 32768
 * indexes past the reference's 32769-entry table (undefined); this project DEFINES it as the same formula
 * evaluated for that value (then clamped), which is what depth_lut_entry does. */
static uint16_t integer_host_to_code(int host_value, int host_depth, int image_depth)
{
    if (host_depth == 8)
    {
        if (image_depth == 8)
        {
            return (uint16_t)host_value;
        }
        return (uint16_t)depth_lut_entry(host_value, 255.0f, (1 << image_depth) - 1);
    }
    return (uint16_t)depth_lut_entry(host_value, 32768.0f, (1 << image_depth) - 1);
}

/* WriteHeifImage.cpp:238-251, 404-417, 700-718 ...: premultiply an integer code by an integer alpha code. */
static uint16_t premultiply_code(uint16_t color, uint16_t alpha, int image_depth)
{
    const uint16_t max_value = (uint16_t)((1 << image_depth) - 1);
    if (alpha < max_value)
    {
        if (alpha == 0)
        {
            return 0;
        }
        if (image_depth == 8)
        {
            return avif_oracle_premultiply_u8((uint8_t)color, (uint8_t)alpha);
        }
        return avif_oracle_premultiply_u16(color, alpha, max_value);
    }
    return color;
}

/* WriteHeifImage.cpp:1079-1091 (colour) / 578-588 (gray): the OETF switch.  hlg_extension != 0 adds the HLG save path
 * this project defines on top of the reference's uncalled LinearToHLG (include/avifgpu.h). */
static int apply_oetf(int transfer, float peak, int is_gray, int hlg_extension, float v, float* out)
{
    switch (transfer)
    {
    case AVIFGPU_TRANSFER_HLG:
        if (is_gray || (hlg_extension != AVIFGPU_HLG_OETF && hlg_extension != AVIFGPU_HLG_INVERSE_OOTF_THEN_OETF)) return 0;
        *out = linear_to_hlg(v);
        return 1;
    case AVIFGPU_TRANSFER_PQ:
        *out = linear_to_pq(v, peak);
        return 1;
    case AVIFGPU_TRANSFER_SMPTE428:
        if (is_gray) return 0; /* WriteHeifImage.cpp:586-587 */
        *out = linear_to_smpte428(v);
        return 1;
    case AVIFGPU_TRANSFER_CLIP:
        *out = v;
        return 1;
    default:
        return 0;
    }
}

/* One host pixel -> up to four integer codes (R,G,B,A or Y,A), exactly as the reference's inner loops.
 * Returns 0 for an unsupported transfer function. */
static int host_pixel_to_codes(const avifgpu_encode_desc* d, const void* pixel, uint16_t codes[4])
{
    const int channels = d->host_channels;
    const int colors = (channels <= 2) ? 1 : 3;
    const int has_alpha = d->alpha_state != AVIFGPU_ALPHA_NONE;
    const int premultiply = d->alpha_state == AVIFGPU_ALPHA_PREMULTIPLIED;
    const int depth = d->image_bit_depth;
    int i;

    if (d->host_depth == 32)
    {
        /* WriteHeifImage.cpp:502-627 (gray), 990-1139 (colour) */
        const float* src = (const float*)pixel;
        const float max_value = (float)((1 << depth) - 1);
        const float peak = (float)d->pq_peak_nits;
        float color[3];
        float alpha = 0.0f;
        for (i = 0; i < colors; ++i) color[i] = src[i];
        if (d->row_matrix_enabled && colors == 3)
        {
            /* The colour-profile step: ConvertRow runs over the raw host row before the per-pixel loop
             * (WriteHeifImage.cpp:1028-1031); alpha is copied (cmsFLAGS_COPY_ALPHA).  This project's definition of the
             * matrix case, see include/avifgpu.h (parity unpinned: lcms2 is not in the reference tree). */
            const float r = color[0], g = color[1], b = color[2];
            const float* m = d->row_matrix;
            color[0] = ((m[0] * r) + (m[1] * g)) + (m[2] * b);
            color[1] = ((m[3] * r) + (m[4] * g)) + (m[5] * b);
            color[2] = ((m[6] * r) + (m[7] * g)) + (m[8] * b);
        }
        if (has_alpha)
        {
            alpha = clamp_f(src[colors], 0.0f, 1.0f);
            if (premultiply)
            {
                if (alpha < 1.0f)
                {
                    if (alpha == 0)
                    {
                        for (i = 0; i < colors; ++i) color[i] = 0;
                    }
                    else
                    {
                        for (i = 0; i < colors; ++i)
                        {
                            color[i] = avif_oracle_premultiply_f32(clamp_f(color[i], 0.0f, 1.0f), alpha, 1.0f);
                        }
                    }
                }
            }
        }
        else if (colors == 1)
        {
            /* WriteHeifImage.cpp:602: the gray no-alpha path clamps to [0,1] before the transfer curve. */
            color[0] = clamp_f(color[0], 0.0f, 1.0f);
        }
        if (colors == 3 && d->transfer == AVIFGPU_TRANSFER_HLG && d->hlg_extension == AVIFGPU_HLG_INVERSE_OOTF_THEN_OETF)
        {
            /* display-referred input: ColorTransfer.cpp:207-220 on the pixel, then the OETF per channel */
            float luma[3];
            if (!d->nclx.present || avif_oracle_get_hlg_luma_coefficients(d->nclx.color_primaries, luma) != AVIFGPU_OK)
            {
                return 0;
            }
            apply_inverse_hlg_ootf(color, luma, d->hlg_display_gamma, (float)d->hlg_peak_nits);
        }
        for (i = 0; i < colors; ++i)
        {
            float curved;
            if (!apply_oetf(d->transfer, peak, colors == 1, d->hlg_extension, color[i], &curved))
            {
                return 0;
            }
            codes[i] = float_to_code(curved, max_value);
        }
        if (has_alpha)
        {
            codes[colors] = float_to_code(alpha, max_value);
        }
        return 1;
    }

    /* Integer hosts: WriteHeifImage.cpp:169-334 / 336-500 (gray), 629-806 / 808-988 (colour). */
    {
        int host[4];
        if (d->host_depth == 8)
        {
            const uint8_t* src = (const uint8_t*)pixel;
            for (i = 0; i < channels; ++i) host[i] = src[i];
        }
        else
        {
            const uint16_t* src = (const uint16_t*)pixel;
            for (i = 0; i < channels; ++i) host[i] = src[i];
        }
        if (d->host_depth == 16 && colors == 1 && d->gray16_curve == AVIFGPU_GRAY16_SMPTE428)
        {
            /* BASELINE.json config 5, defined by this project (SURVEY.md section 8c "C5 caveat"):
             * the quantiser of WriteHeifImage.cpp:618 applied to ColorTransfer.cpp:119-127 of v / 32768. */
            const float max_value = (float)((1 << depth) - 1);
            codes[0] = float_to_code(linear_to_smpte428((float)host[0] / 32768.0f), max_value);
            if (has_alpha)
            {
                codes[1] = integer_host_to_code(host[1], 16, depth);
            }
            return 1;
        }
        for (i = 0; i < channels; ++i)
        {
            codes[i] = integer_host_to_code(host[i], d->host_depth, depth);
        }
        if (has_alpha && premultiply)
        {
            for (i = 0; i < colors; ++i)
            {
                codes[i] = premultiply_code(codes[i], codes[colors], depth);
            }
        }
    }
    return 1;
}

static void store_code(void* row, int index, int bytes_per_sample, uint16_t code)
{
    if (bytes_per_sample == 1)
    {
        ((uint8_t*)row)[index] = (uint8_t)code;
    }
    else
    {
        ((uint16_t*)row)[index] = code;
    }
}

/*
 * Forward matrix (PARITY UNPINNED, see avif_oracle.h): the algebraic inverse of the reference decoder
 * (YuvDecode.cpp:555-557: R = Y + 2(1-kr)Cr, B = Y + 2(1-kb)Cb) on integer codes, full range
 * (WriteMetadata.cpp:46), float32, no contraction, with the two chroma gains as float constants (the way
 * libheif holds its RGB->YCbCr coefficients):
 *     Y  = (kr*R + kg*G) + kb*B
 *     Cb = (B - Y) * cb_scale,  cb_scale = 0.5f / (1 - kb)
 *     Cr = (R - Y) * cr_scale,  cr_scale = 0.5f / (1 - kr)
 *     Ycode = clamp((int)(Y + 0.5f)),  Ccode = clamp((int)((C + 2^(depth-1)) + 0.5f))   (H.273 full range:
 *     Clip(Round(C) + 2^(depth-1)), what libheif's RGB->YCbCr writes; a pure red / blue at max reaches 2^depth and is clipped)
 * Identity matrix (lossless GBR, WriteMetadata.cpp:143-146): Y = G, Cb = B, Cr = R.
 */
typedef struct forward_matrix
{
    float kr, kg, kb;
    float cb_scale; /* 0.5f / (1-kb) */
    float cr_scale; /* 0.5f / (1-kr) */
    int identity;
} forward_matrix;

static void forward_matrix_init(forward_matrix* m, const avifgpu_nclx* nclx)
{
    float k[3];
    avif_oracle_get_yuv_coefficients(nclx, k);
    m->kr = k[0];
    m->kg = k[1];
    m->kb = k[2];
    m->cb_scale = 0.5f / (1.0f - m->kb);
    m->cr_scale = 0.5f / (1.0f - m->kr);
    m->identity = nclx != NULL && nclx->present && nclx->matrix_coefficients == 0;
}

static void forward_pixel(const forward_matrix* m, const uint16_t rgb[3], float* y, float* cb, float* cr)
{
    const float r = (float)rgb[0];
    const float g = (float)rgb[1];
    const float b = (float)rgb[2];
    if (m->identity)
    {
        *y = g;
        *cb = b;
        *cr = r;
        return;
    }
    *y = ((m->kr * r) + (m->kg * g)) + (m->kb * b);
    *cb = (b - *y) * m->cb_scale;
    *cr = (r - *y) * m->cr_scale;
}

static uint16_t quantise_luma(float y, int max_code)
{
    int v = (int)(y + 0.5f);
    return (uint16_t)((v < 0) ? 0 : ((v > max_code) ? max_code : v));
}

static uint16_t quantise_chroma(const forward_matrix* m, float c, int depth)
{
    const int max_code = (1 << depth) - 1;
    int v;
    if (m->identity)
    {
        v = (int)(c + 0.5f);
    }
    else
    {
        /* H.273 full range: chroma zero on 2^(depth-1) (the reference decoder's own zero, max/2, is half a code lower) */
        v = (int)((c + (float)(1 << (depth - 1))) + 0.5f);
    }
    return (uint16_t)((v < 0) ? 0 : ((v > max_code) ? max_code : v));
}

typedef struct encode_job
{
    const avifgpu_encode_desc* desc;
    const uint8_t* rows;
    int64_t row_stride;
    avifgpu_planes dst;
    int y_begin; /* even for 4:2:0 */
    int y_end;
    int status;
} encode_job;

static int encode_rows_reference_layout(const encode_job* job)
{
    const avifgpu_encode_desc* d = job->desc;
    const int channels = d->host_channels;
    const int gray = channels <= 2;
    const int has_alpha = d->alpha_state != AVIFGPU_ALPHA_NONE;
    const int col_bytes = channels * ((d->host_depth + 7) / 8);
    const int bytes_per_sample = d->image_bit_depth > 8 ? 2 : 1;
    int x, y, i;
    for (y = job->y_begin; y < job->y_end; ++y)
    {
        const uint8_t* src = job->rows + (int64_t)y * job->row_stride;
        uint8_t* out0 = (uint8_t*)job->dst.data[0] + (int64_t)y * job->dst.stride[0];
        uint8_t* out_alpha = (gray && has_alpha) ? (uint8_t*)job->dst.data[3] + (int64_t)y * job->dst.stride[3] : NULL;
        for (x = 0; x < d->width; ++x)
        {
            uint16_t codes[4];
            if (!host_pixel_to_codes(d, src + (int64_t)x * col_bytes, codes))
            {
                return fail(AVIFGPU_ERR_UNSUPPORTED, "Unsupported color transfer function.");
            }
            if (gray)
            {
                store_code(out0, x, bytes_per_sample, codes[0]);
                if (has_alpha) store_code(out_alpha, x, bytes_per_sample, codes[1]);
            }
            else
            {
                for (i = 0; i < channels; ++i) store_code(out0, x * channels + i, bytes_per_sample, codes[i]);
            }
        }
    }
    return AVIFGPU_OK;
}

/* Fused host rows -> planar YCbCr(A).  Chroma site (cx, cy) covers pixels x in {2cx, 2cx+1} (4:2:2, 4:2:0) and
 * rows y in {2cy, 2cy+1} (4:2:0).  BOX: ((c00 + c01) + (c10 + c11)) * 0.25f; with a missing column / row at an
 * odd edge the mean is over the samples that exist: (c00 + c10) * 0.5f, (c00 + c01) * 0.5f or c00.
 * TOP_LEFT: c00. */
static int encode_rows_planar_ycbcr(const encode_job* job, const void* codes_image, int64_t codes_stride)
{
    const avifgpu_encode_desc* d = job->desc;
    const int channels = d->host_channels;
    const int has_alpha = d->alpha_state != AVIFGPU_ALPHA_NONE;
    const int col_bytes = channels * ((d->host_depth + 7) / 8);
    const int depth = d->image_bit_depth;
    const int max_code = (1 << depth) - 1;
    const int bytes_per_sample = depth > 8 ? 2 : 1;
    const int xs = (d->chroma == AVIFGPU_CHROMA_420 || d->chroma == AVIFGPU_CHROMA_422) ? 1 : 0;
    const int ys = (d->chroma == AVIFGPU_CHROMA_420) ? 1 : 0;
    const int step_y = 1 << ys;
    const int step_x = 1 << xs;
    forward_matrix m;
    int x, y, dx, dy;
    forward_matrix_init(&m, &d->nclx);
    if (m.identity && (xs || ys))
    {
        return fail(AVIFGPU_ERR_UNSUPPORTED, "identity (GBR) matrix requires 4:4:4");
    }

    for (y = job->y_begin; y < job->y_end; y += step_y)
    {
        uint8_t* out_cb = (uint8_t*)job->dst.data[1] + (int64_t)(y >> ys) * job->dst.stride[1];
        uint8_t* out_cr = (uint8_t*)job->dst.data[2] + (int64_t)(y >> ys) * job->dst.stride[2];
        for (x = 0; x < d->width; x += step_x)
        {
            float cb[2][2], cr[2][2];
            int have[2][2] = { { 0, 0 }, { 0, 0 } };
            for (dy = 0; dy < step_y; ++dy)
            {
                const int yy = y + dy;
                if (yy >= d->height) continue;
                for (dx = 0; dx < step_x; ++dx)
                {
                    const int xx = x + dx;
                    uint16_t codes[4];
                    float yf;
                    if (xx >= d->width) continue;
                    if (codes_image != NULL)
                    {
                        /* pre-quantised interleaved codes (avif_oracle_rgb_codes_to_ycbcr) */
                        const uint8_t* p = (const uint8_t*)codes_image + (int64_t)yy * codes_stride;
                        int i;
                        for (i = 0; i < channels; ++i)
                        {
                            codes[i] = (bytes_per_sample == 1) ? p[xx * channels + i] : ((const uint16_t*)p)[xx * channels + i];
                        }
                    }
                    else if (!host_pixel_to_codes(d, job->rows + (int64_t)yy * job->row_stride + (int64_t)xx * col_bytes, codes))
                    {
                        return fail(AVIFGPU_ERR_UNSUPPORTED, "Unsupported color transfer function.");
                    }
                    forward_pixel(&m, codes, &yf, &cb[dy][dx], &cr[dy][dx]);
                    have[dy][dx] = 1;
                    store_code((uint8_t*)job->dst.data[0] + (int64_t)yy * job->dst.stride[0], xx, bytes_per_sample,
                               quantise_luma(yf, max_code));
                    if (has_alpha)
                    {
                        store_code((uint8_t*)job->dst.data[3] + (int64_t)yy * job->dst.stride[3], xx, bytes_per_sample, codes[3]);
                    }
                }
            }
            {
                float cbv, crv;
                if (d->down_filter == AVIFGPU_DOWN_FILTER_TOP_LEFT || (!xs && !ys))
                {
                    cbv = cb[0][0];
                    crv = cr[0][0];
                }
                else if (have[0][1] && have[1][0])
                {
                    cbv = ((cb[0][0] + cb[0][1]) + (cb[1][0] + cb[1][1])) * 0.25f;
                    crv = ((cr[0][0] + cr[0][1]) + (cr[1][0] + cr[1][1])) * 0.25f;
                }
                else if (have[0][1])
                {
                    cbv = (cb[0][0] + cb[0][1]) * 0.5f;
                    crv = (cr[0][0] + cr[0][1]) * 0.5f;
                }
                else if (have[1][0])
                {
                    cbv = (cb[0][0] + cb[1][0]) * 0.5f;
                    crv = (cr[0][0] + cr[1][0]) * 0.5f;
                }
                else
                {
                    cbv = cb[0][0];
                    crv = cr[0][0];
                }
                store_code(out_cb, x >> xs, bytes_per_sample, quantise_chroma(&m, cbv, depth));
                store_code(out_cr, x >> xs, bytes_per_sample, quantise_chroma(&m, crv, depth));
            }
        }
    }
    return AVIFGPU_OK;
}

static int validate_encode_desc(const avifgpu_encode_desc* d)
{
    if (d == NULL || d->struct_size != sizeof(avifgpu_encode_desc)) return fail(AVIFGPU_ERR_BAD_PARAM, "bad encode desc size");
    if (d->width < 0 || d->height < 0) return fail(AVIFGPU_ERR_BAD_PARAM, "negative image size");
    if (d->host_depth != 8 && d->host_depth != 16 && d->host_depth != 32) return fail(AVIFGPU_ERR_BAD_PARAM, "host depth must be 8, 16 or 32");
    if (d->host_channels < 1 || d->host_channels > 4) return fail(AVIFGPU_ERR_BAD_PARAM, "host channels must be 1..4");
    if (d->image_bit_depth != 8 && d->image_bit_depth != 10 && d->image_bit_depth != 12) return fail(AVIFGPU_ERR_BAD_PARAM, "image bit depth must be 8, 10 or 12");
    {
        const int expects_alpha = d->host_channels == 2 || d->host_channels == 4;
        const int has_alpha = d->alpha_state != AVIFGPU_ALPHA_NONE;
        if (expects_alpha != has_alpha) return fail(AVIFGPU_ERR_BAD_PARAM, "alpha state does not match the channel count");
    }
    if (d->host_depth == 32 && d->image_bit_depth == 8)
    {
        /* WriteHeifImage.cpp writes uint16 samples for 32-bit hosts; an 8-bit heif plane would be overrun. */
        return fail(AVIFGPU_ERR_UNSUPPORTED, "32-bit hosts require a 10- or 12-bit image");
    }
    if (d->layout == AVIFGPU_LAYOUT_PLANAR_YCBCR)
    {
        if (d->host_channels <= 2) return fail(AVIFGPU_ERR_BAD_PARAM, "planar YCbCr needs a colour host");
        if (d->chroma != AVIFGPU_CHROMA_420 && d->chroma != AVIFGPU_CHROMA_422 && d->chroma != AVIFGPU_CHROMA_444) return fail(AVIFGPU_ERR_BAD_PARAM, "bad chroma");
        if (d->nclx.present && !d->nclx.full_range_flag) return fail(AVIFGPU_ERR_UNSUPPORTED, "the encode path is full range only");
    }
    else if (d->layout != AVIFGPU_LAYOUT_REFERENCE)
    {
        return fail(AVIFGPU_ERR_BAD_PARAM, "bad layout");
    }
    return AVIFGPU_OK;
}

static void* encode_thread(void* arg)
{
    encode_job* job = (encode_job*)arg;
    if (job->desc->layout == AVIFGPU_LAYOUT_REFERENCE)
    {
        job->status = encode_rows_reference_layout(job);
    }
    else
    {
        job->status = encode_rows_planar_ycbcr(job, NULL, 0);
    }
    return NULL;
}

/* Row-block boundaries that keep 4:2:0 row pairs together. */
static int split_rows(int height, int parts, int* bounds)
{
    int count = 0;
    int i;
    bounds[count++] = 0;
    for (i = 1; i < parts; ++i)
    {
        int b = (int)(((int64_t)height * i) / parts) & ~1;
        if (b > bounds[count - 1] && b < height)
        {
            bounds[count++] = b;
        }
    }
    bounds[count++] = height;
    return count - 1;
}

#define AVIF_ORACLE_MAX_THREADS 256

int avif_oracle_encode_image_mt(const avifgpu_encode_desc* desc, const void* host_rows, int64_t row_stride,
                                const avifgpu_planes* dst, int32_t threads)
{
    int bounds[AVIF_ORACLE_MAX_THREADS + 2];
    encode_job jobs[AVIF_ORACLE_MAX_THREADS];
    pthread_t handles[AVIF_ORACLE_MAX_THREADS];
    int blocks, i;
    const int status = validate_encode_desc(desc);
    if (status != AVIFGPU_OK) return status;
    if (threads < 1) threads = 1;
    if (threads > AVIF_ORACLE_MAX_THREADS) threads = AVIF_ORACLE_MAX_THREADS;
    blocks = split_rows(desc->height, threads, bounds);
    for (i = 0; i < blocks; ++i)
    {
        jobs[i].desc = desc;
        jobs[i].rows = (const uint8_t*)host_rows;
        jobs[i].row_stride = row_stride;
        jobs[i].dst = *dst;
        jobs[i].y_begin = bounds[i];
        jobs[i].y_end = bounds[i + 1];
        jobs[i].status = AVIFGPU_OK;
    }
    if (blocks == 1)
    {
        encode_thread(&jobs[0]);
        return jobs[0].status;
    }
    for (i = 0; i < blocks; ++i) pthread_create(&handles[i], NULL, encode_thread, &jobs[i]);
    for (i = 0; i < blocks; ++i) pthread_join(handles[i], NULL);
    for (i = 0; i < blocks; ++i)
    {
        if (jobs[i].status != AVIFGPU_OK) return fail(jobs[i].status, "encode block failed (unsupported transfer function?)");
    }
    return AVIFGPU_OK;
}

int avif_oracle_encode_image(const avifgpu_encode_desc* desc, const void* host_rows, int64_t row_stride,
                             const avifgpu_planes* dst)
{
    return avif_oracle_encode_image_mt(desc, host_rows, row_stride, dst, 1);
}

typedef struct codes_job
{
    encode_job job;
    const void* interleaved;
    int64_t interleaved_stride;
} codes_job;

static void* codes_thread(void* arg)
{
    codes_job* c = (codes_job*)arg;
    c->job.status = encode_rows_planar_ycbcr(&c->job, c->interleaved, c->interleaved_stride);
    return NULL;
}

int avif_oracle_rgb_codes_to_ycbcr_mt(const avifgpu_encode_desc* desc, const void* interleaved, int64_t interleaved_stride,
                                      const avifgpu_planes* dst, int32_t threads)
{
    int bounds[AVIF_ORACLE_MAX_THREADS + 2];
    codes_job jobs[AVIF_ORACLE_MAX_THREADS];
    pthread_t handles[AVIF_ORACLE_MAX_THREADS];
    int blocks, i;
    const int status = validate_encode_desc(desc);
    if (status != AVIFGPU_OK) return status;
    if (desc->layout != AVIFGPU_LAYOUT_PLANAR_YCBCR) return fail(AVIFGPU_ERR_BAD_PARAM, "layout must be planar YCbCr");
    if (threads < 1) threads = 1;
    if (threads > AVIF_ORACLE_MAX_THREADS) threads = AVIF_ORACLE_MAX_THREADS;
    blocks = split_rows(desc->height, threads, bounds);
    for (i = 0; i < blocks; ++i)
    {
        jobs[i].job.desc = desc;
        jobs[i].job.rows = NULL;
        jobs[i].job.row_stride = 0;
        jobs[i].job.dst = *dst;
        jobs[i].job.y_begin = bounds[i];
        jobs[i].job.y_end = bounds[i + 1];
        jobs[i].job.status = AVIFGPU_OK;
        jobs[i].interleaved = interleaved;
        jobs[i].interleaved_stride = interleaved_stride;
    }
    if (blocks == 1)
    {
        codes_thread(&jobs[0]);
        return jobs[0].job.status;
    }
    for (i = 0; i < blocks; ++i) pthread_create(&handles[i], NULL, codes_thread, &jobs[i]);
    for (i = 0; i < blocks; ++i) pthread_join(handles[i], NULL);
    for (i = 0; i < blocks; ++i)
    {
        if (jobs[i].job.status != AVIFGPU_OK) return jobs[i].job.status;
    }
    return AVIFGPU_OK;
}

int avif_oracle_rgb_codes_to_ycbcr(const avifgpu_encode_desc* desc, const void* interleaved, int64_t interleaved_stride,
                                   const avifgpu_planes* dst)
{
    return avif_oracle_rgb_codes_to_ycbcr_mt(desc, interleaved, interleaved_stride, dst, 1);
}

/* ---------------------------------------------------------------------------------------------------- */
/* Decode: YuvDecode.cpp + ReadHeifImage.cpp                                                              */
/* ---------------------------------------------------------------------------------------------------- */

typedef struct decode_job
{
    const avifgpu_decode_desc* desc;
    avifgpu_planes src;
    uint8_t* rows;
    int64_t row_stride;
    const float* table_y;
    const float* table_uv;
    const float* table_alpha;
    const float* table_unorm; /* ReadHeifImage.cpp:402-415 */
    float kr, kg, kb;
    float hlg_luma[3];
    int transfer; /* avifgpu_transfer, host_depth 32 only */
    int y_begin;
    int y_end;
    int status;
} decode_job;

static unsigned load_sample(const void* plane, int64_t stride, int x, int y, int bytes_per_sample)
{
    const uint8_t* row = (const uint8_t*)plane + (int64_t)y * stride;
    return (bytes_per_sample == 1) ? row[x] : ((const uint16_t*)row)[x];
}

/* ColorTransfer.cpp:47-67 */
static int transfer_from_nclx(int32_t transfer_characteristics, int* out)
{
    switch (transfer_characteristics)
    {
    case 16: *out = AVIFGPU_TRANSFER_PQ; return 1;
    case 18: *out = AVIFGPU_TRANSFER_HLG; return 1;
    case 17: *out = AVIFGPU_TRANSFER_SMPTE428; return 1;
    default: return 0;
    }
}

/* YuvDecode.cpp:559-588 / 660-689 and ReadHeifImage.cpp:1062-1090: the EOTF switch on an RGB triple. */
static void apply_eotf_rgb(const decode_job* job, const float in[3], float* out)
{
    const avifgpu_decode_desc* d = job->desc;
    switch (job->transfer)
    {
    case AVIFGPU_TRANSFER_PQ:
        out[0] = pq_to_linear(in[0], (float)d->pq_peak_nits);
        out[1] = pq_to_linear(in[1], (float)d->pq_peak_nits);
        out[2] = pq_to_linear(in[2], (float)d->pq_peak_nits);
        break;
    case AVIFGPU_TRANSFER_HLG:
        out[0] = hlg_to_linear(in[0]);
        out[1] = hlg_to_linear(in[1]);
        out[2] = hlg_to_linear(in[2]);
        if (d->hlg_apply_ootf)
        {
            apply_hlg_ootf(out, job->hlg_luma, d->hlg_display_gamma, (float)d->hlg_peak_nits);
        }
        break;
    default: /* SMPTE428 */
        out[0] = smpte428_to_linear(in[0]);
        out[1] = smpte428_to_linear(in[1]);
        out[2] = smpte428_to_linear(in[2]);
        break;
    }
}

/* Rows of a YCbCr image: ReadHeifImage.cpp:83-400 driving YuvDecode.cpp:281-696. */
static void decode_ycbcr_rows(const decode_job* job)
{
    const avifgpu_decode_desc* d = job->desc;
    const int has_alpha = d->alpha_state != AVIFGPU_ALPHA_NONE;
    const int premultiplied = d->alpha_state == AVIFGPU_ALPHA_PREMULTIPLIED;
    const int bps = d->bit_depth > 8 ? 2 : 1;
    const int yuv_max = (1 << d->bit_depth) - 1;
    const int xs = (d->chroma == AVIFGPU_CHROMA_420 || d->chroma == AVIFGPU_CHROMA_422) ? 1 : 0; /* ReadHeifImage.cpp:52-81 */
    const int ys = (d->chroma == AVIFGPU_CHROMA_420) ? 1 : 0;
    const int channels = has_alpha ? 4 : 3;
    const float kr = job->kr, kg = job->kg, kb = job->kb;
    int x, y;
    for (y = job->y_begin; y < job->y_end; ++y)
    {
        const int uv_j = y >> ys; /* ReadHeifImage.cpp:359 */
        uint8_t* dst_row = job->rows + (int64_t)y * job->row_stride;
        for (x = 0; x < d->width; ++x)
        {
            const int uv_i = x >> xs;
            unsigned unorm_y = load_sample(job->src.data[0], job->src.stride[0], x, y, bps);
            unsigned unorm_u = load_sample(job->src.data[1], job->src.stride[1], uv_i, uv_j, bps);
            unsigned unorm_v = load_sample(job->src.data[2], job->src.stride[2], uv_i, uv_j, bps);
            unsigned unorm_a = has_alpha ? load_sample(job->src.data[3], job->src.stride[3], x, y, bps) : 0;
            float Y, Cb, Cr, R, G, B, A = 0.0f;
            if (d->host_depth != 8)
            {
                /* the 16-bit-container variants clamp codes (YuvDecode.cpp:425-427 ...); the 8-bit ones cannot overflow */
                if (unorm_y > (unsigned)yuv_max) unorm_y = (unsigned)yuv_max;
                if (unorm_u > (unsigned)yuv_max) unorm_u = (unsigned)yuv_max;
                if (unorm_v > (unsigned)yuv_max) unorm_v = (unsigned)yuv_max;
                if (unorm_a > (unsigned)yuv_max) unorm_a = (unsigned)yuv_max;
            }
            Y = job->table_y[unorm_y];
            Cb = job->table_uv[unorm_u];
            Cr = job->table_uv[unorm_v];

            /* YuvDecode.cpp:306-312 (identical in all six colour variants) */
            R = Y + (2 * (1 - kr)) * Cr;
            B = Y + (2 * (1 - kb)) * Cb;
            G = Y - ((2 * ((kr * (1 - kr) * Cr) + (kb * (1 - kb) * Cb))) / kg);
            R = clamp_f(R, 0.0f, 1.0f);
            G = clamp_f(G, 0.0f, 1.0f);
            B = clamp_f(B, 0.0f, 1.0f);

            if (has_alpha)
            {
                A = job->table_alpha[unorm_a];
                if (premultiplied)
                {
                    if (unorm_a < (unsigned)yuv_max)
                    {
                        if (unorm_a == 0)
                        {
                            R = 0; G = 0; B = 0;
                        }
                        else
                        {
                            R = avif_oracle_unpremultiply_f32(R, A, 1.0f);
                            G = avif_oracle_unpremultiply_f32(G, A, 1.0f);
                            B = avif_oracle_unpremultiply_f32(B, A, 1.0f);
                        }
                    }
                }
            }

            if (d->host_depth == 8)
            {
                /* YuvDecode.cpp:318-320, 391-394 */
                uint8_t* p = dst_row + x * channels;
                p[0] = (uint8_t)(0.5f + (R * 255.0f));
                p[1] = (uint8_t)(0.5f + (G * 255.0f));
                p[2] = (uint8_t)(0.5f + (B * 255.0f));
                if (has_alpha) p[3] = (uint8_t)unorm_a;
            }
            else if (d->host_depth == 16)
            {
                /* YuvDecode.cpp:441-443, 511-514 */
                uint16_t* p = (uint16_t*)dst_row + x * channels;
                p[0] = (uint16_t)(0.5f + (R * 32768.0f));
                p[1] = (uint16_t)(0.5f + (G * 32768.0f));
                p[2] = (uint16_t)(0.5f + (B * 32768.0f));
                if (has_alpha) p[3] = (uint16_t)(0.5f + (A * 32768.0f));
            }
            else
            {
                /* YuvDecode.cpp:559-592, 660-692 */
                float* p = (float*)dst_row + x * channels;
                const float rgb[3] = { R, G, B };
                apply_eotf_rgb(job, rgb, p);
                if (has_alpha) p[3] = A;
            }
        }
    }
}

/* Rows of a monochrome image: ReadHeifImage.cpp:418-559, 863-947 driving YuvDecode.cpp:55-279. */
static void decode_gray_rows(const decode_job* job)
{
    const avifgpu_decode_desc* d = job->desc;
    const int has_alpha = d->alpha_state != AVIFGPU_ALPHA_NONE;
    const int premultiplied = d->alpha_state == AVIFGPU_ALPHA_PREMULTIPLIED;
    const int bps = d->bit_depth > 8 ? 2 : 1;
    const int yuv_max = (1 << d->bit_depth) - 1;
    const int channels = has_alpha ? 2 : 1;
    int x, y;
    for (y = job->y_begin; y < job->y_end; ++y)
    {
        uint8_t* dst_row = job->rows + (int64_t)y * job->row_stride;
        for (x = 0; x < d->width; ++x)
        {
            unsigned unorm_y = load_sample(job->src.data[0], job->src.stride[0], x, y, bps);
            unsigned unorm_a = has_alpha ? load_sample(job->src.data[3], job->src.stride[3], x, y, bps) : 0;
            if (d->host_depth != 8)
            {
                if (unorm_y > (unsigned)yuv_max) unorm_y = (unsigned)yuv_max;
                if (unorm_a > (unsigned)yuv_max) unorm_a = (unsigned)yuv_max;
            }
            if (d->host_depth == 32)
            {
                /* YuvDecode.cpp:195-279: GrayAlpha32 un-premultiplies in the INTEGER domain, before the table. */
                float* p = (float*)dst_row + x * channels;
                if (has_alpha && premultiplied && unorm_a < (unsigned)yuv_max)
                {
                    if (unorm_a == 0)
                    {
                        unorm_y = 0;
                    }
                    else
                    {
                        unorm_y = avif_oracle_unpremultiply_u16((uint16_t)unorm_y, (uint16_t)unorm_a, (uint16_t)yuv_max);
                    }
                }
                p[0] = pq_to_linear(job->table_y[unorm_y], (float)d->pq_peak_nits); /* PQ only, :214-221 */
                if (has_alpha) p[1] = job->table_alpha[unorm_a];
            }
            else
            {
                float Y = job->table_y[unorm_y];
                float A = 0.0f;
                if (has_alpha)
                {
                    A = job->table_alpha[unorm_a];
                    if (premultiplied && unorm_a < (unsigned)yuv_max)
                    {
                        if (unorm_a == 0)
                        {
                            Y = 0;
                        }
                        else
                        {
                            Y = avif_oracle_unpremultiply_f32(Y, A, 1.0f);
                        }
                    }
                }
                if (d->host_depth == 8)
                {
                    /* YuvDecode.cpp:55-124 */
                    uint8_t* p = dst_row + x * channels;
                    p[0] = (uint8_t)(0.5f + (Y * 255.0f));
                    if (has_alpha) p[1] = (uint8_t)unorm_a;
                }
                else
                {
                    /* YuvDecode.cpp:126-193 */
                    uint16_t* p = (uint16_t*)dst_row + x * channels;
                    p[0] = (uint16_t)(0.5f + (Y * 32768.0f));
                    if (has_alpha) p[1] = (uint16_t)(0.5f + (A * 32768.0f));
                }
            }
        }
    }
}

/* Rows of a planar-RGB image: ReadHeifImage.cpp:561-712 (8), 714-861 (16), 949-1178 (32). */
static void decode_planar_rgb_rows(const decode_job* job)
{
    const avifgpu_decode_desc* d = job->desc;
    const int has_alpha = d->alpha_state != AVIFGPU_ALPHA_NONE;
    const int premultiplied = d->alpha_state == AVIFGPU_ALPHA_PREMULTIPLIED;
    const int bps = d->bit_depth > 8 ? 2 : 1;
    const unsigned max_value = (unsigned)((1 << d->bit_depth) - 1);
    const int channels = has_alpha ? 4 : 3;
    int x, y, i;
    for (y = job->y_begin; y < job->y_end; ++y)
    {
        uint8_t* dst_row = job->rows + (int64_t)y * job->row_stride;
        for (x = 0; x < d->width; ++x)
        {
            unsigned c[3];
            unsigned a = has_alpha ? load_sample(job->src.data[3], job->src.stride[3], x, y, bps) : 0;
            for (i = 0; i < 3; ++i) c[i] = load_sample(job->src.data[i], job->src.stride[i], x, y, bps);
            if (d->host_depth == 16)
            {
                /* ReadHeifImage.cpp:789-792, 846-848: samples are masked, not clamped */
                for (i = 0; i < 3; ++i) c[i] &= max_value;
                a &= max_value;
            }
            else if (d->host_depth == 32)
            {
                /* ReadHeifImage.cpp:1034-1060 indexes the table with the raw sample; a value above max_value is
                 * out of bounds there (undefined).  This project DEFINES it as clamped to max_value. */
                for (i = 0; i < 3; ++i) if (c[i] > max_value) c[i] = max_value;
                if (a > max_value) a = max_value;
            }
            if (has_alpha && premultiplied && a < max_value)
            {
                for (i = 0; i < 3; ++i)
                {
                    if (a == 0)
                    {
                        c[i] = 0;
                    }
                    else if (d->host_depth == 8)
                    {
                        c[i] = avif_oracle_unpremultiply_u8((uint8_t)c[i], (uint8_t)a);
                    }
                    else
                    {
                        c[i] = avif_oracle_unpremultiply_u16((uint16_t)c[i], (uint16_t)a, (uint16_t)max_value);
                    }
                }
            }
            if (d->host_depth == 8)
            {
                uint8_t* p = dst_row + x * channels;
                for (i = 0; i < 3; ++i) p[i] = (uint8_t)c[i];
                if (has_alpha) p[3] = (uint8_t)a;
            }
            else if (d->host_depth == 16)
            {
                uint16_t* p = (uint16_t*)dst_row + x * channels;
                for (i = 0; i < 3; ++i) p[i] = (uint16_t)c[i];
                if (has_alpha) p[3] = (uint16_t)a;
            }
            else
            {
                float* p = (float*)dst_row + x * channels;
                const float rgb[3] = { job->table_unorm[c[0]], job->table_unorm[c[1]], job->table_unorm[c[2]] };
                apply_eotf_rgb(job, rgb, p);
                if (has_alpha) p[3] = job->table_unorm[a];
            }
        }
    }
}

static void* decode_thread(void* arg)
{
    decode_job* job = (decode_job*)arg;
    switch (job->desc->colorspace)
    {
    case AVIFGPU_COLORSPACE_YCBCR: decode_ycbcr_rows(job); break;
    case AVIFGPU_COLORSPACE_MONOCHROME: decode_gray_rows(job); break;
    default: decode_planar_rgb_rows(job); break;
    }
    job->status = AVIFGPU_OK;
    return NULL;
}

static int validate_decode_desc(const avifgpu_decode_desc* d, int* transfer)
{
    *transfer = AVIFGPU_TRANSFER_CLIP;
    if (d == NULL || d->struct_size != sizeof(avifgpu_decode_desc)) return fail(AVIFGPU_ERR_BAD_PARAM, "bad decode desc size");
    if (d->width < 0 || d->height < 0) return fail(AVIFGPU_ERR_BAD_PARAM, "negative image size");
    if (d->host_depth != 8 && d->host_depth != 16 && d->host_depth != 32) return fail(AVIFGPU_ERR_BAD_PARAM, "host depth must be 8, 16 or 32");
    if (d->bit_depth != 8 && d->bit_depth != 10 && d->bit_depth != 12 && d->bit_depth != 16) return fail(AVIFGPU_ERR_UNSUPPORTED, "The image has an unsupported bit depth, must be 8, 10, 12 or 16.");
    if (d->colorspace != AVIFGPU_COLORSPACE_YCBCR && d->colorspace != AVIFGPU_COLORSPACE_RGB && d->colorspace != AVIFGPU_COLORSPACE_MONOCHROME) return fail(AVIFGPU_ERR_UNSUPPORTED, "Unsupported image color space, expected RGB.");
    /* The 8-bit host variants read uint8 planes, the others uint16 planes (ReadHeifImage.cpp:127-131 vs 240-244). */
    if ((d->host_depth == 8) != (d->bit_depth == 8)) return fail(AVIFGPU_ERR_UNSUPPORTED, "host depth 8 pairs with 8-bit planes only");
    if (d->colorspace == AVIFGPU_COLORSPACE_YCBCR && d->chroma != AVIFGPU_CHROMA_420 && d->chroma != AVIFGPU_CHROMA_422 && d->chroma != AVIFGPU_CHROMA_444) return fail(AVIFGPU_ERR_BAD_PARAM, "bad chroma");
    if (d->host_depth == 32)
    {
        if (!d->nclx.present) return fail(AVIFGPU_ERR_UNSUPPORTED, "The nclxProfile is null."); /* ReadHeifImage.cpp:870-873, 956-959 */
        if (!transfer_from_nclx(d->nclx.transfer_characteristics, transfer)) return fail(AVIFGPU_ERR_UNSUPPORTED, "Unsupported NCLX transfer characteristic.");
        if (d->colorspace == AVIFGPU_COLORSPACE_MONOCHROME && *transfer != AVIFGPU_TRANSFER_PQ) return fail(AVIFGPU_ERR_UNSUPPORTED, "Unsupported color transfer function.");
    }
    return AVIFGPU_OK;
}

int avif_oracle_decode_image_mt(const avifgpu_decode_desc* desc, const avifgpu_planes* src, void* host_rows,
                                int64_t row_stride, int32_t threads)
{
    int bounds[AVIF_ORACLE_MAX_THREADS + 2];
    decode_job jobs[AVIF_ORACLE_MAX_THREADS];
    pthread_t handles[AVIF_ORACLE_MAX_THREADS];
    decode_job base;
    int transfer, blocks, i;
    float k[3];
    float* tables = NULL;
    const int status = validate_decode_desc(desc, &transfer);
    if (status != AVIFGPU_OK) return status;

    memset(&base, 0, sizeof(base));
    base.desc = desc;
    base.src = *src;
    base.rows = (uint8_t*)host_rows;
    base.row_stride = row_stride;
    base.transfer = transfer;
    {
        const size_t count = (size_t)1 << desc->bit_depth;
        tables = (float*)malloc(4 * count * sizeof(float));
        if (tables == NULL) return fail(AVIFGPU_ERR_OOM, "out of memory");
        if (desc->colorspace == AVIFGPU_COLORSPACE_RGB)
        {
            /* ReadHeifImage.cpp:402-415 */
            size_t j;
            const float max_value = (float)(count - 1);
            for (j = 0; j < count; ++j) tables[j] = (float)j / max_value;
            base.table_unorm = tables;
        }
        else
        {
            const int mono = desc->colorspace == AVIFGPU_COLORSPACE_MONOCHROME;
            const int has_alpha = desc->alpha_state != AVIFGPU_ALPHA_NONE;
            const int st = build_tables(&desc->nclx, desc->bit_depth, mono, has_alpha, tables, tables + count, tables + 2 * count);
            if (st != AVIFGPU_OK) { free(tables); return st; }
            base.table_y = tables;
            base.table_uv = tables + count;
            base.table_alpha = tables + 2 * count;
        }
    }
    avif_oracle_get_yuv_coefficients(&desc->nclx, k);
    base.kr = k[0]; base.kg = k[1]; base.kb = k[2];
    if (desc->host_depth == 32 && transfer == AVIFGPU_TRANSFER_HLG && desc->hlg_apply_ootf)
    {
        const int st = avif_oracle_get_hlg_luma_coefficients(desc->nclx.color_primaries, base.hlg_luma);
        if (st != AVIFGPU_OK) { free(tables); return st; }
    }

    if (threads < 1) threads = 1;
    if (threads > AVIF_ORACLE_MAX_THREADS) threads = AVIF_ORACLE_MAX_THREADS;
    blocks = split_rows(desc->height, threads, bounds);
    for (i = 0; i < blocks; ++i)
    {
        jobs[i] = base;
        jobs[i].y_begin = bounds[i];
        jobs[i].y_end = bounds[i + 1];
    }
    if (blocks == 1)
    {
        decode_thread(&jobs[0]);
    }
    else
    {
        for (i = 0; i < blocks; ++i) pthread_create(&handles[i], NULL, decode_thread, &jobs[i]);
        for (i = 0; i < blocks; ++i) pthread_join(handles[i], NULL);
    }
    free(tables);
    return AVIFGPU_OK;
}

int avif_oracle_decode_image(const avifgpu_decode_desc* desc, const avifgpu_planes* src, void* host_rows,
                             int64_t row_stride)
{
    return avif_oracle_decode_image_mt(desc, src, host_rows, row_stride, 1);
}
