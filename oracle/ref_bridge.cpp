/*
 * oracle/ref_bridge.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * extern "C" doorway into the UNMODIFIED reference translation units, which oracle/Makefile compiles where
 * they lie under /root/reference/src/common:
 *     ColorTransfer.cpp  PremultipliedAlpha.cpp  YuvDecode.cpp  YuvLookupTables.cpp  YUVCoefficiants.cpp
 *     WriteHeifImage.cpp  ReadHeifImage.cpp
 * against the header shim in oracle/shim (mock Photoshop host + mock heif_image).  The result,
 * oracle/_ref/libavifref.so, is "the reference itself run here": it pins the C restatement in
 * oracle/avif_oracle.c and serves as the CPU baseline (bench.py cpu_baseline.kind == "reference").
 *
 * Everything in this file is glue written for this project: it builds a mock FormatRecord whose advanceState
 * feeds / drains rows from plain buffers, wraps caller planes in a mock heif_image, and calls the reference's
 * own entry points (WriteHeifImage.h:29-63, ReadHeifImage.h:27-63, YUVDecode.h, ColorTransfer.h,
 * PremultipliedAlpha.h, YUVCoefficiants.h, YUVLookupTables.h).
 */
#include "../include/avifgpu.h"

#include "AvifFormat.h"
#include "AlphaState.h"
#include "ColorTransfer.h"
#include "PremultipliedAlpha.h"
#include "ReadHeifImage.h"
#include "WriteHeifImage.h"
#include "YUVCoefficiants.h"
#include "YUVLookupTables.h"
#include "OSErrException.h"
#include "LibHeifException.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#define AVIFREF_EXPORT extern "C" __attribute__((visibility("default")))

namespace
{
    thread_local std::string g_lastError;

    // ---- mock host state (one per thread so row blocks can be converted on several threads) -------------

    struct Session
    {
        FormatRecord record{};
        BufferProcs bufferProcs{};
        // encode: rows come from here; decode: rows go there.
        const uint8_t* sourceRows = nullptr;
        uint8_t* destRows = nullptr;
        int64_t rowStride = 0;
        int64_t rowPayloadBytes = 0;
        int64_t advanceCalls = 0;
    };

    thread_local Session* g_session = nullptr;

    OSErr MockAllocateBuffer(int32 size, BufferID* bufferID)
    {
        void* memory = std::malloc(static_cast<size_t>(size > 0 ? size : 1));
        if (memory == nullptr)
        {
            return memFullErr;
        }
        *bufferID = reinterpret_cast<BufferID>(memory);
        return noErr;
    }

    Ptr MockLockBuffer(BufferID bufferID, Boolean) { return reinterpret_cast<Ptr>(bufferID); }
    void MockUnlockBuffer(BufferID) {}
    void MockFreeBuffer(BufferID bufferID) { std::free(reinterpret_cast<void*>(bufferID)); }
    Boolean MockAbort() { return 0; }
    void MockProgress(int32, int32) {}

    OSErr MockAdvanceState()
    {
        Session* s = g_session;
        FormatRecord& r = s->record;
        s->advanceCalls++;
        const int32 top = r.theRect32.top;
        const int32 bottom = r.theRect32.bottom;
        for (int32 y = top; y < bottom; ++y)
        {
            uint8_t* hostRow = static_cast<uint8_t*>(r.data) + static_cast<int64_t>(y - top) * r.rowBytes;
            if (s->sourceRows != nullptr)
            {
                std::memcpy(hostRow, s->sourceRows + static_cast<int64_t>(y) * s->rowStride,
                            static_cast<size_t>(s->rowPayloadBytes));
            }
            else if (s->destRows != nullptr)
            {
                std::memcpy(s->destRows + static_cast<int64_t>(y) * s->rowStride, hostRow,
                            static_cast<size_t>(s->rowPayloadBytes));
            }
        }
        return noErr;
    }

    int16 ImageModeFor(int channels, int depth)
    {
        const bool gray = channels <= 2;
        switch (depth)
        {
        case 8: return gray ? plugInModeGrayScale : plugInModeRGBColor;
        case 16: return gray ? plugInModeGray16 : plugInModeRGB48;
        default: return gray ? plugInModeGray32 : plugInModeRGB96;
        }
    }

    void InitSession(Session& s, int width, int height, int channels, int depth)
    {
        s.bufferProcs.allocateProc = MockAllocateBuffer;
        s.bufferProcs.lockProc = MockLockBuffer;
        s.bufferProcs.unlockProc = MockUnlockBuffer;
        s.bufferProcs.freeProc = MockFreeBuffer;
        FormatRecord& r = s.record;
        r.planes = static_cast<int16>(channels);
        r.depth = static_cast<int16>(depth);
        r.imageMode = ImageModeFor(channels, depth);
        r.imageSize32.h = width;
        r.imageSize32.v = height;
        r.imageSize.h = static_cast<int16>(std::min(width, 32767));
        r.imageSize.v = static_cast<int16>(std::min(height, 32767));
        r.HostSupports32BitCoordinates = 1;
        r.PluginUsing32BitCoordinates = 1;
        r.advanceState = MockAdvanceState;
        r.abortProc = MockAbort;
        r.progressProc = MockProgress;
        r.bufferProcs = &s.bufferProcs;
        r.transparencyPlane = static_cast<int16>((channels == 2 || channels == 4) ? channels - 1 : -1);
    }

    heif_color_profile_nclx ToHeifNclx(const avifgpu_nclx& n)
    {
        heif_color_profile_nclx out{};
        out.version = 1;
        out.color_primaries = static_cast<heif_color_primaries>(n.color_primaries);
        out.transfer_characteristics = static_cast<heif_transfer_characteristics>(n.transfer_characteristics);
        out.matrix_coefficients = static_cast<heif_matrix_coefficients>(n.matrix_coefficients);
        out.full_range_flag = static_cast<uint8_t>(n.full_range_flag != 0);
        return out;
    }

    template <typename F>
    int Guarded(F&& body)
    {
        try
        {
            body();
            return AVIFGPU_OK;
        }
        catch (const std::bad_alloc&)
        {
            g_lastError = "std::bad_alloc";
            return AVIFGPU_ERR_OOM;
        }
        catch (const OSErrException& e)
        {
            g_lastError = "OSErrException " + std::to_string(e.GetErrorCode());
            return e.GetErrorCode() == userCanceledErr ? AVIFGPU_ERR_CANCELED : AVIFGPU_ERR_BAD_PARAM;
        }
        catch (const LibHeifException& e)
        {
            g_lastError = std::string("LibHeifException ") + e.what();
            return AVIFGPU_ERR_UNSUPPORTED;
        }
        catch (const std::exception& e)
        {
            g_lastError = e.what();
            return AVIFGPU_ERR_UNSUPPORTED;
        }
    }

    SaveUIOptions ToSaveOptions(const avifgpu_encode_desc& d)
    {
        SaveUIOptions o{};
        o.quality = 85;
        o.chromaSubsampling = ChromaSubsampling::Yuv422;
        o.compressionSpeed = CompressionSpeed::Default;
        switch (d.image_bit_depth)
        {
        case 8: o.imageBitDepth = ImageBitDepth::Eight; break;
        case 10: o.imageBitDepth = ImageBitDepth::Ten; break;
        case 12: o.imageBitDepth = ImageBitDepth::Twelve; break;
        default: o.imageBitDepth = static_cast<ImageBitDepth>(99); break; // -> formatCannotRead in the reference
        }
        o.hdrTransferFunction = static_cast<ColorTransferFunction>(d.transfer);
        o.pq.nominalPeakBrightness = d.pq_peak_nits;
        o.keepColorProfile = true; // no ICC conversion on this path
        o.premultipliedAlpha = d.alpha_state == AVIFGPU_ALPHA_PREMULTIPLIED;
        return o;
    }

    LoadUIOptions ToLoadOptions(const avifgpu_decode_desc& d)
    {
        LoadUIOptions o{};
        o.format = LoadOptionsHDRFormat::Unknown;
        o.hlg.applyOOTF = d.hlg_apply_ootf != 0;
        o.hlg.displayGamma = d.hlg_display_gamma;
        o.hlg.nominalPeakBrightness = d.hlg_peak_nits;
        o.pq.nominalPeakBrightness = d.pq_peak_nits;
        return o;
    }

    void CopyPlaneOut(const heif_image* image, heif_channel channel, void* dst, int64_t dstStride,
                      int64_t payloadBytes, int rows)
    {
        int stride = 0;
        const uint8_t* src = heif_image_get_plane_readonly(image, channel, &stride);
        if (src == nullptr || dst == nullptr)
        {
            throw std::runtime_error("ref_bridge: missing plane");
        }
        for (int y = 0; y < rows; ++y)
        {
            std::memcpy(static_cast<uint8_t*>(dst) + y * dstStride, src + static_cast<int64_t>(y) * stride,
                        static_cast<size_t>(payloadBytes));
        }
    }

    void CopyPlaneIn(heif_image* image, heif_channel channel, const void* src, int64_t srcStride,
                     int64_t payloadBytes, int rows)
    {
        int stride = 0;
        uint8_t* dst = heif_image_get_plane(image, channel, &stride);
        if (src == nullptr || dst == nullptr)
        {
            throw std::runtime_error("ref_bridge: missing plane");
        }
        for (int y = 0; y < rows; ++y)
        {
            std::memcpy(dst + static_cast<int64_t>(y) * stride, static_cast<const uint8_t*>(src) + y * srcStride,
                        static_cast<size_t>(payloadBytes));
        }
    }

    // One image (or one row block presented as an image) through the reference's encode row shuttle.
    void EncodeImage(const avifgpu_encode_desc& d, const void* hostRows, int64_t rowStride, const avifgpu_planes& dst)
    {
        if (d.layout != AVIFGPU_LAYOUT_REFERENCE)
        {
            throw std::runtime_error("ref_bridge: the reference only produces AVIFGPU_LAYOUT_REFERENCE");
        }
        if (d.host_depth == 16 && d.host_channels <= 2 && d.gray16_curve != AVIFGPU_GRAY16_LUT)
        {
            throw std::runtime_error("ref_bridge: gray16 SMPTE428 is not a reference path");
        }

        Session session;
        InitSession(session, d.width, d.height, d.host_channels, d.host_depth);
        FormatRecord& r = session.record;

        // Write.cpp:279-299
        r.planeBytes = static_cast<int16>((r.depth + 7) / 8);
        r.loPlane = 0;
        r.hiPlane = static_cast<int16>(r.planes - 1);
        r.colBytes = static_cast<int16>(r.planes * r.planeBytes);
        r.rowBytes = static_cast<int32>(static_cast<int64_t>(d.width) * r.colBytes);
        std::vector<uint8_t> rowBuffer(static_cast<size_t>(std::max<int32>(r.rowBytes, 1)));
        r.data = rowBuffer.data();

        session.sourceRows = static_cast<const uint8_t*>(hostRows);
        session.rowStride = rowStride;
        session.rowPayloadBytes = r.rowBytes;

        const AlphaState alphaState = static_cast<AlphaState>(d.alpha_state);
        const VPoint imageSize{ d.height, d.width };
        const SaveUIOptions options = ToSaveOptions(d);
        const bool gray = d.host_channels <= 2;

        g_session = &session;
        ScopedHeifImage image;
        // Write.cpp:303-336
        if (gray)
        {
            switch (d.host_depth)
            {
            case 8: image = CreateHeifImageGrayEightBit(&r, alphaState, imageSize, options); break;
            case 16: image = CreateHeifImageGraySixteenBit(&r, alphaState, imageSize, options); break;
            case 32: image = CreateHeifImageGrayThirtyTwoBit(&r, alphaState, imageSize, options); break;
            default: throw OSErrException(formatBadParameters);
            }
        }
        else
        {
            switch (d.host_depth)
            {
            case 8: image = CreateHeifImageRGBEightBit(&r, alphaState, imageSize, options); break;
            case 16: image = CreateHeifImageRGBSixteenBit(&r, alphaState, imageSize, options); break;
            case 32: image = CreateHeifImageRGBThirtyTwoBit(&r, alphaState, imageSize, options); break;
            default: throw OSErrException(formatBadParameters);
            }
        }
        g_session = nullptr;

        const int bytesPerSample = d.image_bit_depth > 8 ? 2 : 1;
        if (gray)
        {
            CopyPlaneOut(image.get(), heif_channel_Y, dst.data[0], dst.stride[0],
                         static_cast<int64_t>(d.width) * bytesPerSample, d.height);
            if (alphaState != AlphaState::None)
            {
                CopyPlaneOut(image.get(), heif_channel_Alpha, dst.data[3], dst.stride[3],
                             static_cast<int64_t>(d.width) * bytesPerSample, d.height);
            }
        }
        else
        {
            CopyPlaneOut(image.get(), heif_channel_interleaved, dst.data[0], dst.stride[0],
                         static_cast<int64_t>(d.width) * d.host_channels * bytesPerSample, d.height);
        }
    }

    void DecodeImage(const avifgpu_decode_desc& d, const avifgpu_planes& src, void* hostRows, int64_t rowStride)
    {
        const bool hasAlpha = d.alpha_state != AVIFGPU_ALPHA_NONE;
        const int bytesPerSample = d.bit_depth > 8 ? 2 : 1;
        heif_colorspace colorspace = static_cast<heif_colorspace>(d.colorspace);
        heif_chroma chroma = static_cast<heif_chroma>(d.chroma);
        int channels;
        if (colorspace == heif_colorspace_monochrome)
        {
            chroma = heif_chroma_monochrome;
            channels = hasAlpha ? 2 : 1;
        }
        else
        {
            channels = hasAlpha ? 4 : 3;
            if (colorspace == heif_colorspace_RGB)
            {
                chroma = heif_chroma_444;
            }
        }

        heif_image* raw = nullptr;
        LibHeifException::ThrowIfError(heif_image_create(d.width, d.height, colorspace, chroma, &raw));
        ScopedHeifImage image(raw);

        const int cw = (chroma == heif_chroma_420 || chroma == heif_chroma_422) ? (d.width + 1) / 2 : d.width;
        const int ch = (chroma == heif_chroma_420) ? (d.height + 1) / 2 : d.height;

        auto addPlane = [&](heif_channel channel, int index, int w, int h)
        {
            LibHeifException::ThrowIfError(heif_image_add_plane(image.get(), channel, w, h, d.bit_depth));
            CopyPlaneIn(image.get(), channel, src.data[index], src.stride[index],
                        static_cast<int64_t>(w) * bytesPerSample, h);
        };

        if (colorspace == heif_colorspace_YCbCr)
        {
            addPlane(heif_channel_Y, 0, d.width, d.height);
            addPlane(heif_channel_Cb, 1, cw, ch);
            addPlane(heif_channel_Cr, 2, cw, ch);
        }
        else if (colorspace == heif_colorspace_RGB)
        {
            addPlane(heif_channel_R, 0, d.width, d.height);
            addPlane(heif_channel_G, 1, d.width, d.height);
            addPlane(heif_channel_B, 2, d.width, d.height);
        }
        else
        {
            addPlane(heif_channel_Y, 0, d.width, d.height);
        }
        if (hasAlpha)
        {
            addPlane(heif_channel_Alpha, 3, d.width, d.height);
        }

        Session session;
        InitSession(session, d.width, d.height, channels, d.host_depth);
        FormatRecord& r = session.record;
        session.destRows = static_cast<uint8_t*>(hostRows);
        session.rowStride = rowStride;
        session.rowPayloadBytes = static_cast<int64_t>(d.width) * channels * ((d.host_depth + 7) / 8);

        const heif_color_profile_nclx nclx = ToHeifNclx(d.nclx);
        const heif_color_profile_nclx* nclxPtr = d.nclx.present ? &nclx : nullptr;
        const AlphaState alphaState = static_cast<AlphaState>(d.alpha_state);
        const LoadUIOptions loadOptions = ToLoadOptions(d);
        const bool gray = colorspace == heif_colorspace_monochrome;

        g_session = &session;
        // Read.cpp:587-630
        if (gray)
        {
            switch (d.host_depth)
            {
            case 8: ReadHeifImageGrayEightBit(image.get(), alphaState, nclxPtr, &r); break;
            case 16: ReadHeifImageGraySixteenBit(image.get(), alphaState, nclxPtr, &r); break;
            case 32: ReadHeifImageGrayThirtyTwoBit(image.get(), alphaState, nclxPtr, loadOptions, &r); break;
            default: throw OSErrException(formatBadParameters);
            }
        }
        else
        {
            switch (d.host_depth)
            {
            case 8: ReadHeifImageRGBEightBit(image.get(), alphaState, nclxPtr, &r); break;
            case 16: ReadHeifImageRGBSixteenBit(image.get(), alphaState, nclxPtr, &r); break;
            case 32: ReadHeifImageRGBThirtyTwoBit(image.get(), alphaState, nclxPtr, loadOptions, &r); break;
            default: throw OSErrException(formatBadParameters);
            }
        }
        g_session = nullptr;
    }

    // Splits [0, height) into `parts` row blocks whose boundaries are even (4:2:0 row pairs stay together).
    std::vector<int> EvenRowSplits(int height, int parts)
    {
        std::vector<int> bounds;
        bounds.push_back(0);
        for (int i = 1; i < parts; ++i)
        {
            int b = static_cast<int>((static_cast<int64_t>(height) * i) / parts);
            b &= ~1;
            if (b > bounds.back() && b < height)
            {
                bounds.push_back(b);
            }
        }
        bounds.push_back(height);
        return bounds;
    }
}

AVIFREF_EXPORT const char* avifref_last_error(void) { return g_lastError.c_str(); }

AVIFREF_EXPORT const char* avifref_libm_version(void)
{
#if defined(__GLIBC__)
    static char text[64];
    std::snprintf(text, sizeof(text), "glibc %d.%d", __GLIBC__, __GLIBC_MINOR__);
    return text;
#else
    return "unknown libm";
#endif
}

// ---- scalar pixel math (ColorTransfer.cpp, PremultipliedAlpha.cpp) --------------------------------------

AVIFREF_EXPORT int avifref_transfer_f32(int32_t function, float param, const float* in, float* out, size_t n)
{
    switch (function)
    {
    case AVIFGPU_FN_LINEAR_TO_PQ: for (size_t i = 0; i < n; ++i) out[i] = LinearToPQ(in[i], param); break;
    case AVIFGPU_FN_PQ_TO_LINEAR: for (size_t i = 0; i < n; ++i) out[i] = PQToLinear(in[i], param); break;
    case AVIFGPU_FN_LINEAR_TO_SMPTE428: for (size_t i = 0; i < n; ++i) out[i] = LinearToSMPTE428(in[i]); break;
    case AVIFGPU_FN_SMPTE428_TO_LINEAR: for (size_t i = 0; i < n; ++i) out[i] = SMPTE428ToLinear(in[i]); break;
    case AVIFGPU_FN_HLG_TO_LINEAR: for (size_t i = 0; i < n; ++i) out[i] = HLGToLinear(in[i]); break;
    case AVIFGPU_FN_LINEAR_TO_HLG: for (size_t i = 0; i < n; ++i) out[i] = LinearToHLG(in[i]); break;
    case AVIFGPU_FN_POWF: for (size_t i = 0; i < n; ++i) out[i] = powf(in[i], param); break;
    case AVIFGPU_FN_EXPF: for (size_t i = 0; i < n; ++i) out[i] = expf(in[i]); break;
    case AVIFGPU_FN_LOGF: for (size_t i = 0; i < n; ++i) out[i] = logf(in[i]); break;
    default: g_lastError = "unknown function"; return AVIFGPU_ERR_BAD_PARAM;
    }
    return AVIFGPU_OK;
}

AVIFREF_EXPORT int avifref_hlg_ootf(float* rgb, size_t pixels, int32_t colorPrimaries, float displayGamma, float peakNits)
{
    return Guarded([&]
    {
        const HLGLumaCoefficiants luma = GetHLGLumaCoefficients(static_cast<heif_color_primaries>(colorPrimaries));
        for (size_t i = 0; i < pixels; ++i)
        {
            ApplyHLGOOTF(rgb + 3 * i, luma, displayGamma, peakNits);
        }
    });
}

AVIFREF_EXPORT int avifref_hlg_inverse_ootf(float* rgb, size_t pixels, int32_t colorPrimaries, float displayGamma, float peakNits)
{
    return Guarded([&]
    {
        const HLGLumaCoefficiants luma = GetHLGLumaCoefficients(static_cast<heif_color_primaries>(colorPrimaries));
        for (size_t i = 0; i < pixels; ++i)
        {
            ApplyInverseHLGOOTF(rgb + 3 * i, luma, displayGamma, peakNits);
        }
    });
}

AVIFREF_EXPORT uint8_t avifref_premultiply_u8(uint8_t color, uint8_t alpha) { return PremultiplyColor(color, alpha); }
AVIFREF_EXPORT uint16_t avifref_premultiply_u16(uint16_t color, uint16_t alpha, uint16_t maxValue) { return PremultiplyColor(color, alpha, maxValue); }
AVIFREF_EXPORT float avifref_premultiply_f32(float color, float alpha, float maxValue) { return PremultiplyColor(color, alpha, maxValue); }
AVIFREF_EXPORT uint8_t avifref_unpremultiply_u8(uint8_t color, uint8_t alpha) { return UnpremultiplyColor(color, alpha); }
AVIFREF_EXPORT uint16_t avifref_unpremultiply_u16(uint16_t color, uint16_t alpha, uint16_t maxValue) { return UnpremultiplyColor(color, alpha, maxValue); }
AVIFREF_EXPORT float avifref_unpremultiply_f32(float color, float alpha, float maxValue) { return UnpremultiplyColor(color, alpha, maxValue); }

// Exhaustive tables for the test-suite: out[c * count + a] for c, a in [0, count).
AVIFREF_EXPORT void avifref_premultiply_table_u16(uint16_t maxValue, int unpremultiply, uint16_t* out)
{
    const int count = static_cast<int>(maxValue) + 1;
    for (int c = 0; c < count; ++c)
    {
        for (int a = 0; a < count; ++a)
        {
            uint16_t v;
            if (unpremultiply)
            {
                v = a == 0 ? 0 : UnpremultiplyColor(static_cast<uint16_t>(c), static_cast<uint16_t>(a), maxValue);
            }
            else
            {
                v = PremultiplyColor(static_cast<uint16_t>(c), static_cast<uint16_t>(a), maxValue);
            }
            out[static_cast<size_t>(c) * count + a] = v;
        }
    }
}

AVIFREF_EXPORT void avifref_premultiply_table_u8(int unpremultiply, uint8_t* out)
{
    for (int c = 0; c < 256; ++c)
    {
        for (int a = 0; a < 256; ++a)
        {
            uint8_t v;
            if (unpremultiply)
            {
                v = a == 0 ? 0 : UnpremultiplyColor(static_cast<uint8_t>(c), static_cast<uint8_t>(a));
            }
            else
            {
                v = PremultiplyColor(static_cast<uint8_t>(c), static_cast<uint8_t>(a));
            }
            out[c * 256 + a] = v;
        }
    }
}

// ---- parameter derivation (YUVCoefficiants.cpp, YuvLookupTables.cpp, ColorTransfer.cpp:31-45) -----------

AVIFREF_EXPORT int avifref_get_yuv_coefficients(const avifgpu_nclx* nclx, float* out)
{
    return Guarded([&]
    {
        heif_color_profile_nclx h{};
        const heif_color_profile_nclx* p = nullptr;
        if (nclx != nullptr && nclx->present)
        {
            h = ToHeifNclx(*nclx);
            p = &h;
        }
        YUVCoefficiants c{};
        GetYUVCoefficiants(p, c);
        out[0] = c.kr;
        out[1] = c.kg;
        out[2] = c.kb;
    });
}

AVIFREF_EXPORT int avifref_get_hlg_luma_coefficients(int32_t primaries, float* out)
{
    return Guarded([&]
    {
        const HLGLumaCoefficiants c = GetHLGLumaCoefficients(static_cast<heif_color_primaries>(primaries));
        out[0] = c.red;
        out[1] = c.green;
        out[2] = c.blue;
    });
}

AVIFREF_EXPORT int avifref_build_yuv_tables(const avifgpu_nclx* nclx, int32_t bitDepth, int32_t monochrome,
                                            float* outY, float* outUV, float* outAlpha)
{
    return Guarded([&]
    {
        heif_color_profile_nclx h{};
        const heif_color_profile_nclx* p = nullptr;
        if (nclx != nullptr && nclx->present)
        {
            h = ToHeifNclx(*nclx);
            p = &h;
        }
        const YUVLookupTables tables(p, bitDepth, monochrome != 0, outAlpha != nullptr);
        const size_t count = static_cast<size_t>(1) << bitDepth;
        if (outY) std::memcpy(outY, tables.unormFloatTableY.get(), count * sizeof(float));
        if (outUV && !monochrome) std::memcpy(outUV, tables.unormFloatTableUV.get(), count * sizeof(float));
        if (outAlpha) std::memcpy(outAlpha, tables.unormFloatTableAlpha.get(), count * sizeof(float));
    });
}

// ---- whole-image drivers through the reference's own row shuttle ----------------------------------------

AVIFREF_EXPORT int avifref_encode_image(const avifgpu_encode_desc* desc, const void* hostRows, int64_t rowStride,
                                        const avifgpu_planes* dst)
{
    return Guarded([&] { EncodeImage(*desc, hostRows, rowStride, *dst); });
}

AVIFREF_EXPORT int avifref_decode_image(const avifgpu_decode_desc* desc, const avifgpu_planes* src, void* hostRows,
                                        int64_t rowStride)
{
    return Guarded([&] { DecodeImage(*desc, *src, hostRows, rowStride); });
}

// Row-block parallel variants: each thread presents its block to the (single-threaded) reference loops as an
// image of its own.  Rows are independent (4:2:0: row pairs), so the result equals the one-thread result.
AVIFREF_EXPORT int avifref_encode_image_mt(const avifgpu_encode_desc* desc, const void* hostRows, int64_t rowStride,
                                           const avifgpu_planes* dst, int32_t threads)
{
    if (threads <= 1)
    {
        return avifref_encode_image(desc, hostRows, rowStride, dst);
    }
    const std::vector<int> bounds = EvenRowSplits(desc->height, threads);
    std::vector<int> results(bounds.size() - 1, AVIFGPU_OK);
    std::vector<std::string> errors(bounds.size() - 1);
    std::vector<std::thread> pool;
    for (size_t i = 0; i + 1 < bounds.size(); ++i)
    {
        pool.emplace_back([&, i]
        {
            avifgpu_encode_desc d = *desc;
            d.height = bounds[i + 1] - bounds[i];
            avifgpu_planes p = *dst;
            for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
            {
                if (p.data[k]) p.data[k] = static_cast<uint8_t*>(p.data[k]) + static_cast<int64_t>(bounds[i]) * p.stride[k];
            }
            results[i] = avifref_encode_image(&d, static_cast<const uint8_t*>(hostRows) + static_cast<int64_t>(bounds[i]) * rowStride,
                                              rowStride, &p);
            errors[i] = g_lastError;
        });
    }
    for (auto& t : pool) t.join();
    for (size_t i = 0; i < results.size(); ++i)
    {
        if (results[i] != AVIFGPU_OK)
        {
            g_lastError = errors[i];
            return results[i];
        }
    }
    return AVIFGPU_OK;
}

AVIFREF_EXPORT int avifref_decode_image_mt(const avifgpu_decode_desc* desc, const avifgpu_planes* src, void* hostRows,
                                           int64_t rowStride, int32_t threads)
{
    if (threads <= 1)
    {
        return avifref_decode_image(desc, src, hostRows, rowStride);
    }
    const std::vector<int> bounds = EvenRowSplits(desc->height, threads);
    std::vector<int> results(bounds.size() - 1, AVIFGPU_OK);
    std::vector<std::string> errors(bounds.size() - 1);
    std::vector<std::thread> pool;
    const bool is420 = desc->colorspace == AVIFGPU_COLORSPACE_YCBCR && desc->chroma == AVIFGPU_CHROMA_420;
    for (size_t i = 0; i + 1 < bounds.size(); ++i)
    {
        pool.emplace_back([&, i]
        {
            avifgpu_decode_desc d = *desc;
            d.height = bounds[i + 1] - bounds[i];
            avifgpu_planes p = *src;
            for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
            {
                if (!p.data[k]) continue;
                const bool chromaPlane = is420 && (k == 1 || k == 2);
                const int64_t row = chromaPlane ? bounds[i] / 2 : bounds[i];
                p.data[k] = static_cast<uint8_t*>(p.data[k]) + row * p.stride[k];
            }
            results[i] = avifref_decode_image(&d, &p, static_cast<uint8_t*>(hostRows) + static_cast<int64_t>(bounds[i]) * rowStride,
                                              rowStride);
            errors[i] = g_lastError;
        });
    }
    for (auto& t : pool) t.join();
    for (size_t i = 0; i < results.size(); ++i)
    {
        if (results[i] != AVIFGPU_OK)
        {
            g_lastError = errors[i];
            return results[i];
        }
    }
    return AVIFGPU_OK;
}
