// avifgpu_api.cu -- the extern "C" surface declared in include/avifgpu.h: context, validation, PCIe staging for
// the host-pointer entry points, and dispatch to the kernels.  No CPU fallback anywhere: every entry point that
// converts pixels launches a CUDA kernel or fails.
#include "../../include/avifgpu.h"

#include "curve_tables.h"
#include "host_params.h"
#include "kernel_params.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

namespace avifgpu
{
int LaunchEncodeGeneric(const EncodeParams& params, int hostDepth, void* stream);
int LaunchDecodeGeneric(const DecodeParams& params, void* stream);
int LaunchEncodeFast(const EncodeParams& params, int hostDepth, void* stream);   // 0 = not applicable
int LaunchEncodeFastInteger(const EncodeParams& params, int hostDepth, void* stream); // 0 = not applicable
cudaError_t BuildGray16Lut(uint16_t* deviceLut, int smpte428, uint32_t maxCode, void* stream);
long long VerifyHlgDivisions(void* stream);
long long VerifyGreenDivision(const DecodeParams& params, void* stream);
int LaunchDecodeFast(const DecodeParams& params, void* stream);                  // 0 = not applicable
int LaunchDecodeFastInteger(const DecodeParams& params, void* stream);           // 0 = not applicable
int LaunchHlgOotf(int inverse, const float luma[3], float displayGamma, float peak, const float* in, float* out, size_t pixels, void* stream);

namespace
{
    thread_local int g_launchFailure = 0; // cudaError_t of the last failed launch on this thread, 0 = none
}

int ReportLaunchFailure(int cudaErrorCode)
{
    g_launchFailure = cudaErrorCode;
    return AVIFGPU_ERR_CUDA;
}

static int TakeLaunchFailure()
{
    const int code = g_launchFailure;
    g_launchFailure = 0;
    return code;
}

int LaunchEncode(const EncodeParams& params, int hostDepth, void* stream)
{
    int fast = LaunchEncodeFast(params, hostDepth, stream);
    if (fast != 0)
    {
        return fast;
    }
    fast = LaunchEncodeFastInteger(params, hostDepth, stream);
    if (fast != 0)
    {
        return fast;
    }
    return LaunchEncodeGeneric(params, hostDepth, stream);
}

int LaunchDecode(const DecodeParams& params, void* stream)
{
    int fast = LaunchDecodeFast(params, stream);
    if (fast != 0)
    {
        return fast;
    }
    fast = LaunchDecodeFastInteger(params, stream);
    if (fast != 0)
    {
        return fast;
    }
    return LaunchDecodeGeneric(params, stream);
}
} // namespace avifgpu

using namespace avifgpu;

namespace
{
    thread_local std::string g_creationError;

    // Number of row-block slices a host-pointer call is cut into so that the H2D copy of slice i+1, the kernel
    // of slice i and the D2H copy of slice i-1 overlap (three streams would not help: PCIe is full duplex, one
    // copy engine per direction).
    constexpr int kPipelineStreams = 2;
}

struct avifgpu_context
{
    int device = -1;
    cudaStream_t streams[kPipelineStreams] = {};
    cudaEvent_t sliceDone[kPipelineStreams] = {};
    std::string lastError;
    int64_t launches = 0;
    int smCount = 0;

    // Device staging for the host-pointer entry points (grow-only).
    struct Buffer
    {
        void* ptr = nullptr;
        size_t bytes = 0;
    };
    Buffer deviceRows[kPipelineStreams];
    Buffer devicePlanes[kPipelineStreams][AVIFGPU_MAX_PLANES];
    // Pinned bounce buffers for pageable caller memory (grow-only).
    Buffer pinnedRows[kPipelineStreams];
    Buffer pinnedPlanes[kPipelineStreams][AVIFGPU_MAX_PLANES];
    Buffer transferScratch[2];
    std::vector<CurveTable*> curveTables; // exact step tables, built per (curve, param, depth): explicitly or once they pay off
    struct PendingTable
    {
        int curve, param, depth;
        int64_t pixels; // converted with the exact kernel so far
    };
    std::vector<PendingTable> pendingTables;
    int64_t tableAutoBuildPixels = AVIFGPU_TABLE_AUTOBUILD_DEFAULT;
    struct Gray16Lut
    {
        int depth = 0;
        int smpte428 = 0;
        uint16_t* device = nullptr;
    };
    std::vector<Gray16Lut> gray16Luts;
    int hlgDivisionState = -1; // -1 not checked yet, 0 keep IEEE divisions, 1 fast divisions verified exact

    // HLG decode replaces two constant divisions by a 3-instruction form, but only after comparing it with the
    // IEEE division over every numerator the call sites can produce, on this device.
    int VerifiedHlgDivisions()
    {
        if (hlgDivisionState < 0)
        {
            const long long disagreements = VerifyHlgDivisions(streams[0]);
            hlgDivisionState = disagreements == 0 ? 1 : 0;
            launches += 1;
        }
        return hlgDivisionState;
    }

    // YuvDecode.cpp:308 divides by the per-image constant kg; the tuned decode kernels use a 3-instruction form after
    // it has been compared with the IEEE division for every (Cb, Cr) code pair of the configuration, on this device.
    struct GreenDivision
    {
        avifpix::InverseMatrix matrix;
        avifpix::RangeParams range;
        uint32_t maxCode;
        int state;
    };
    std::vector<GreenDivision> greenDivisions;
    int VerifiedGreenDivision(const DecodeParams& p)
    {
        if (p.colorspace != AVIFGPU_COLORSPACE_YCBCR || p.bitDepth > 12)
        {
            return 0;
        }
        for (const GreenDivision& g : greenDivisions)
        {
            if (std::memcmp(&g.matrix, &p.matrix, sizeof(g.matrix)) == 0 && std::memcmp(&g.range, &p.range, sizeof(g.range)) == 0 && g.maxCode == p.maxCode)
            {
                return g.state;
            }
        }
        GreenDivision g{};
        g.matrix = p.matrix;
        g.range = p.range;
        g.maxCode = p.maxCode;
        g.state = VerifyGreenDivision(p, streams[0]) == 0 ? 1 : 0;
        launches += 1;
        greenDivisions.push_back(g);
        return g.state;
    }

    // The 65536-entry code table of a Gray16 host configuration (built on the device on first use), or nullptr.
    const uint16_t* Gray16LutFor(const avifgpu_encode_desc& d)
    {
        if (d.host_depth != 16 || d.host_channels != 1 || d.layout != AVIFGPU_LAYOUT_REFERENCE || d.image_bit_depth <= 8)
        {
            return nullptr;
        }
        const int smpte428 = d.gray16_curve == AVIFGPU_GRAY16_SMPTE428 ? 1 : 0;
        for (const Gray16Lut& l : gray16Luts)
        {
            if (l.depth == d.image_bit_depth && l.smpte428 == smpte428)
            {
                return l.device;
            }
        }
        Gray16Lut lut;
        lut.depth = d.image_bit_depth;
        lut.smpte428 = smpte428;
        if (cudaMalloc(&lut.device, 65536 * sizeof(uint16_t)) != cudaSuccess)
        {
            cudaGetLastError();
            return nullptr;
        }
        if (BuildGray16Lut(lut.device, smpte428, (1u << d.image_bit_depth) - 1u, streams[0]) != cudaSuccess ||
            cudaStreamSynchronize(streams[0]) != cudaSuccess)
        {
            cudaGetLastError();
            cudaFree(lut.device);
            return nullptr;
        }
        launches += 1;
        gray16Luts.push_back(lut);
        return lut.device;
    }

    // The verified step table for a float-host encode description, or nullptr when the description does not use
    // one / the table could not be verified / building it has not paid off yet (then the generic exact kernel serves
    // the call).  `pixels` = the size of the call that asks; `force` = avifgpu_prepare_encode.
    CurveTable* CurveTableFor(const avifgpu_encode_desc& d, int64_t pixels, bool force)
    {
        if (d.host_depth != 32 || d.image_bit_depth > 12)
        {
            return nullptr;
        }
        int curve;
        int param = 0;
        if (d.transfer == AVIFGPU_TRANSFER_PQ)
        {
            curve = kCurveLinearToPQ;
            param = d.pq_peak_nits;
        }
        else if (d.transfer == AVIFGPU_TRANSFER_SMPTE428)
        {
            curve = kCurveLinearToSMPTE428;
        }
        else if (d.transfer == AVIFGPU_TRANSFER_HLG)
        {
            curve = kCurveLinearToHLG;
        }
        else
        {
            return nullptr;
        }
        for (CurveTable* t : curveTables)
        {
            if (t->curve == curve && t->param == param && t->depth == d.image_bit_depth)
            {
                return t;
            }
        }
        if (!force)
        {
            // ~40 ms of sweeps buy a ~7x faster kernel: worth it once a configuration has seen enough pixels
            PendingTable* pending = nullptr;
            for (PendingTable& candidate : pendingTables)
            {
                if (candidate.curve == curve && candidate.param == param && candidate.depth == d.image_bit_depth)
                {
                    pending = &candidate;
                }
            }
            if (pending == nullptr)
            {
                pendingTables.push_back(PendingTable{ curve, param, d.image_bit_depth, 0 });
                pending = &pendingTables.back();
            }
            pending->pixels += pixels;
            if (tableAutoBuildPixels < 0 || pending->pixels <= tableAutoBuildPixels)
            {
                return nullptr;
            }
        }
        CurveTable* t = new (std::nothrow) CurveTable();
        if (t == nullptr)
        {
            return nullptr;
        }
        BuildCurveTable(curve, param, d.image_bit_depth, streams[0], t);
        launches += t->stats.sweptInputs ? (t->stats.bandBitmapBytes ? 3 : 2) : 0; // sweep, (band bitmap,) verify
        curveTables.push_back(t);
        return t;
    }

    int Fail(int status, const std::string& message)
    {
        lastError = message;
        return status;
    }

    // A launcher returned a negative status: report the CUDA error it recorded (it has already cleared CUDA's own
    // slot), after draining the pipeline streams so that no copy of an earlier slice is still writing into caller memory
    // when the entry point returns.
    int LaunchFailed(int status, const char* what)
    {
        const int code = TakeLaunchFailure();
        for (cudaStream_t stream : streams)
        {
            if (stream)
            {
                cudaStreamSynchronize(stream);
            }
        }
        cudaGetLastError();
        lastError = std::string(what) + " failed: " + (code != 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : avifgpu_status_string(status));
        return status;
    }

    // Any failure in the middle of a host-pointer call: same draining, then the status.
    int Abandon(int status)
    {
        for (cudaStream_t stream : streams)
        {
            if (stream)
            {
                cudaStreamSynchronize(stream);
            }
        }
        cudaGetLastError();
        return status;
    }

    int Cuda(cudaError_t e, const char* what)
    {
        if (e == cudaSuccess)
        {
            return AVIFGPU_OK;
        }
        lastError = std::string(what) + ": " + cudaGetErrorString(e);
        return (e == cudaErrorMemoryAllocation) ? AVIFGPU_ERR_OOM : AVIFGPU_ERR_CUDA;
    }

    int EnsureDevice(Buffer& b, size_t bytes)
    {
        if (b.bytes >= bytes)
        {
            return AVIFGPU_OK;
        }
        if (b.ptr)
        {
            cudaFree(b.ptr);
            b.ptr = nullptr;
            b.bytes = 0;
        }
        const size_t rounded = ((bytes + (1u << 20) - 1) >> 20) << 20;
        const int status = Cuda(cudaMalloc(&b.ptr, rounded), "cudaMalloc");
        if (status == AVIFGPU_OK)
        {
            b.bytes = rounded;
        }
        return status;
    }

    int EnsurePinned(Buffer& b, size_t bytes)
    {
        if (b.bytes >= bytes)
        {
            return AVIFGPU_OK;
        }
        if (b.ptr)
        {
            cudaFreeHost(b.ptr);
            b.ptr = nullptr;
            b.bytes = 0;
        }
        const size_t rounded = ((bytes + (1u << 20) - 1) >> 20) << 20;
        const int status = Cuda(cudaHostAlloc(&b.ptr, rounded, cudaHostAllocDefault), "cudaHostAlloc");
        if (status == AVIFGPU_OK)
        {
            b.bytes = rounded;
        }
        return status;
    }
};

namespace
{
    bool IsPinned(const void* p)
    {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, p) != cudaSuccess)
        {
            cudaGetLastError();
            return false;
        }
        return attr.type == cudaMemoryTypeHost;
    }

    struct DeviceGuard
    {
        int previous = -1;
        explicit DeviceGuard(int device)
        {
            cudaGetDevice(&previous);
            if (previous != device)
            {
                cudaSetDevice(device);
            }
            else
            {
                previous = -1;
            }
        }
        ~DeviceGuard()
        {
            if (previous >= 0)
            {
                cudaSetDevice(previous);
            }
        }
    };

    int CheckBlock(avifgpu_context* ctx, int height, int ys, int y0, int nrows)
    {
        if (y0 < 0 || nrows < 0 || y0 > height || nrows > height - y0)
        {
            return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "row block outside the image");
        }
        if (ys && (y0 & 1))
        {
            return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "4:2:0 row blocks must start on an even row");
        }
        if (ys && (nrows & 1) && y0 + nrows != height)
        {
            return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "4:2:0 row blocks must have an even height unless they end the image");
        }
        return AVIFGPU_OK;
    }
}

// ---- context ------------------------------------------------------------------------------------------------

extern "C" {

AVIFGPU_EXPORT int avifgpu_api_version(void) { return AVIFGPU_API_VERSION; }

AVIFGPU_EXPORT const char* avifgpu_status_string(int status)
{
    switch (status)
    {
    case AVIFGPU_OK: return "ok";
    case AVIFGPU_ERR_BAD_PARAM: return "bad parameter";
    case AVIFGPU_ERR_UNSUPPORTED: return "unsupported";
    case AVIFGPU_ERR_NO_DEVICE: return "no usable CUDA device";
    case AVIFGPU_ERR_CUDA: return "CUDA error";
    case AVIFGPU_ERR_OOM: return "out of memory";
    case AVIFGPU_ERR_CANCELED: return "canceled";
    default: return "unknown status";
    }
}

AVIFGPU_EXPORT int avifgpu_create(int device_ordinal, avifgpu_context** out_ctx)
{
    if (out_ctx == nullptr)
    {
        g_creationError = "out_ctx is NULL";
        return AVIFGPU_ERR_BAD_PARAM;
    }
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
    {
        cudaGetLastError();
        g_creationError = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                          " (this library has no CPU fallback)";
        return AVIFGPU_ERR_NO_DEVICE;
    }
    if (device_ordinal < 0 || device_ordinal >= count)
    {
        g_creationError = "device ordinal out of range";
        return AVIFGPU_ERR_BAD_PARAM;
    }
    cudaDeviceProp prop{};
    e = cudaGetDeviceProperties(&prop, device_ordinal);
    if (e != cudaSuccess)
    {
        g_creationError = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return AVIFGPU_ERR_NO_DEVICE;
    }
    if (prop.major != 10)
    {
        char text[160];
        std::snprintf(text, sizeof(text), "device %d is sm_%d%d; this library ships sm_100a code only", device_ordinal, prop.major, prop.minor);
        g_creationError = text;
        return AVIFGPU_ERR_NO_DEVICE;
    }
    avifgpu_context* ctx = new (std::nothrow) avifgpu_context();
    if (ctx == nullptr)
    {
        g_creationError = "out of host memory";
        return AVIFGPU_ERR_OOM;
    }
    ctx->device = device_ordinal;
    ctx->smCount = prop.multiProcessorCount;
    DeviceGuard guard(device_ordinal);
    for (int i = 0; i < kPipelineStreams; ++i)
    {
        if (cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->sliceDone[i], cudaEventDisableTiming) != cudaSuccess)
        {
            g_creationError = std::string("stream/event creation failed: ") + cudaGetErrorString(cudaGetLastError());
            avifgpu_destroy(ctx);
            return AVIFGPU_ERR_CUDA;
        }
    }
    *out_ctx = ctx;
    return AVIFGPU_OK;
}

AVIFGPU_EXPORT void avifgpu_destroy(avifgpu_context* ctx)
{
    if (ctx == nullptr)
    {
        return;
    }
    DeviceGuard guard(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kPipelineStreams; ++i)
    {
        if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
        if (ctx->sliceDone[i]) cudaEventDestroy(ctx->sliceDone[i]);
        if (ctx->deviceRows[i].ptr) cudaFree(ctx->deviceRows[i].ptr);
        if (ctx->pinnedRows[i].ptr) cudaFreeHost(ctx->pinnedRows[i].ptr);
        for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
        {
            if (ctx->devicePlanes[i][k].ptr) cudaFree(ctx->devicePlanes[i][k].ptr);
            if (ctx->pinnedPlanes[i][k].ptr) cudaFreeHost(ctx->pinnedPlanes[i][k].ptr);
        }
    }
    for (auto& b : ctx->transferScratch)
    {
        if (b.ptr) cudaFree(b.ptr);
    }
    for (CurveTable* t : ctx->curveTables)
    {
        FreeCurveTable(t);
        delete t;
    }
    for (auto& l : ctx->gray16Luts)
    {
        cudaFree(l.device);
    }
    delete ctx;
}

AVIFGPU_EXPORT const char* avifgpu_last_error(const avifgpu_context* ctx)
{
    return ctx ? ctx->lastError.c_str() : g_creationError.c_str();
}

AVIFGPU_EXPORT int64_t avifgpu_launch_count(const avifgpu_context* ctx) { return ctx ? ctx->launches : 0; }

AVIFGPU_EXPORT int avifgpu_synchronize(avifgpu_context* ctx)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    DeviceGuard guard(ctx->device);
    return ctx->Cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
}

AVIFGPU_EXPORT int avifgpu_host_alloc(avifgpu_context* ctx, size_t bytes, void** out_ptr)
{
    if (ctx == nullptr || out_ptr == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    DeviceGuard guard(ctx->device);
    *out_ptr = nullptr;
    return ctx->Cuda(cudaHostAlloc(out_ptr, bytes ? bytes : 1, cudaHostAllocDefault), "cudaHostAlloc");
}

AVIFGPU_EXPORT int avifgpu_host_free(avifgpu_context* ctx, void* ptr)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    DeviceGuard guard(ctx->device);
    return ctx->Cuda(cudaFreeHost(ptr), "cudaFreeHost");
}

// ---- geometry and parameter derivation (no device needed) -----------------------------------------------------

AVIFGPU_EXPORT int avifgpu_encode_host_col_bytes(const avifgpu_encode_desc* desc)
{
    return ValidateEncodeDesc(desc, nullptr) == AVIFGPU_OK ? EncodeHostColBytes(*desc) : AVIFGPU_ERR_BAD_PARAM;
}

AVIFGPU_EXPORT int avifgpu_decode_host_col_bytes(const avifgpu_decode_desc* desc)
{
    int32_t transfer;
    return ValidateDecodeDesc(desc, &transfer, nullptr) == AVIFGPU_OK ? DecodeHostColBytes(*desc) : AVIFGPU_ERR_BAD_PARAM;
}

static int ReportGeometry(const PlaneGeometry& g, int32_t* w, int32_t* h, int32_t* b)
{
    if (w) *w = g.present ? g.widthSamples : 0;
    if (h) *h = g.present ? g.height : 0;
    if (b) *b = g.present ? g.bytesPerSample : 0;
    return g.present ? 1 : 0;
}

AVIFGPU_EXPORT int avifgpu_encode_plane_geometry(const avifgpu_encode_desc* desc, int index, int32_t* w, int32_t* h, int32_t* b)
{
    const int status = ValidateEncodeDesc(desc, nullptr);
    if (status != AVIFGPU_OK || index < 0 || index >= AVIFGPU_MAX_PLANES)
    {
        ReportGeometry(PlaneGeometry(), w, h, b);
        return status != AVIFGPU_OK ? status : AVIFGPU_ERR_BAD_PARAM;
    }
    return ReportGeometry(EncodePlaneGeometry(*desc, index), w, h, b);
}

AVIFGPU_EXPORT int avifgpu_decode_plane_geometry(const avifgpu_decode_desc* desc, int index, int32_t* w, int32_t* h, int32_t* b)
{
    int32_t transfer;
    const int status = ValidateDecodeDesc(desc, &transfer, nullptr);
    if (status != AVIFGPU_OK || index < 0 || index >= AVIFGPU_MAX_PLANES)
    {
        ReportGeometry(PlaneGeometry(), w, h, b);
        return status != AVIFGPU_OK ? status : AVIFGPU_ERR_BAD_PARAM;
    }
    return ReportGeometry(DecodePlaneGeometry(*desc, index), w, h, b);
}

AVIFGPU_EXPORT int avifgpu_get_yuv_coefficients(const avifgpu_nclx* nclx, float* out)
{
    if (out == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    GetYuvCoefficients(nclx, out);
    return AVIFGPU_OK;
}

AVIFGPU_EXPORT int avifgpu_get_hlg_luma_coefficients(int32_t color_primaries, float* out)
{
    if (out == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    return GetHlgLumaCoefficients(color_primaries, out) ? AVIFGPU_OK : AVIFGPU_ERR_UNSUPPORTED;
}

AVIFGPU_EXPORT int avifgpu_build_yuv_tables(const avifgpu_nclx* nclx, int32_t bit_depth, int32_t monochrome, float* out_y,
                                            float* out_uv, float* out_alpha)
{
    if (bit_depth != 8 && bit_depth != 10 && bit_depth != 12 && bit_depth != 16)
    {
        return AVIFGPU_ERR_UNSUPPORTED;
    }
    const avifpix::RangeParams range = MakeRangeParams(nclx, bit_depth, monochrome != 0);
    const uint32_t count = 1u << bit_depth;
    for (uint32_t i = 0; i < count; ++i)
    {
        if (out_y) out_y[i] = avifpix::UnormToFloatY(i, range);
        if (out_uv && !monochrome) out_uv[i] = avifpix::UnormToFloatUV(i, range);
        if (out_alpha) out_alpha[i] = avifpix::UnormToFloatPlain(i, range.maxChannelFloat);
    }
    return AVIFGPU_OK;
}

// ---- device-pointer entry points -------------------------------------------------------------------------------

AVIFGPU_EXPORT int avifgpu_encode_rows_device(avifgpu_context* ctx, const avifgpu_encode_desc* desc, const void* device_rows,
                                              int64_t row_stride_bytes, int32_t y0, int32_t nrows,
                                              const avifgpu_planes* device_dst, void* cuda_stream)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    std::string error;
    int status = ValidateEncodeDesc(desc, &error);
    if (status != AVIFGPU_OK)
    {
        return ctx->Fail(status, error);
    }
    if (device_dst == nullptr || (device_rows == nullptr && nrows > 0 && desc->width > 0))
    {
        return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "NULL buffer");
    }
    EncodeParams p;
    FillEncodeParams(*desc, &p);
    status = CheckBlock(ctx, desc->height, p.ys, y0, nrows);
    if (status != AVIFGPU_OK)
    {
        return status;
    }
    if (nrows == 0 || desc->width == 0)
    {
        return AVIFGPU_OK;
    }
    p.rows = device_rows;
    p.rowStride = row_stride_bytes;
    p.rowCount = nrows;
    for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
    {
        const PlaneGeometry g = EncodePlaneGeometry(*desc, k);
        if (!g.present)
        {
            continue;
        }
        if (device_dst->data[k] == nullptr)
        {
            return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "missing destination plane");
        }
        p.plane[k] = static_cast<uint8_t*>(device_dst->data[k]) + static_cast<int64_t>(y0 >> g.ys) * device_dst->stride[k];
        p.planeStride[k] = device_dst->stride[k];
    }
    DeviceGuard guard(ctx->device);
    p.smCount = ctx->smCount;
    if (CurveTable* table = ctx->CurveTableFor(*desc, static_cast<int64_t>(desc->width) * nrows, false))
    {
        p.curveTable = table->valid ? &table->view : nullptr;
    }
    p.gray16Lut = ctx->Gray16LutFor(*desc);
    const int launched = LaunchEncode(p, desc->host_depth, cuda_stream);
    if (launched < 0)
    {
        return ctx->LaunchFailed(launched, "encode kernel launch");
    }
    ctx->launches += launched;
    return AVIFGPU_OK;
}

AVIFGPU_EXPORT int avifgpu_decode_rows_device(avifgpu_context* ctx, const avifgpu_decode_desc* desc,
                                              const avifgpu_planes* device_src, int32_t y0, int32_t nrows, void* device_rows,
                                              int64_t row_stride_bytes, void* cuda_stream)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    std::string error;
    int32_t transfer;
    int status = ValidateDecodeDesc(desc, &transfer, &error);
    if (status != AVIFGPU_OK)
    {
        return ctx->Fail(status, error);
    }
    if (device_src == nullptr || (device_rows == nullptr && nrows > 0 && desc->width > 0))
    {
        return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "NULL buffer");
    }
    DecodeParams p;
    if (!FillDecodeParams(*desc, transfer, &p, &error))
    {
        return ctx->Fail(AVIFGPU_ERR_UNSUPPORTED, error);
    }
    if (y0 < 0 || nrows < 0 || y0 > desc->height || nrows > desc->height - y0)
    {
        return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "row block outside the image");
    }
    if (nrows == 0 || desc->width == 0)
    {
        return AVIFGPU_OK;
    }
    p.rows = device_rows;
    p.rowStride = row_stride_bytes;
    p.rowCount = nrows;
    p.yPhase = y0 & p.ys;
    p.smCount = ctx->smCount;
    DeviceGuard deviceGuardForTables(ctx->device);
    p.verifiedHlgDivisions = (desc->host_depth == 32 && transfer == AVIFGPU_TRANSFER_HLG) ? ctx->VerifiedHlgDivisions() : 0;
    p.verifiedGreenDivision = ctx->VerifiedGreenDivision(p);
    for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
    {
        const PlaneGeometry g = DecodePlaneGeometry(*desc, k);
        if (!g.present)
        {
            continue;
        }
        if (device_src->data[k] == nullptr)
        {
            return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "missing source plane");
        }
        p.plane[k] = static_cast<const uint8_t*>(device_src->data[k]) + static_cast<int64_t>(y0 >> g.ys) * device_src->stride[k];
        p.planeStride[k] = device_src->stride[k];
    }
    DeviceGuard guard(ctx->device);
    const int launched = LaunchDecode(p, cuda_stream);
    if (launched < 0)
    {
        return ctx->LaunchFailed(launched, "decode kernel launch");
    }
    ctx->launches += launched;
    return AVIFGPU_OK;
}

// ---- host-pointer entry points (PCIe inside) ---------------------------------------------------------------------

// Rows per pipeline slice: big enough to amortise launch + copy latency, small enough that two slices overlap.
static int SliceRows(int nrows, int64_t bytesPerRow)
{
    const int64_t target = 32ll << 20; // ~32 MiB of host rows per slice
    int64_t rows = bytesPerRow > 0 ? target / bytesPerRow : nrows;
    rows = std::max<int64_t>(rows, 2);
    rows &= ~1ll; // keep 4:2:0 row pairs together
    return static_cast<int>(std::min<int64_t>(rows, nrows));
}

AVIFGPU_EXPORT int avifgpu_encode_rows(avifgpu_context* ctx, const avifgpu_encode_desc* desc, const void* host_rows,
                                       int64_t row_stride_bytes, int32_t y0, int32_t nrows, const avifgpu_planes* dst)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    std::string error;
    int status = ValidateEncodeDesc(desc, &error);
    if (status != AVIFGPU_OK)
    {
        return ctx->Fail(status, error);
    }
    if (dst == nullptr || (host_rows == nullptr && nrows > 0 && desc->width > 0))
    {
        return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "NULL buffer");
    }
    EncodeParams base;
    FillEncodeParams(*desc, &base);
    status = CheckBlock(ctx, desc->height, base.ys, y0, nrows);
    if (status != AVIFGPU_OK)
    {
        return status;
    }
    if (nrows == 0 || desc->width == 0)
    {
        return AVIFGPU_OK;
    }
    PlaneGeometry geometry[AVIFGPU_MAX_PLANES];
    for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
    {
        geometry[k] = EncodePlaneGeometry(*desc, k);
        if (geometry[k].present && dst->data[k] == nullptr)
        {
            return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "missing destination plane");
        }
    }

    DeviceGuard guard(ctx->device);
    base.smCount = ctx->smCount;
    if (CurveTable* table = ctx->CurveTableFor(*desc, static_cast<int64_t>(desc->width) * nrows, false))
    {
        base.curveTable = table->valid ? &table->view : nullptr;
    }
    base.gray16Lut = ctx->Gray16LutFor(*desc);
    const int64_t rowPayload = static_cast<int64_t>(desc->width) * EncodeHostColBytes(*desc);
    const int64_t deviceRowStride = (rowPayload + 255) & ~255ll;
    const bool rowsPinned = IsPinned(host_rows);
    const int sliceRows = SliceRows(nrows, rowPayload);

    int slot = 0;
    for (int begin = 0; begin < nrows; begin += sliceRows, slot = (slot + 1) % kPipelineStreams)
    {
        const int rows = std::min(sliceRows, nrows - begin);
        cudaStream_t stream = ctx->streams[slot];
        // The slot's buffers are free once its previous slice has been copied back.
        if ((status = ctx->Cuda(cudaEventSynchronize(ctx->sliceDone[slot]), "cudaEventSynchronize")) != AVIFGPU_OK) return status;

        if ((status = ctx->EnsureDevice(ctx->deviceRows[slot], static_cast<size_t>(deviceRowStride) * rows)) != AVIFGPU_OK) return status;
        const uint8_t* source = static_cast<const uint8_t*>(host_rows) + static_cast<int64_t>(begin) * row_stride_bytes;
        int64_t sourceStride = row_stride_bytes;
        if (!rowsPinned)
        {
            // Pageable caller memory: bounce through pinned memory so the DMA is asynchronous and full speed.
            if ((status = ctx->EnsurePinned(ctx->pinnedRows[slot], static_cast<size_t>(rowPayload) * rows)) != AVIFGPU_OK) return status;
            uint8_t* bounce = static_cast<uint8_t*>(ctx->pinnedRows[slot].ptr);
            for (int r = 0; r < rows; ++r)
            {
                std::memcpy(bounce + static_cast<int64_t>(r) * rowPayload, source + static_cast<int64_t>(r) * row_stride_bytes,
                            static_cast<size_t>(rowPayload));
            }
            source = bounce;
            sourceStride = rowPayload;
        }
        if ((status = ctx->Cuda(cudaMemcpy2DAsync(ctx->deviceRows[slot].ptr, static_cast<size_t>(deviceRowStride), source,
                                                  static_cast<size_t>(sourceStride), static_cast<size_t>(rowPayload),
                                                  static_cast<size_t>(rows), cudaMemcpyHostToDevice, stream),
                                "H2D rows")) != AVIFGPU_OK) return status;

        EncodeParams p = base;
        p.rows = ctx->deviceRows[slot].ptr;
        p.rowStride = deviceRowStride;
        p.rowCount = rows;
        int64_t planeStride[AVIFGPU_MAX_PLANES] = {};
        int planeRows[AVIFGPU_MAX_PLANES] = {};
        int64_t planePayload[AVIFGPU_MAX_PLANES] = {};
        for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
        {
            const PlaneGeometry& g = geometry[k];
            if (!g.present)
            {
                continue;
            }
            planePayload[k] = static_cast<int64_t>(g.widthSamples) * g.bytesPerSample;
            planeStride[k] = (planePayload[k] + 255) & ~255ll;
            planeRows[k] = (rows + g.ys) >> g.ys;
            if ((status = ctx->EnsureDevice(ctx->devicePlanes[slot][k], static_cast<size_t>(planeStride[k]) * planeRows[k])) != AVIFGPU_OK) return status;
            p.plane[k] = ctx->devicePlanes[slot][k].ptr;
            p.planeStride[k] = planeStride[k];
        }
        const int launched = LaunchEncode(p, desc->host_depth, stream);
        if (launched < 0)
        {
            return ctx->LaunchFailed(launched, "encode kernel launch");
        }
        ctx->launches += launched;

        for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
        {
            const PlaneGeometry& g = geometry[k];
            if (!g.present)
            {
                continue;
            }
            uint8_t* target = static_cast<uint8_t*>(dst->data[k]) + static_cast<int64_t>((y0 + begin) >> g.ys) * dst->stride[k];
            if ((status = ctx->Cuda(cudaMemcpy2DAsync(target, static_cast<size_t>(dst->stride[k]), p.plane[k],
                                                      static_cast<size_t>(planeStride[k]), static_cast<size_t>(planePayload[k]),
                                                      static_cast<size_t>(planeRows[k]), cudaMemcpyDeviceToHost, stream),
                                    "D2H plane")) != AVIFGPU_OK) return status;
        }
        if ((status = ctx->Cuda(cudaEventRecord(ctx->sliceDone[slot], stream), "cudaEventRecord")) != AVIFGPU_OK) return status;
    }
    for (int i = 0; i < kPipelineStreams; ++i)
    {
        if ((status = ctx->Cuda(cudaStreamSynchronize(ctx->streams[i]), "cudaStreamSynchronize")) != AVIFGPU_OK) return status;
    }
    return AVIFGPU_OK;
}

AVIFGPU_EXPORT int avifgpu_decode_rows(avifgpu_context* ctx, const avifgpu_decode_desc* desc, const avifgpu_planes* src,
                                       int32_t y0, int32_t nrows, void* host_rows, int64_t row_stride_bytes)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    std::string error;
    int32_t transfer;
    int status = ValidateDecodeDesc(desc, &transfer, &error);
    if (status != AVIFGPU_OK)
    {
        return ctx->Fail(status, error);
    }
    if (src == nullptr || (host_rows == nullptr && nrows > 0 && desc->width > 0))
    {
        return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "NULL buffer");
    }
    DecodeParams base;
    if (!FillDecodeParams(*desc, transfer, &base, &error))
    {
        return ctx->Fail(AVIFGPU_ERR_UNSUPPORTED, error);
    }
    if (y0 < 0 || nrows < 0 || y0 > desc->height || nrows > desc->height - y0)
    {
        return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "row block outside the image");
    }
    if (nrows == 0 || desc->width == 0)
    {
        return AVIFGPU_OK;
    }
    PlaneGeometry geometry[AVIFGPU_MAX_PLANES];
    for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
    {
        geometry[k] = DecodePlaneGeometry(*desc, k);
        if (geometry[k].present && src->data[k] == nullptr)
        {
            return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "missing source plane");
        }
    }

    DeviceGuard guard(ctx->device);
    base.verifiedHlgDivisions = (desc->host_depth == 32 && transfer == AVIFGPU_TRANSFER_HLG) ? ctx->VerifiedHlgDivisions() : 0;
    base.verifiedGreenDivision = ctx->VerifiedGreenDivision(base);
    const int64_t rowPayload = static_cast<int64_t>(desc->width) * DecodeHostColBytes(*desc);
    const int64_t deviceRowStride = (rowPayload + 255) & ~255ll;
    const int sliceRows = SliceRows(nrows, rowPayload);
    const bool rowsPinned = IsPinned(host_rows);

    struct Pending
    {
        bool active = false;
        int begin = 0;
        int rows = 0;
    } pending[kPipelineStreams];

    auto drain = [&](int slot) -> int
    {
        // Copies a finished slice from the pinned bounce buffer to pageable caller rows.
        if (!pending[slot].active)
        {
            return AVIFGPU_OK;
        }
        int st = ctx->Cuda(cudaEventSynchronize(ctx->sliceDone[slot]), "cudaEventSynchronize");
        if (st != AVIFGPU_OK) return st;
        if (!rowsPinned)
        {
            const uint8_t* bounce = static_cast<const uint8_t*>(ctx->pinnedRows[slot].ptr);
            uint8_t* target = static_cast<uint8_t*>(host_rows) + static_cast<int64_t>(pending[slot].begin) * row_stride_bytes;
            for (int r = 0; r < pending[slot].rows; ++r)
            {
                std::memcpy(target + static_cast<int64_t>(r) * row_stride_bytes, bounce + static_cast<int64_t>(r) * rowPayload,
                            static_cast<size_t>(rowPayload));
            }
        }
        pending[slot].active = false;
        return AVIFGPU_OK;
    };

    int slot = 0;
    for (int begin = 0; begin < nrows; begin += sliceRows, slot = (slot + 1) % kPipelineStreams)
    {
        const int rows = std::min(sliceRows, nrows - begin);
        cudaStream_t stream = ctx->streams[slot];
        if ((status = drain(slot)) != AVIFGPU_OK) return status;
        if ((status = ctx->Cuda(cudaEventSynchronize(ctx->sliceDone[slot]), "cudaEventSynchronize")) != AVIFGPU_OK) return status;

        const int yFirst = y0 + begin;
        DecodeParams p = base;
        p.rowCount = rows;
        p.yPhase = yFirst & p.ys;
        p.smCount = ctx->smCount;
        for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
        {
            const PlaneGeometry& g = geometry[k];
            if (!g.present)
            {
                continue;
            }
            const int firstRow = yFirst >> g.ys;
            const int lastRow = (yFirst + rows - 1) >> g.ys;
            const int planeRows = lastRow - firstRow + 1;
            const int64_t payload = static_cast<int64_t>(g.widthSamples) * g.bytesPerSample;
            const int64_t stride = (payload + 255) & ~255ll;
            if ((status = ctx->EnsureDevice(ctx->devicePlanes[slot][k], static_cast<size_t>(stride) * planeRows)) != AVIFGPU_OK) return status;
            const uint8_t* source = static_cast<const uint8_t*>(src->data[k]) + static_cast<int64_t>(firstRow) * src->stride[k];
            int64_t sourceStride = src->stride[k];
            if (!IsPinned(source))
            {
                if ((status = ctx->EnsurePinned(ctx->pinnedPlanes[slot][k], static_cast<size_t>(payload) * planeRows)) != AVIFGPU_OK) return status;
                uint8_t* bounce = static_cast<uint8_t*>(ctx->pinnedPlanes[slot][k].ptr);
                for (int r = 0; r < planeRows; ++r)
                {
                    std::memcpy(bounce + static_cast<int64_t>(r) * payload, source + static_cast<int64_t>(r) * src->stride[k],
                                static_cast<size_t>(payload));
                }
                source = bounce;
                sourceStride = payload;
            }
            if ((status = ctx->Cuda(cudaMemcpy2DAsync(ctx->devicePlanes[slot][k].ptr, static_cast<size_t>(stride), source,
                                                      static_cast<size_t>(sourceStride), static_cast<size_t>(payload),
                                                      static_cast<size_t>(planeRows), cudaMemcpyHostToDevice, stream),
                                    "H2D plane")) != AVIFGPU_OK) return status;
            p.plane[k] = ctx->devicePlanes[slot][k].ptr;
            p.planeStride[k] = stride;
        }
        if ((status = ctx->EnsureDevice(ctx->deviceRows[slot], static_cast<size_t>(deviceRowStride) * rows)) != AVIFGPU_OK) return status;
        p.rows = ctx->deviceRows[slot].ptr;
        p.rowStride = deviceRowStride;
        const int launched = LaunchDecode(p, stream);
        if (launched < 0)
        {
            return ctx->LaunchFailed(launched, "decode kernel launch");
        }
        ctx->launches += launched;

        uint8_t* target = static_cast<uint8_t*>(host_rows) + static_cast<int64_t>(begin) * row_stride_bytes;
        int64_t targetStride = row_stride_bytes;
        if (!rowsPinned)
        {
            if ((status = ctx->EnsurePinned(ctx->pinnedRows[slot], static_cast<size_t>(rowPayload) * rows)) != AVIFGPU_OK) return status;
            target = static_cast<uint8_t*>(ctx->pinnedRows[slot].ptr);
            targetStride = rowPayload;
        }
        if ((status = ctx->Cuda(cudaMemcpy2DAsync(target, static_cast<size_t>(targetStride), ctx->deviceRows[slot].ptr,
                                                  static_cast<size_t>(deviceRowStride), static_cast<size_t>(rowPayload),
                                                  static_cast<size_t>(rows), cudaMemcpyDeviceToHost, stream),
                                "D2H rows")) != AVIFGPU_OK) return status;
        if ((status = ctx->Cuda(cudaEventRecord(ctx->sliceDone[slot], stream), "cudaEventRecord")) != AVIFGPU_OK) return status;
        pending[slot].active = true;
        pending[slot].begin = begin;
        pending[slot].rows = rows;
    }
    for (int i = 0; i < kPipelineStreams; ++i)
    {
        if ((status = drain(i)) != AVIFGPU_OK) return status;
        if ((status = ctx->Cuda(cudaStreamSynchronize(ctx->streams[i]), "cudaStreamSynchronize")) != AVIFGPU_OK) return status;
    }
    return AVIFGPU_OK;
}

// ---- preparation ------------------------------------------------------------------------------------------------

AVIFGPU_EXPORT int avifgpu_prepare_encode(avifgpu_context* ctx, const avifgpu_encode_desc* desc, avifgpu_curve_stats* out_stats)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    std::string error;
    const int status = ValidateEncodeDesc(desc, &error);
    if (status != AVIFGPU_OK)
    {
        return ctx->Fail(status, error);
    }
    DeviceGuard guard(ctx->device);
    CurveTable* table = ctx->CurveTableFor(*desc, 0, true);
    ctx->Gray16LutFor(*desc);
    if (out_stats != nullptr)
    {
        std::memset(out_stats, 0, sizeof(*out_stats));
        if (table != nullptr)
        {
            out_stats->applicable = 1;
            out_stats->valid = table->valid ? 1 : 0;
            out_stats->steps = table->stats.steps;
            out_stats->bands = table->stats.bands;
            out_stats->widest_band_ulps = table->stats.widestBand;
            out_stats->bucket_count = table->view.bucketCount;
            out_stats->swept_inputs = table->stats.sweptInputs;
            out_stats->in_band_inputs = table->stats.inBandInputs;
            out_stats->verify_mismatches = table->stats.verifyMismatches;
            out_stats->build_ms = table->stats.buildMilliseconds;
        }
    }
    if (table != nullptr && !table->valid)
    {
        ctx->lastError = "step table not used (generic exact kernel serves this configuration): " + table->error;
    }
    return AVIFGPU_OK;
}

AVIFGPU_EXPORT int avifgpu_set_table_autobuild(avifgpu_context* ctx, int64_t pixels)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    ctx->tableAutoBuildPixels = pixels;
    return AVIFGPU_OK;
}

// ---- primitives -------------------------------------------------------------------------------------------------

AVIFGPU_EXPORT int avifgpu_transfer_f32(avifgpu_context* ctx, int32_t function, float param, const float* in, float* out, size_t n)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    if (function < AVIFGPU_FN_LINEAR_TO_PQ || function > AVIFGPU_FN_LOGF)
    {
        return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "unknown function");
    }
    if (n == 0)
    {
        return AVIFGPU_OK;
    }
    if (in == nullptr || out == nullptr)
    {
        return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "NULL buffer");
    }
    DeviceGuard guard(ctx->device);
    int status;
    const size_t bytes = n * sizeof(float);
    if ((status = ctx->EnsureDevice(ctx->transferScratch[0], bytes)) != AVIFGPU_OK) return status;
    if ((status = ctx->EnsureDevice(ctx->transferScratch[1], bytes)) != AVIFGPU_OK) return status;
    cudaStream_t stream = ctx->streams[0];
    if ((status = ctx->Cuda(cudaMemcpyAsync(ctx->transferScratch[0].ptr, in, bytes, cudaMemcpyHostToDevice, stream), "H2D")) != AVIFGPU_OK) return status;
    const int launched = LaunchTransfer(function, param, static_cast<const float*>(ctx->transferScratch[0].ptr),
                                        static_cast<float*>(ctx->transferScratch[1].ptr), n, stream);
    if (launched < 0)
    {
        return ctx->LaunchFailed(launched, "transfer kernel launch");
    }
    ctx->launches += launched;
    if ((status = ctx->Cuda(cudaMemcpyAsync(out, ctx->transferScratch[1].ptr, bytes, cudaMemcpyDeviceToHost, stream), "D2H")) != AVIFGPU_OK) return status;
    return ctx->Cuda(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
}

AVIFGPU_EXPORT int avifgpu_hlg_ootf_f32(avifgpu_context* ctx, int32_t inverse, int32_t color_primaries, float display_gamma,
                                        float nominal_peak_nits, const float* rgb_in, float* rgb_out, size_t pixels)
{
    if (ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    float luma[3];
    if (!GetHlgLumaCoefficients(color_primaries, luma))
    {
        return ctx->Fail(AVIFGPU_ERR_UNSUPPORTED, "no HLG luma coefficients for these colour primaries");
    }
    if (pixels == 0)
    {
        return AVIFGPU_OK;
    }
    if (rgb_in == nullptr || rgb_out == nullptr)
    {
        return ctx->Fail(AVIFGPU_ERR_BAD_PARAM, "NULL buffer");
    }
    DeviceGuard guard(ctx->device);
    int status;
    const size_t bytes = pixels * 3 * sizeof(float);
    if ((status = ctx->EnsureDevice(ctx->transferScratch[0], bytes)) != AVIFGPU_OK) return status;
    if ((status = ctx->EnsureDevice(ctx->transferScratch[1], bytes)) != AVIFGPU_OK) return status;
    cudaStream_t stream = ctx->streams[0];
    if ((status = ctx->Cuda(cudaMemcpyAsync(ctx->transferScratch[0].ptr, rgb_in, bytes, cudaMemcpyHostToDevice, stream), "H2D")) != AVIFGPU_OK) return status;
    const int launched = LaunchHlgOotf(inverse != 0, luma, display_gamma, nominal_peak_nits, static_cast<const float*>(ctx->transferScratch[0].ptr),
                                       static_cast<float*>(ctx->transferScratch[1].ptr), pixels, stream);
    if (launched < 0)
    {
        return ctx->LaunchFailed(launched, "OOTF kernel launch");
    }
    ctx->launches += launched;
    if ((status = ctx->Cuda(cudaMemcpyAsync(rgb_out, ctx->transferScratch[1].ptr, bytes, cudaMemcpyDeviceToHost, stream), "D2H")) != AVIFGPU_OK) return status;
    return ctx->Cuda(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
}

} // extern "C"
