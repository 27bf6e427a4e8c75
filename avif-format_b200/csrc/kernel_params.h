// kernel_params.h -- plain parameter blocks handed to the CUDA kernels (device pointers, block-relative).
//
// "Block" = the row block [y0, y0 + rows) one launch converts.  All pointers already point at the first row of
// the block in their buffer (for sub-sampled chroma planes: at chroma row y0 >> ys), so kernels index rows from
// zero and never see y0 -- except `yPhase`, the parity of y0 for 4:2:0 decode, where an odd first row shares its
// chroma row with the row above it (ReadHeifImage.cpp:359 uvJ = y >> yChromaShift).
#ifndef AVIF_KERNEL_PARAMS_H
#define AVIF_KERNEL_PARAMS_H

#include <stdint.h>

#include <atomic>

#include "curve_tables.h"
#include "pixel_math.cuh"

namespace avifgpu
{

#if defined(__CUDACC__)
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting; `configuredDevices` (one static per kernel
// instantiation) remembers the devices it has been made on, so a process that drives several GPUs configures each.
template <typename Kernel>
inline cudaError_t AllowDynamicShared(Kernel kernel, int bytes, std::atomic<uint64_t>& configuredDevices)
{
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess)
    {
        return e;
    }
    const uint64_t bit = 1ull << (device & 63);
    if (configuredDevices.load(std::memory_order_acquire) & bit)
    {
        return cudaSuccess;
    }
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess)
    {
        configuredDevices.fetch_or(bit, std::memory_order_release);
    }
    return e;
}
#endif

struct EncodeParams
{
    const void* rows;      // interleaved host pixels (formatRecord->data layout), device memory
    int64_t rowStride;     // bytes
    void* plane[4];        // REFERENCE layout: [0] interleaved or Y, [3] alpha (gray); PLANAR: Y, Cb, Cr, A
    int64_t planeStride[4];
    int32_t width;
    int32_t rowCount;      // rows in this block
    int32_t channels;      // 1..4
    int32_t hasAlpha;
    int32_t premultiply;
    int32_t imageDepth;    // 8, 10, 12
    uint32_t maxCode;
    float maxCodeFloat;
    int32_t transfer;      // avifgpu_transfer (float hosts)
    float pqMultiplier;    // peak / 10000.0f
    int32_t gray16Smpte428;
    int32_t planar;        // AVIFGPU_LAYOUT_PLANAR_YCBCR
    int32_t xs, ys;        // chroma shifts
    int32_t topLeft;       // AVIFGPU_DOWN_FILTER_TOP_LEFT
    avifpix::ForwardMatrix matrix;
    float chromaOffset;
    int32_t hlgInverseOotf; // AVIFGPU_HLG_INVERSE_OOTF_THEN_OETF: ApplyInverseHLGOOTF on the pixel before LinearToHLG
    float hlgLuma[3];
    float hlgDisplayGamma;
    float hlgPeak;
    int32_t rowMatrixEnabled; // avifgpu_encode_desc.row_matrix: the colour-profile 3x3 ahead of everything else (float colour hosts)
    float rowMatrix[9];
    // Float hosts with a transfer curve: the flat step table + band bitmap in global memory (a by-value copy of
    // *curveTable made by the generic launcher), or useCurveView = 0 -> every sample takes the exact powf.
    CurveTableView curveView;
    int32_t useCurveView;
    // Host-side extras for the launcher (ignored by the kernels):
    const CurveTableView* curveTable; // verified exact step table for `transfer`, or nullptr
    const uint16_t* gray16Lut;        // 65536-entry code table for Gray16 hosts (device memory), or nullptr
    int32_t smCount;
    int32_t verifiedPremultiply;      // 1 once the context has verified FastPremultiplyBiased for this image depth on this device
};

struct DecodeParams
{
    const void* plane[4];  // YCbCr: Y, Cb, Cr, A; mono: Y, -, -, A; planar RGB: R, G, B, A
    int64_t planeStride[4];
    void* rows;            // interleaved host pixels, device memory
    int64_t rowStride;
    int32_t width;
    int32_t rowCount;
    int32_t yPhase;        // y0 & ys
    int32_t colorspace;    // avifgpu_colorspace
    int32_t xs, ys;
    int32_t hasAlpha;
    int32_t premultiplied;
    int32_t bitDepth;
    uint32_t maxCode;
    avifpix::RangeParams range;
    avifpix::InverseMatrix matrix;
    int32_t hostDepth;     // 8, 16, 32
    int32_t transfer;      // avifgpu_transfer (host depth 32)
    float pqMultiplier;    // 10000.0f / peak
    int32_t applyOotf;
    float lumaR, lumaG, lumaB;
    float gammaMinusOne;
    float hlgPeak;
    int32_t smCount;              // host-side extra for the launcher
    int32_t verifiedHlgDivisions; // 1 once the context has verified HLGToLinearUnit's fast divisions on this device
    int32_t verifiedGreenDivision; // 1 once the context has verified the fast `/ kg` of YuvDecode.cpp:308 for this matrix, depth, range
    int32_t verifiedPqRatio;       // 1 once the context has verified the branch-free division inside PQToLinear on this device
};

// A launcher that sees a CUDA error has already consumed it (cudaGetLastError clears the slot), so it leaves the code
// here -- a thread-local slot in avifgpu_api.cu -- and returns AVIFGPU_ERR_CUDA; the API reports it from there instead
// of asking CUDA a second time (which would answer cudaSuccess and turn a failed launch into AVIFGPU_OK).
int ReportLaunchFailure(int cudaErrorCode);

// Launchers implemented in kernels_*.cu.  They only enqueue work on `stream` and return the number of kernels
// launched (>= 1) or a negative avifgpu_status.
int LaunchEncode(const EncodeParams& params, int hostDepth, void* stream);
int LaunchDecode(const DecodeParams& params, void* stream);
int LaunchTransfer(int function, float param, const float* in, float* out, size_t count, void* stream);

} // namespace avifgpu

#endif
