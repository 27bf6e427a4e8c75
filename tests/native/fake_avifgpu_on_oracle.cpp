// tests/native/fake_avifgpu_on_oracle.cpp -- TEST INFRASTRUCTURE: a stand-in for the GPU-touching entry points of
// include/avifgpu.h that converts with the CPU oracle, so that the C++ row shuttle (avif-format_b200/host/GpuRowShuttle.cpp)
// can be exercised on a machine without a GPU: row protocol, block partition, staging budget, strides, double buffering,
// error mapping, the colour-profile guard and seam.  Linked into tests/native/host_shuttle_test.cpp by
// tests/test_host_shuttle_cpu.py; the executable's own definitions take precedence over libavifgpu.so's, whose pure host
// helpers (avifgpu_encode_host_col_bytes, avifgpu_icc_to_rec2020_linear_matrix, avifgpu_status_string, ...) are still the
// real ones.  Never part of the product.
//
// A row block [y0, y0 + nrows) goes to the oracle as an image of nrows rows whose planes start at the block's first row
// (avif_oracle.h: "a row block presented as an image"), exactly the contract of avifgpu_encode_rows / avifgpu_decode_rows.
#include "../../include/avifgpu.h"
#include "../../oracle/avif_oracle.h"

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct avifgpu_context
{
    std::string lastError;
    int64_t lastTicket = 0;
    int64_t calls = 0;
};

namespace
{

int ChromaShiftY(int32_t chroma) { return chroma == AVIFGPU_CHROMA_420 ? 1 : 0; }

int Fail(avifgpu_context* ctx, int status, const char* message)
{
    if (ctx != nullptr)
    {
        ctx->lastError = message;
    }
    return status;
}

// plane k of the block that starts at image row y0 (chroma planes of 4:2:0 images have half the rows)
avifgpu_planes BlockPlanes(const avifgpu_planes& whole, int32_t y0, bool yCbCr, int32_t chroma)
{
    avifgpu_planes block = whole;
    for (int k = 0; k < AVIFGPU_MAX_PLANES; ++k)
    {
        if (whole.data[k] == nullptr)
        {
            continue;
        }
        const int shift = (yCbCr && (k == 1 || k == 2)) ? ChromaShiftY(chroma) : 0;
        block.data[k] = static_cast<uint8_t*>(whole.data[k]) + static_cast<int64_t>(y0 >> shift) * whole.stride[k];
    }
    return block;
}

} // namespace

extern "C"
{

int avifgpu_create(int, avifgpu_context** out_ctx)
{
    if (out_ctx == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    *out_ctx = new avifgpu_context();
    return AVIFGPU_OK;
}

void avifgpu_destroy(avifgpu_context* ctx) { delete ctx; }

const char* avifgpu_last_error(const avifgpu_context* ctx) { return ctx != nullptr ? ctx->lastError.c_str() : "no context"; }

int64_t avifgpu_launch_count(const avifgpu_context* ctx) { return ctx != nullptr ? ctx->calls : 0; }

int avifgpu_synchronize(avifgpu_context*) { return AVIFGPU_OK; }

int avifgpu_host_alloc(avifgpu_context*, size_t bytes, void** out_ptr)
{
    if (out_ptr == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    void* p = nullptr;
    if (posix_memalign(&p, 4096, bytes > 0 ? bytes : 1) != 0)
    {
        return AVIFGPU_ERR_OOM;
    }
    std::memset(p, 0xA5, bytes); // a shuttle that reads staging it has not filled shows up in the comparison
    *out_ptr = p;
    return AVIFGPU_OK;
}

int avifgpu_host_free(avifgpu_context*, void* ptr)
{
    std::free(ptr);
    return AVIFGPU_OK;
}

int avifgpu_prepare_encode(avifgpu_context*, const avifgpu_encode_desc*, avifgpu_curve_stats* out_stats)
{
    if (out_stats != nullptr)
    {
        std::memset(out_stats, 0, sizeof(*out_stats));
    }
    return AVIFGPU_OK;
}

int avifgpu_set_table_autobuild(avifgpu_context*, int64_t) { return AVIFGPU_OK; }

int avifgpu_encode_rows(avifgpu_context* ctx, const avifgpu_encode_desc* desc, const void* host_rows, int64_t row_stride_bytes, int32_t y0,
                        int32_t nrows, const avifgpu_planes* dst)
{
    if (ctx == nullptr || desc == nullptr || dst == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    ++ctx->lastTicket;
    ++ctx->calls;
    if (y0 < 0 || nrows < 0 || y0 + nrows > desc->height)
    {
        return Fail(ctx, AVIFGPU_ERR_BAD_PARAM, "row block outside the image");
    }
    const bool planar = desc->layout == AVIFGPU_LAYOUT_PLANAR_YCBCR;
    const int ys = planar ? ChromaShiftY(desc->chroma) : 0;
    if (ys && ((y0 & 1) || ((nrows & 1) && y0 + nrows != desc->height)))
    {
        return Fail(ctx, AVIFGPU_ERR_BAD_PARAM, "4:2:0 row blocks start on even rows and hold an even number of rows");
    }
    if (nrows == 0 || desc->width == 0)
    {
        return AVIFGPU_OK;
    }
    avifgpu_encode_desc block = *desc;
    block.height = nrows;
    const avifgpu_planes planes = BlockPlanes(*dst, y0, planar, desc->chroma);
    const int status = avif_oracle_encode_image(&block, host_rows, row_stride_bytes, &planes);
    if (status != 0)
    {
        return Fail(ctx, status, avif_oracle_last_error());
    }
    return AVIFGPU_OK;
}

int avifgpu_decode_rows(avifgpu_context* ctx, const avifgpu_decode_desc* desc, const avifgpu_planes* src, int32_t y0, int32_t nrows,
                        void* host_rows, int64_t row_stride_bytes)
{
    if (ctx == nullptr || desc == nullptr || src == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    ++ctx->lastTicket;
    ++ctx->calls;
    if (y0 < 0 || nrows < 0 || y0 + nrows > desc->height)
    {
        return Fail(ctx, AVIFGPU_ERR_BAD_PARAM, "row block outside the image");
    }
    if (nrows == 0 || desc->width == 0)
    {
        return AVIFGPU_OK;
    }
    const bool yCbCr = desc->colorspace == AVIFGPU_COLORSPACE_YCBCR;
    const int ys = yCbCr ? ChromaShiftY(desc->chroma) : 0;
    const int lead = y0 & ys; // a block may start on an odd row of a 4:2:0 image: decode from the row above it and drop that row
    avifgpu_decode_desc block = *desc;
    block.height = nrows + lead;
    const avifgpu_planes planes = BlockPlanes(*src, y0 - lead, yCbCr, desc->chroma);
    int status;
    if (lead == 0)
    {
        status = avif_oracle_decode_image(&block, &planes, host_rows, row_stride_bytes);
    }
    else
    {
        std::vector<uint8_t> scratch(static_cast<size_t>(row_stride_bytes) * block.height);
        status = avif_oracle_decode_image(&block, &planes, scratch.data(), row_stride_bytes);
        if (status == 0)
        {
            const int64_t payload = static_cast<int64_t>(desc->width) * avifgpu_decode_host_col_bytes(desc);
            for (int r = 0; r < nrows; ++r)
            {
                std::memcpy(static_cast<uint8_t*>(host_rows) + r * row_stride_bytes, scratch.data() + (r + lead) * row_stride_bytes, payload);
            }
        }
    }
    if (status != 0)
    {
        return Fail(ctx, status, avif_oracle_last_error());
    }
    return AVIFGPU_OK;
}

int avifgpu_encode_rows_async(avifgpu_context* ctx, const avifgpu_encode_desc* desc, const void* host_rows, int64_t row_stride_bytes, int32_t y0,
                              int32_t nrows, const avifgpu_planes* dst, int64_t* out_ticket)
{
    const int status = avifgpu_encode_rows(ctx, desc, host_rows, row_stride_bytes, y0, nrows, dst);
    if (out_ticket != nullptr && ctx != nullptr)
    {
        *out_ticket = ctx->lastTicket;
    }
    return status;
}

int avifgpu_decode_rows_async(avifgpu_context* ctx, const avifgpu_decode_desc* desc, const avifgpu_planes* src, int32_t y0, int32_t nrows,
                              void* host_rows, int64_t row_stride_bytes, int64_t* out_ticket)
{
    const int status = avifgpu_decode_rows(ctx, desc, src, y0, nrows, host_rows, row_stride_bytes);
    if (out_ticket != nullptr && ctx != nullptr)
    {
        *out_ticket = ctx->lastTicket;
    }
    return status;
}

int avifgpu_wait(avifgpu_context* ctx, int64_t) { return ctx != nullptr ? AVIFGPU_OK : AVIFGPU_ERR_BAD_PARAM; }

} // extern "C"
