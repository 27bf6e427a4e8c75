/*
 * avifgpu.h -- C ABI of the B200-native colour-conversion hot path of the avif-format plug-in.
 *
 * This header is the drop-in boundary: plain pointers and sizes, no C++ / torch types.  Every entry point
 * replaces one seam of the reference (citations are relative to the reference tree, src/common/):
 *
 *   avifgpu_encode_rows*          the per-row loops of CreateHeifImage{Gray,RGB}{Eight,Sixteen,ThirtyTwo}Bit
 *                                 (WriteHeifImage.cpp:169-1139; called from Write.cpp:303-336) and, for
 *                                 AVIFGPU_LAYOUT_PLANAR_YCBCR, additionally the RGB->YCbCr matrix + chroma
 *                                 down-sampling that the reference delegates to libheif inside
 *                                 heif_context_encode_image (Write.cpp:44; matrix chosen at
 *                                 WriteMetadata.cpp:107-149).
 *   avifgpu_decode_rows*          the per-row loops of ReadHeifImage{Gray,RGB}{Eight,Sixteen,ThirtyTwo}Bit
 *                                 (ReadHeifImage.cpp:83-1178; called from Read.cpp:587-630), i.e. the twelve
 *                                 Decode*Row* functions of YUVDecode.h:29-147 plus the planar-RGB branches.
 *   avifgpu_get_yuv_coefficients  GetYUVCoefficiants (YUVCoefficiants.cpp:154-188).
 *   avifgpu_build_yuv_tables      YUVLookupTables::YUVLookupTables (YuvLookupTables.cpp:115-192).
 *   avifgpu_transfer_f32          LinearToPQ / PQToLinear / LinearToSMPTE428 / SMPTE428ToLinear /
 *                                 HLGToLinear / LinearToHLG (ColorTransfer.cpp:69-190), evaluated on the GPU
 *                                 with the device libm of csrc/device_math.cuh (used by the primitive-level
 *                                 parity gates).
 *
 * The "_device" variants take device pointers and a cudaStream_t (as void*) and never synchronise: they are
 * what a caller that keeps frames resident in HBM uses, and what bench.py times for `value`.  The plain
 * variants take HOST pointers, move the row block over PCIe through pinned staging owned by the context and
 * return when the destination host memory is valid: they are what the plug-in's FormatRecord row shuttle
 * binds, and what bench.py times for `e2e`.
 *
 * There is no CPU fallback.  If no CUDA device is usable avifgpu_create fails with AVIFGPU_ERR_NO_DEVICE.
 *
 * Error convention (reference: exceptions mapped to OSErr at Write.cpp:345-364 / Read.cpp:659-678): every
 * function returns 0 on success or a negative avifgpu_status; nothing throws across this boundary.  The C++
 * host mirror (avif-format_b200/host) turns the codes back into OSErrException / std::runtime_error.
 */
#ifndef AVIFGPU_H
#define AVIFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32) && defined(AVIFGPU_BUILDING_LIBRARY)
#define AVIFGPU_EXPORT __declspec(dllexport)
#elif defined(_WIN32)
#define AVIFGPU_EXPORT __declspec(dllimport) /* the plug-in (an MSVC project) consuming avifgpu.dll */
#else
#define AVIFGPU_EXPORT __attribute__((visibility("default")))
#endif

#define AVIFGPU_API_VERSION 4

typedef enum avifgpu_status
{
    AVIFGPU_OK = 0,
    AVIFGPU_ERR_BAD_PARAM = -1,    /* -> formatBadParameters                                   */
    AVIFGPU_ERR_UNSUPPORTED = -2,  /* -> std::runtime_error("Unsupported ...") in the reference */
    AVIFGPU_ERR_NO_DEVICE = -3,    /* -> errPlugInHostInsufficient                              */
    AVIFGPU_ERR_CUDA = -4,         /* -> std::runtime_error                                     */
    AVIFGPU_ERR_OOM = -5,          /* -> memFullErr / std::bad_alloc                            */
    AVIFGPU_ERR_CANCELED = -6      /* -> userCanceledErr (abortProc polled between row blocks)  */
} avifgpu_status;

/* AlphaState.h:24-29 (same order). */
typedef enum avifgpu_alpha_state
{
    AVIFGPU_ALPHA_NONE = 0,
    AVIFGPU_ALPHA_STRAIGHT = 1,
    AVIFGPU_ALPHA_PREMULTIPLIED = 2
} avifgpu_alpha_state;

/* ColorTransfer.h:28-34 (same order). */
typedef enum avifgpu_transfer
{
    AVIFGPU_TRANSFER_PQ = 0,
    AVIFGPU_TRANSFER_HLG = 1,
    AVIFGPU_TRANSFER_SMPTE428 = 2,
    AVIFGPU_TRANSFER_CLIP = 3
} avifgpu_transfer;

/* Numeric values = libheif's heif_chroma for the planar kinds. */
typedef enum avifgpu_chroma
{
    AVIFGPU_CHROMA_MONOCHROME = 0,
    AVIFGPU_CHROMA_420 = 1,
    AVIFGPU_CHROMA_422 = 2,
    AVIFGPU_CHROMA_444 = 3
} avifgpu_chroma;

/* Numeric values = libheif's heif_colorspace. */
typedef enum avifgpu_colorspace
{
    AVIFGPU_COLORSPACE_YCBCR = 0,
    AVIFGPU_COLORSPACE_RGB = 1,
    AVIFGPU_COLORSPACE_MONOCHROME = 2
} avifgpu_colorspace;

typedef enum avifgpu_layout
{
    /* What the reference itself produces: heif_channel_interleaved RGB(A) for colour hosts
     * (WriteHeifImage.cpp:63-85, 639-646) and heif_channel_Y (+ heif_channel_Alpha) planes for gray hosts
     * (WriteHeifImage.cpp:177-194).  plane[0] = interleaved / Y, plane[3] = Alpha (gray only). */
    AVIFGPU_LAYOUT_REFERENCE = 0,
    /* Fused: additionally applies the forward matrix + chroma down-filter and writes heif_channel_Y / Cb /
     * Cr (/ Alpha) planes at image_bit_depth.  plane[0..3] = Y, Cb, Cr, Alpha.  Colour hosts only. */
    AVIFGPU_LAYOUT_PLANAR_YCBCR = 1
} avifgpu_layout;

typedef enum avifgpu_down_filter
{
    AVIFGPU_DOWN_FILTER_BOX = 0,     /* mean of the 2x1 / 2x2 float chroma samples, then quantise */
    AVIFGPU_DOWN_FILTER_TOP_LEFT = 1 /* co-sited top-left sample (libheif 1.14-like)              */
} avifgpu_down_filter;

/* Extra transfer option outside the reference's ColorTransferFunction enum: BASELINE.json config 5
 * (Gray16 -> 12-bit monochrome SMPTE 428-1), a composition this project defines. */
typedef enum avifgpu_gray16_curve
{
    AVIFGPU_GRAY16_LUT = 0,      /* reference behaviour: BuildSixteenBitToHeifImageLookup, WriteHeifImage.cpp:140-166 */
    AVIFGPU_GRAY16_SMPTE428 = 1  /* code = (u16)clamp(LinearToSMPTE428(v/32768f)*max, 0, max)                         */
} avifgpu_gray16_curve;

/* Mirrors heif_color_profile_nclx as far as the path reads it (YuvLookupTables.cpp:143-144,
 * YUVCoefficiants.cpp:110-152, ColorTransfer.cpp:31-67).  present == 0 means "nclx == nullptr". */
typedef struct avifgpu_nclx
{
    int32_t present;
    int32_t color_primaries;          /* H.273 code point */
    int32_t transfer_characteristics; /* H.273 code point */
    int32_t matrix_coefficients;      /* H.273 code point */
    int32_t full_range_flag;
} avifgpu_nclx;

#define AVIFGPU_MAX_PLANES 4

/* A set of image planes.  stride in BYTES (libheif pads rows: heif_image_get_plane's out_stride).
 * Samples deeper than 8 bit are native-endian uint16 with the value in the low bits. */
typedef struct avifgpu_planes
{
    void* data[AVIFGPU_MAX_PLANES];
    int64_t stride[AVIFGPU_MAX_PLANES];
} avifgpu_planes;

/* Parameter block of the encode direction = FormatRecord fields + SaveUIOptions fields the loops read
 * (AvifFormat.h:87-101, Write.cpp:229-258 fix-ups are the CALLER's job and are mirrored in host/). */
typedef enum avifgpu_hlg_extension
{
    AVIFGPU_HLG_REJECT = 0,                /* what the reference does */
    AVIFGPU_HLG_OETF = 1,                  /* scene-referred input: code = quantise(LinearToHLG(c)) */
    AVIFGPU_HLG_INVERSE_OOTF_THEN_OETF = 2 /* display-referred input: ApplyInverseHLGOOTF(rgb) first; needs nclx primaries */
} avifgpu_hlg_extension;

typedef struct avifgpu_encode_desc
{
    uint32_t struct_size;    /* sizeof(avifgpu_encode_desc) */
    int32_t width;           /* imageSize.h */
    int32_t height;          /* imageSize.v */
    int32_t host_depth;      /* formatRecord->depth: 8, 16 (0..32768) or 32 (float) */
    int32_t host_channels;   /* formatRecord->planes: 1 Gray, 2 Gray+A, 3 RGB, 4 RGB+A */
    int32_t alpha_state;     /* avifgpu_alpha_state */
    int32_t image_bit_depth; /* SaveUIOptions.imageBitDepth: 8, 10 or 12 */
    int32_t transfer;        /* SaveUIOptions.hdrTransferFunction (32-bit hosts only) */
    int32_t pq_peak_nits;    /* SaveUIOptions.pq.nominalPeakBrightness */
    int32_t layout;          /* avifgpu_layout */
    int32_t chroma;          /* avifgpu_chroma, for AVIFGPU_LAYOUT_PLANAR_YCBCR */
    int32_t down_filter;     /* avifgpu_down_filter */
    int32_t gray16_curve;    /* avifgpu_gray16_curve, Gray16 hosts only */
    avifgpu_nclx nclx;       /* matrix for PLANAR_YCBCR (WriteMetadata.cpp:113-146); full range only */
    /* HLG save path.  The reference ships LinearToHLG and ApplyInverseHLGOOTF (ColorTransfer.cpp:141-164,
     * 207-220) but no caller: with hlg_extension = 0 transfer = AVIFGPU_TRANSFER_HLG is rejected with the reference's own
     * "Unsupported color transfer function." (WriteHeifImage.cpp:1085-1087).  Colour float hosts only. */
    int32_t hlg_extension;       /* avifgpu_hlg_extension */
    float hlg_display_gamma;     /* for AVIFGPU_HLG_INVERSE_OOTF_THEN_OETF (LoadUIOptions.hlg.displayGamma's counterpart) */
    int32_t hlg_peak_nits;       /* nominal peak brightness of the display-referred input */
    /* Colour-profile step on the GPU, colour float hosts only.  The reference converts every host row
     * with lcms2 before anything else (ColorProfileConversion::ConvertRow at WriteHeifImage.cpp:1028-1031); for a
     * linear-light document in a matrix profile that conversion is one 3x3 matrix, applied here per pixel in binary32,
     * alpha untouched (cmsFLAGS_COPY_ALPHA), before the clamp / premultiply / transfer curve:
     *     r' = (m[0] r + m[1] g) + m[2] b,  g' = (m[3] r + m[4] g) + m[5] b,  b' = (m[6] r + m[7] g) + m[8] b
     * avifgpu_icc_to_rec2020_linear_matrix() derives it from an ICC profile.  Parity unpinned (lcms2 is not in the tree). */
    int32_t row_matrix_enabled;
    float row_matrix[9];
} avifgpu_encode_desc;

/* Parameter block of the decode direction = heif_image properties + nclx + LoadUIOptions
 * (AvifFormat.h:61-85). */
typedef struct avifgpu_decode_desc
{
    uint32_t struct_size;   /* sizeof(avifgpu_decode_desc) */
    int32_t width;
    int32_t height;
    int32_t colorspace;     /* avifgpu_colorspace */
    int32_t chroma;         /* avifgpu_chroma (YCbCr / monochrome); ignored for planar RGB */
    int32_t bit_depth;      /* heif_image_get_bits_per_pixel_range: 8, 10, 12 or 16 */
    int32_t alpha_state;    /* avifgpu_alpha_state */
    int32_t host_depth;     /* 8, 16 or 32: which ReadHeifImage*Bit variant */
    avifgpu_nclx nclx;
    int32_t hlg_apply_ootf;      /* LoadUIOptions.hlg.applyOOTF */
    float hlg_display_gamma;     /* LoadUIOptions.hlg.displayGamma */
    int32_t hlg_peak_nits;       /* LoadUIOptions.hlg.nominalPeakBrightness */
    int32_t pq_peak_nits;        /* LoadUIOptions.pq.nominalPeakBrightness */
} avifgpu_decode_desc;

typedef struct avifgpu_context avifgpu_context;

/* ---- context --------------------------------------------------------------------------------------- */

AVIFGPU_EXPORT int avifgpu_api_version(void);

/* Binds to CUDA device `device_ordinal` (must be compute capability 10.x).  No CPU fallback. */
AVIFGPU_EXPORT int avifgpu_create(int device_ordinal, avifgpu_context** out_ctx);
AVIFGPU_EXPORT void avifgpu_destroy(avifgpu_context* ctx);

/* Human-readable text for the last failure on this context (never NULL). ctx may be NULL for creation errors. */
AVIFGPU_EXPORT const char* avifgpu_last_error(const avifgpu_context* ctx);
AVIFGPU_EXPORT const char* avifgpu_status_string(int status);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
AVIFGPU_EXPORT int64_t avifgpu_launch_count(const avifgpu_context* ctx);

/* Blocks until all work issued through this context has finished. */
AVIFGPU_EXPORT int avifgpu_synchronize(avifgpu_context* ctx);

/* Pinned host memory for row buffers (the plug-in's replacement for its one-row ScopedBufferSuiteBuffer,
 * Write.cpp:297-299 / ReadHeifImage.cpp:113-115). */
AVIFGPU_EXPORT int avifgpu_host_alloc(avifgpu_context* ctx, size_t bytes, void** out_ptr);
AVIFGPU_EXPORT int avifgpu_host_free(avifgpu_context* ctx, void* ptr);

/* ---- geometry helpers (pure host arithmetic, usable without a device) ------------------------------- */

/* Bytes per host pixel (formatRecord->colBytes) for an encode / decode description. */
AVIFGPU_EXPORT int avifgpu_encode_host_col_bytes(const avifgpu_encode_desc* desc);
AVIFGPU_EXPORT int avifgpu_decode_host_col_bytes(const avifgpu_decode_desc* desc);

/* Width / height / bytes-per-sample of plane `index` of the encode destination or decode source.
 * Returns 0 and zeroes the outputs for planes that do not exist in that configuration. */
AVIFGPU_EXPORT int avifgpu_encode_plane_geometry(const avifgpu_encode_desc* desc, int index,
                                                 int32_t* out_width_samples, int32_t* out_height,
                                                 int32_t* out_bytes_per_sample);
AVIFGPU_EXPORT int avifgpu_decode_plane_geometry(const avifgpu_decode_desc* desc, int index,
                                                 int32_t* out_width_samples, int32_t* out_height,
                                                 int32_t* out_bytes_per_sample);

/* ---- parameter derivation (host arithmetic identical to the reference's) ---------------------------- */

/* GetYUVCoefficiants, YUVCoefficiants.cpp:154-188: out_kr_kg_kb[3]. */
AVIFGPU_EXPORT int avifgpu_get_yuv_coefficients(const avifgpu_nclx* nclx, float* out_kr_kg_kb);

/* GetHLGLumaCoefficients, ColorTransfer.cpp:31-45: out_rgb[3]; AVIFGPU_ERR_UNSUPPORTED for other primaries. */
AVIFGPU_EXPORT int avifgpu_get_hlg_luma_coefficients(int32_t color_primaries, float* out_rgb);

/* YUVLookupTables ctor, YuvLookupTables.cpp:115-192.  Each non-NULL table receives 1 << bit_depth floats. */
AVIFGPU_EXPORT int avifgpu_build_yuv_tables(const avifgpu_nclx* nclx, int32_t bit_depth, int32_t monochrome,
                                            float* out_table_y, float* out_table_uv, float* out_table_alpha);

/* ICC profile -> the matrix of avifgpu_encode_desc.row_matrix for an HDR save (document RGB, linear light -> linear
 * Rec.2020, what ColorProfileConversion::InitializeForRec2020Conversion sets up, ColorProfileConversion.cpp:240-266).
 * AVIFGPU_OK: out_matrix9 filled; *out_is_rec2020 = 1 when the profile already is Rec.2020 (the reference then converts
 * nothing, ColorProfileConversion.cpp:128-131).  AVIFGPU_ERR_UNSUPPORTED: not a matrix / TRC RGB profile with identity
 * tone curves (LUT profiles, gamma-encoded profiles: those stay with the host's lcms2).  Pure host arithmetic. */
AVIFGPU_EXPORT int avifgpu_icc_to_rec2020_linear_matrix(const void* icc_profile, size_t size, float* out_matrix9, int32_t* out_is_rec2020);

/* ---- the hot path: host-pointer variants (PCIe inside) ---------------------------------------------- */

/*
 * Converts host rows [y0, y0 + nrows) and stores them into the destination planes.
 *   host_rows   interleaved host pixels of row y0 (formatRecord->data after advanceState for
 *               theRect = {top = y0, bottom = y0 + nrows}); row_stride_bytes = formatRecord->rowBytes.
 *   dst         HOST pointers to the origin (row 0) of each whole-image plane, e.g. heif_image_get_plane().
 * For 4:2:0 output y0 must be even and nrows even unless y0 + nrows == height.
 */
AVIFGPU_EXPORT int avifgpu_encode_rows(avifgpu_context* ctx, const avifgpu_encode_desc* desc,
                                       const void* host_rows, int64_t row_stride_bytes,
                                       int32_t y0, int32_t nrows, const avifgpu_planes* dst);

/*
 * Converts image rows [y0, y0 + nrows) of the source planes into interleaved host rows.
 *   src         HOST pointers to the origin of each whole-image plane (heif_image_get_plane_readonly()):
 *               YCbCr: Y, Cb, Cr, Alpha; monochrome: Y, -, -, Alpha; planar RGB: R, G, B, Alpha.
 *   host_rows   receives row y0 first; row_stride_bytes = formatRecord->rowBytes.
 */
AVIFGPU_EXPORT int avifgpu_decode_rows(avifgpu_context* ctx, const avifgpu_decode_desc* desc,
                                       const avifgpu_planes* src, int32_t y0, int32_t nrows,
                                       void* host_rows, int64_t row_stride_bytes);

/* ---- the hot path: asynchronous host-pointer variants ------------------------------------------------------ */

/*
 * The same conversions, enqueued: the call returns once the work is on the device's queues, so the caller's thread can
 * produce the next row block (the plug-in: advanceState, i.e. Photoshop filling the next rows, Write.cpp:297-336) while
 * the GPU converts and copies this one.  Single-threaded contract like the rest of a context: calls on one context come
 * from one thread at a time.
 *   encode: on return the rows of THIS call may still be read by the copy engine when host_rows is page-locked memory
 *           (avifgpu_host_alloc); they may be overwritten once the NEXT call on this context has returned (two row
 *           buffers make a double-buffered shuttle) or avifgpu_wait() has.  Pageable rows are consumed before return.
 *           The destination planes are complete after avifgpu_wait(ticket).
 *   decode: the source planes must stay valid and the host rows are complete after avifgpu_wait(ticket).
 * out_ticket (may be NULL) identifies the call; tickets increase by one per host-pointer call on a context.
 */
AVIFGPU_EXPORT int avifgpu_encode_rows_async(avifgpu_context* ctx, const avifgpu_encode_desc* desc,
                                             const void* host_rows, int64_t row_stride_bytes,
                                             int32_t y0, int32_t nrows, const avifgpu_planes* dst, int64_t* out_ticket);
AVIFGPU_EXPORT int avifgpu_decode_rows_async(avifgpu_context* ctx, const avifgpu_decode_desc* desc,
                                             const avifgpu_planes* src, int32_t y0, int32_t nrows,
                                             void* host_rows, int64_t row_stride_bytes, int64_t* out_ticket);
/* Blocks until every host-pointer call up to and including `ticket` has completed and its results are in the
 * caller's memory; ticket <= 0 waits for everything issued so far. */
AVIFGPU_EXPORT int avifgpu_wait(avifgpu_context* ctx, int64_t ticket);

/* ---- the hot path: device-pointer variants (no copies, no synchronisation) --------------------------- */

/* Same contracts, but host_rows / planes are DEVICE pointers valid on the context's device and the work is
 * enqueued on `cuda_stream` (a cudaStream_t; NULL = the legacy default stream). */
AVIFGPU_EXPORT int avifgpu_encode_rows_device(avifgpu_context* ctx, const avifgpu_encode_desc* desc,
                                              const void* device_rows, int64_t row_stride_bytes,
                                              int32_t y0, int32_t nrows, const avifgpu_planes* device_dst,
                                              void* cuda_stream);

AVIFGPU_EXPORT int avifgpu_decode_rows_device(avifgpu_context* ctx, const avifgpu_decode_desc* desc,
                                              const avifgpu_planes* device_src, int32_t y0, int32_t nrows,
                                              void* device_rows, int64_t row_stride_bytes,
                                              void* cuda_stream);

/* ---- the hot path across several GPUs of one box ------------------------------------------------ */

/*
 * An image shards by row block (4:2:0: even boundaries, no halo; tables are replicated), so one process can put every
 * GPU of the box behind the same seam: the reference's row loops (WriteHeifImage.cpp:808-988 for a 16-bit RGBA save,
 * ReadHeifImage.cpp:290-400 for a float load, ...) have no cross-row state.  A shard group owns one context per device
 * and splits every call into `size` contiguous row blocks with avifgpu_shard_row_blocks().
 *
 *   host-pointer calls   block r is staged over device r's OWN PCIe link, converted there and copied back into the
 *                        caller's planes / rows: the PCIe-bound end-to-end rate scales with the number of links.
 *   device-pointer call  block r's rows already sit in device r's HBM; every device's conversion kernel stores its
 *                        part of the planes straight into the OWNER device's memory over NVLink / NVSwitch (peer
 *                        mapping inside the process), so the planar image is assembled when the kernels end -- the
 *                        "gather of the final planar buffer" is fused into the conversion, no collective pass.
 */
typedef struct avifgpu_shard_group avifgpu_shard_group;

/* One context per entry of device_ordinals[0 .. count) (distinct CUDA devices, compute capability 10.x); enables peer
 * access between them where the hardware offers it. */
AVIFGPU_EXPORT int avifgpu_shard_group_create(const int32_t* device_ordinals, int32_t count, avifgpu_shard_group** out_group);
AVIFGPU_EXPORT void avifgpu_shard_group_destroy(avifgpu_shard_group* group);
AVIFGPU_EXPORT int32_t avifgpu_shard_group_size(const avifgpu_shard_group* group);
/* The context of member `index` (owned by the group), e.g. for avifgpu_host_alloc or avifgpu_launch_count. */
AVIFGPU_EXPORT avifgpu_context* avifgpu_shard_group_context(avifgpu_shard_group* group, int32_t index);
/* 1 when member `from` can address member `to`'s memory (needed by the device-pointer call), else 0. */
AVIFGPU_EXPORT int avifgpu_shard_group_peer_access(const avifgpu_shard_group* group, int32_t from, int32_t to);
AVIFGPU_EXPORT const char* avifgpu_shard_group_last_error(const avifgpu_shard_group* group);
/* avifgpu_prepare_encode on every member (concurrently). */
AVIFGPU_EXPORT int avifgpu_shard_group_prepare_encode(avifgpu_shard_group* group, const avifgpu_encode_desc* desc);
/* Blocks until every member's work has finished. */
AVIFGPU_EXPORT int avifgpu_shard_group_synchronize(avifgpu_shard_group* group);

/* Row blocks of rows [y0, y0 + nrows) for `parts` members: contiguous, in order, every inner boundary even (a 2x2
 * chroma site never straddles two blocks); trailing blocks may be empty.  Pure host arithmetic. */
AVIFGPU_EXPORT int avifgpu_shard_row_blocks(int32_t y0, int32_t nrows, int32_t parts, int32_t* out_y0, int32_t* out_nrows);

/* avifgpu_encode_rows / avifgpu_decode_rows with the row block split across the group (same arguments, same result,
 * bit for bit).  y0 even for 4:2:0. */
AVIFGPU_EXPORT int avifgpu_encode_rows_sharded(avifgpu_shard_group* group, const avifgpu_encode_desc* desc,
                                               const void* host_rows, int64_t row_stride_bytes,
                                               int32_t y0, int32_t nrows, const avifgpu_planes* dst);
AVIFGPU_EXPORT int avifgpu_decode_rows_sharded(avifgpu_shard_group* group, const avifgpu_decode_desc* desc,
                                               const avifgpu_planes* src, int32_t y0, int32_t nrows,
                                               void* host_rows, int64_t row_stride_bytes);

/*
 * Device-resident frame, rows distributed: device_rows[r] / row_stride_bytes[r] = the first row of member r's block
 * (avifgpu_shard_row_blocks(0, desc->height, size)) in member r's memory.  owner_planes = whole-image planes in member
 * `owner`'s memory; every member must have peer access to it.  Enqueues one conversion per member on that member's
 * own stream and returns; avifgpu_shard_group_synchronize() (or the next sharded call) orders after it.
 */
AVIFGPU_EXPORT int avifgpu_encode_rows_sharded_device(avifgpu_shard_group* group, const avifgpu_encode_desc* desc,
                                                      const void* const* device_rows, const int64_t* row_stride_bytes,
                                                      const avifgpu_planes* owner_planes, int32_t owner);

/* ---- preparation (optional) --------------------------------------------------------------------------- */

/* Statistics of the exact float->code step table behind a float-host encode configuration
 * (avif-format_b200/csrc/curve_tables.h).  The table is derived on the device from the exact curve by sweeping
 * every non-negative float, and verified against it the same way, the first time a configuration is used. */
typedef struct avifgpu_curve_stats
{
    int32_t applicable;          /* 1 if this description uses a step table (float host, PQ or SMPTE 428) */
    int32_t valid;               /* 1 if the table was built and verified; 0 -> the exact generic kernel is used */
    int32_t steps;               /* thresholds found */
    int32_t bands;               /* thresholds with a non-empty fuzzy (non-monotone) band */
    uint32_t widest_band_ulps;
    int32_t bucket_count;
    uint64_t swept_inputs;       /* floats evaluated by the sweep */
    uint64_t in_band_inputs;     /* floats the kernel hands to the exact path */
    uint64_t verify_mismatches;  /* must be 0 for valid == 1 */
    double build_ms;
} avifgpu_curve_stats;

/* Builds whatever device-side tables `desc` needs, now.  out_stats may be NULL.
 * The exact step tables of the float PQ / SMPTE 428 encode path cost about 40 ms to build and verify (once per
 * context and configuration) and make that path several times faster, which pays off after roughly two gigapixels:
 * frame pipelines call this up front; a caller that converts one image does not need to (see
 * avifgpu_set_table_autobuild). */
AVIFGPU_EXPORT int avifgpu_prepare_encode(avifgpu_context* ctx, const avifgpu_encode_desc* desc,
                                          avifgpu_curve_stats* out_stats);

/* Without avifgpu_prepare_encode() the encode calls convert with the exact kernel (glibc-identical powf per sample)
 * until `pixels` pixels of one configuration have gone through this context, and build the step tables then.
 * Default AVIFGPU_TABLE_AUTOBUILD_DEFAULT; 0 = build at first use; negative = never build automatically.
 * Results are bit-identical either way. */
#define AVIFGPU_TABLE_AUTOBUILD_DEFAULT (((int64_t)1) << 31)
AVIFGPU_EXPORT int avifgpu_set_table_autobuild(avifgpu_context* ctx, int64_t pixels);

/* ---- primitive-level entry points (parity gates G2/G5; not on the plug-in's call path) -------------- */

typedef enum avifgpu_function
{
    AVIFGPU_FN_LINEAR_TO_PQ = 0,      /* param = peak nits  ColorTransfer.cpp:69-92   */
    AVIFGPU_FN_PQ_TO_LINEAR = 1,      /* param = peak nits  ColorTransfer.cpp:94-117  */
    AVIFGPU_FN_LINEAR_TO_SMPTE428 = 2,/*                    ColorTransfer.cpp:119-127 */
    AVIFGPU_FN_SMPTE428_TO_LINEAR = 3,/*                    ColorTransfer.cpp:129-139 */
    AVIFGPU_FN_HLG_TO_LINEAR = 4,     /*                    ColorTransfer.cpp:166-190 */
    AVIFGPU_FN_LINEAR_TO_HLG = 5,     /*                    ColorTransfer.cpp:141-164 */
    AVIFGPU_FN_POWF = 6,              /* param = exponent   libm powf                 */
    AVIFGPU_FN_EXPF = 7,              /*                    libm expf                 */
    AVIFGPU_FN_LOGF = 8               /*                    libm logf                 */
} avifgpu_function;

/* out[i] = fn(in[i], param) for n floats; in/out are HOST pointers (copied through the device). */
AVIFGPU_EXPORT int avifgpu_transfer_f32(avifgpu_context* ctx, int32_t function, float param,
                                        const float* in, float* out, size_t n);

/* ApplyHLGOOTF (ColorTransfer.cpp:192-205; inverse = 0) or ApplyInverseHLGOOTF (ColorTransfer.cpp:207-220; inverse = 1)
 * over `pixels` interleaved RGB float triples, luma coefficients from GetHLGLumaCoefficients(color_primaries)
 * (ColorTransfer.cpp:31-45).  rgb_in / rgb_out are HOST pointers (may alias). */
AVIFGPU_EXPORT int avifgpu_hlg_ootf_f32(avifgpu_context* ctx, int32_t inverse, int32_t color_primaries, float display_gamma,
                                        float nominal_peak_nits, const float* rgb_in, float* rgb_out, size_t pixels);

#ifdef __cplusplus
}
#endif

#endif /* AVIFGPU_H */
