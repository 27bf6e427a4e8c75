// GpuRowShuttle.h -- the plug-in's row shuttle entry points, backed by the B200 library (include/avifgpu.h).
//
// Same names, arguments and error behaviour as the reference's
//     CreateHeifImage{Gray,RGB}{Eight,Sixteen,ThirtyTwo}Bit   (src/common/WriteHeifImage.h:29-63)
//     ReadHeifImage{Gray,RGB}{Eight,Sixteen,ThirtyTwo}Bit     (src/common/ReadHeifImage.h:27-63)
// so the call sites in Write.cpp:303-336 and Read.cpp:587-630 do not change.  Differences, all behind the seam:
//   * the host is asked for blocks of rows (theRect top..bottom spanning many rows -- the FormatRecord protocol
//     allows any rectangle) into a pinned staging buffer instead of one row at a time, abortProc is still polled
//     once per block, and each block is one avifgpu_encode_rows / avifgpu_decode_rows call;
//   * colour images are handed to libheif as planar YCbCr at the encoder's chroma format (matrix + down-filter
//     fused on the GPU) instead of interleaved RGB, so heif_context_encode_image's own colour conversion
//     degenerates to a no-op; gray images keep the reference's Y (+ Alpha) planes;
//   * the optional ICC row transform (ColorProfileConversion, lcms2) is not on the accelerated path: the shuttle
//     requires "no transform" (keepColorProfile or no document profile) and throws formatBadParameters otherwise.
// Errors surface exactly like the reference's: OSErrException (userCanceledErr, formatBadParameters, ...),
// std::runtime_error for unsupported configurations, std::bad_alloc for memory.
#ifndef AVIFGPU_HOST_GPU_ROW_SHUTTLE_H
#define AVIFGPU_HOST_GPU_ROW_SHUTTLE_H

#include "compat/PluginTypes.h"

struct avifgpu_context;

namespace avifgpu_host
{

// Process-wide context for the plug-in (Photoshop calls PluginMain on one thread).  Throws
// OSErrException(errPlugInHostInsufficient) when there is no usable B200: there is no CPU fallback.
avifgpu_context* SharedContext();
void ReleaseSharedContext();

// Rows requested from / delivered to the host per advanceState call (rounded down to an even count).
void SetRowsPerBlock(int32 rows);

} // namespace avifgpu_host

ScopedHeifImage CreateHeifImageGrayEightBit(FormatRecordPtr formatRecord, AlphaState alphaState, const VPoint& imageSize, const SaveUIOptions& saveOptions);
ScopedHeifImage CreateHeifImageGraySixteenBit(FormatRecordPtr formatRecord, AlphaState alphaState, const VPoint& imageSize, const SaveUIOptions& saveOptions);
ScopedHeifImage CreateHeifImageGrayThirtyTwoBit(FormatRecordPtr formatRecord, AlphaState alphaState, const VPoint& imageSize, const SaveUIOptions& saveOptions);
ScopedHeifImage CreateHeifImageRGBEightBit(FormatRecordPtr formatRecord, AlphaState alphaState, const VPoint& imageSize, const SaveUIOptions& saveOptions);
ScopedHeifImage CreateHeifImageRGBSixteenBit(FormatRecordPtr formatRecord, AlphaState alphaState, const VPoint& imageSize, const SaveUIOptions& saveOptions);
ScopedHeifImage CreateHeifImageRGBThirtyTwoBit(FormatRecordPtr formatRecord, AlphaState alphaState, const VPoint& imageSize, const SaveUIOptions& saveOptions);

void ReadHeifImageGrayEightBit(const heif_image* image, AlphaState alphaState, const heif_color_profile_nclx* nclxProfile, FormatRecordPtr formatRecord);
void ReadHeifImageGraySixteenBit(const heif_image* image, AlphaState alphaState, const heif_color_profile_nclx* nclxProfile, FormatRecordPtr formatRecord);
void ReadHeifImageGrayThirtyTwoBit(const heif_image* image, AlphaState alphaState, const heif_color_profile_nclx* nclxProfile, const LoadUIOptions& loadOptions, FormatRecordPtr formatRecord);
void ReadHeifImageRGBEightBit(const heif_image* image, AlphaState alphaState, const heif_color_profile_nclx* nclxProfile, FormatRecordPtr formatRecord);
void ReadHeifImageRGBSixteenBit(const heif_image* image, AlphaState alphaState, const heif_color_profile_nclx* nclxProfile, FormatRecordPtr formatRecord);
void ReadHeifImageRGBThirtyTwoBit(const heif_image* image, AlphaState alphaState, const heif_color_profile_nclx* nclxProfile, const LoadUIOptions& loadOptions, FormatRecordPtr formatRecord);

#endif
