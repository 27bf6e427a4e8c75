/*
 * oracle/shim/mock_heif.cpp -- TEST INFRASTRUCTURE ONLY.
 * A heap-backed stand-in for the few heif_image_* calls the reference's row shuttle makes
 * (WriteHeifImage.cpp:32-39,175-194,639-646; ReadHeifImage.cpp:104-111 and siblings).  Rows are padded the
 * way libheif pads them (stride rounded up to 16 bytes plus one extra 16-byte lane) so stride handling is
 * exercised.  Written from the public API description; not derived from libheif sources.
 */
#include "libheif/heif.h"
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>

namespace
{
    struct MockPlane
    {
        int width = 0;
        int height = 0;
        int bitDepth = 0;
        int stride = 0;
        uint8_t* memory = nullptr;
    };

    const heif_error kOk = { heif_error_Ok, heif_suberror_Unspecified, "Success" };
}

struct heif_image
{
    int width;
    int height;
    heif_colorspace colorspace;
    heif_chroma chroma;
    std::map<int, MockPlane> planes;
};

extern "C" {

heif_error heif_image_create(int width, int height, heif_colorspace colorspace, heif_chroma chroma,
                             heif_image** out_image)
{
    heif_image* image = new (std::nothrow) heif_image();
    if (image == nullptr)
    {
        return heif_error{ heif_error_Memory_allocation_error, heif_suberror_Unspecified, "out of memory" };
    }
    image->width = width;
    image->height = height;
    image->colorspace = colorspace;
    image->chroma = chroma;
    *out_image = image;
    return kOk;
}

heif_error heif_image_add_plane(heif_image* image, heif_channel channel, int width, int height, int bit_depth)
{
    int samplesPerPixel = 1;
    if (channel == heif_channel_interleaved)
    {
        switch (image->chroma)
        {
        case heif_chroma_interleaved_RGB:
        case heif_chroma_interleaved_RRGGBB_BE:
        case heif_chroma_interleaved_RRGGBB_LE:
            samplesPerPixel = 3;
            break;
        case heif_chroma_interleaved_RGBA:
        case heif_chroma_interleaved_RRGGBBAA_BE:
        case heif_chroma_interleaved_RRGGBBAA_LE:
            samplesPerPixel = 4;
            break;
        default:
            return heif_error{ heif_error_Usage_error, heif_suberror_Invalid_parameter_value, "bad chroma" };
        }
    }

    const int bytesPerSample = (bit_depth + 7) / 8;
    MockPlane plane;
    plane.width = width;
    plane.height = height;
    plane.bitDepth = bit_depth;
    plane.stride = ((width * samplesPerPixel * bytesPerSample + 15) / 16) * 16 + 16;
    const size_t bytes = static_cast<size_t>(plane.stride) * static_cast<size_t>(height > 0 ? height : 1);
    plane.memory = static_cast<uint8_t*>(std::malloc(bytes));
    if (plane.memory == nullptr)
    {
        return heif_error{ heif_error_Memory_allocation_error, heif_suberror_Unspecified, "out of memory" };
    }
    std::memset(plane.memory, 0xCD, bytes);

    auto existing = image->planes.find(channel);
    if (existing != image->planes.end())
    {
        std::free(existing->second.memory);
    }
    image->planes[channel] = plane;
    return kOk;
}

uint8_t* heif_image_get_plane(heif_image* image, heif_channel channel, int* out_stride)
{
    auto it = image->planes.find(channel);
    if (it == image->planes.end())
    {
        if (out_stride) *out_stride = 0;
        return nullptr;
    }
    if (out_stride) *out_stride = it->second.stride;
    return it->second.memory;
}

const uint8_t* heif_image_get_plane_readonly(const heif_image* image, heif_channel channel, int* out_stride)
{
    return heif_image_get_plane(const_cast<heif_image*>(image), channel, out_stride);
}

int heif_image_get_bits_per_pixel_range(const heif_image* image, heif_channel channel)
{
    auto it = image->planes.find(channel);
    return it == image->planes.end() ? -1 : it->second.bitDepth;
}

heif_chroma heif_image_get_chroma_format(const heif_image* image) { return image->chroma; }
heif_colorspace heif_image_get_colorspace(const heif_image* image) { return image->colorspace; }

int heif_image_get_width(const heif_image* image, heif_channel channel)
{
    auto it = image->planes.find(channel);
    return it == image->planes.end() ? -1 : it->second.width;
}

int heif_image_get_height(const heif_image* image, heif_channel channel)
{
    auto it = image->planes.find(channel);
    return it == image->planes.end() ? -1 : it->second.height;
}

void heif_image_release(const heif_image* image)
{
    if (image == nullptr) return;
    for (auto& entry : image->planes)
    {
        std::free(entry.second.memory);
    }
    delete image;
}

void heif_context_free(heif_context*) {}
void heif_encoder_release(heif_encoder*) {}
void heif_encoding_options_free(heif_encoding_options*) {}
void heif_image_handle_release(const heif_image_handle*) {}
void heif_nclx_color_profile_free(heif_color_profile_nclx*) {}

} // extern "C"
