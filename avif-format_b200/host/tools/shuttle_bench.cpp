// shuttle_bench.cpp -- times the plug-in's row shuttle (GpuRowShuttle.cpp: CreateHeifImageRGBThirtyTwoBit /
// CreateHeifImageRGBSixteenBit exactly as Write.cpp:303-336 calls them) end to end on this box: a mock Photoshop host
// serves row blocks through FormatRecord::advanceState into the shuttle's pinned buffers, the planes belong to a
// heap-backed heif_image (pageable memory with padded strides, like libheif's), the GPU path is whatever the shuttle
// uses.  bench.py runs it for its `e2e_shuttle` figure.
//
//   shuttle_bench <c2|c3|c4> <width> <height> <steps> <copy|resident> <fresh|warm> [device ...]
//
//   c3        the read side: ReadHeifImageRGBThirtyTwoBit on a 10-bit HLG 4:2:0 image (Read.cpp:587-630); the host takes the
//             rows it is offered (copy: memcpy into a pageable frame; resident: leaves them in the staging buffer)
//   copy      advanceState memcpy's the requested rows out of a pageable source frame (what any real host at least does)
//   resident  advanceState leaves the staging buffers as they are after their first fill: the host's own cost removed,
//             what remains is the shuttle + the library (the figure to hold against bench.py's e2e)
//   fresh     every image's planes are new malloc'd memory, as libheif allocates them: the first write to each page
//             faults it in (the kernel zeroes 100 MB per 8K frame) -- a cost of the allocation, paid by whoever writes
//             the planes first, CPU loop or GPU shuttle alike
//   warm      the planes come from an arena whose pages have been touched before (an allocator that recycles memory)
// Prints one JSON object.
#include "../GpuRowShuttle.h"
#include "../../../include/avifgpu.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

// ---- a heap-backed heif_image (only what the shuttle calls) -----------------------------------------------------------
namespace
{
    // "warm" planes: a bump arena over memory that has been written once.
    uint8_t* g_arena = nullptr;
    size_t g_arenaBytes = 0, g_arenaUsed = 0;
    bool g_warm = false;
}

struct heif_image
{
    int width, height;
    heif_colorspace colorspace;
    heif_chroma chroma;
    struct Plane { int width, height, depth, stride; uint8_t* memory; };
    std::map<int, Plane> planes;
};

extern "C" {
heif_error heif_image_create(int width, int height, heif_colorspace colorspace, heif_chroma chroma, heif_image** out)
{
    *out = new heif_image{ width, height, colorspace, chroma, {} };
    return heif_error{ heif_error_Ok, heif_suberror_Unspecified, "Success" };
}
heif_error heif_image_add_plane(heif_image* image, heif_channel channel, int width, int height, int depth)
{
    heif_image::Plane p{ width, height, depth, 0, nullptr };
    p.stride = ((width * (depth > 8 ? 2 : 1) + 15) / 16 + 1) * 16; // padded rows, like libheif's
    const size_t bytes = static_cast<size_t>(p.stride) * (height > 0 ? height : 1);
    if (g_warm && g_arenaUsed + bytes + 64 <= g_arenaBytes)
    {
        p.memory = g_arena + g_arenaUsed;
        g_arenaUsed += (bytes + 63) & ~static_cast<size_t>(63);
    }
    else
    {
        p.memory = static_cast<uint8_t*>(std::malloc(bytes));
    }
    image->planes[channel] = p;
    return heif_error{ p.memory ? heif_error_Ok : heif_error_Memory_allocation_error, heif_suberror_Unspecified, "" };
}
uint8_t* heif_image_get_plane(heif_image* image, heif_channel channel, int* stride)
{
    auto it = image->planes.find(channel);
    if (it == image->planes.end()) return nullptr;
    if (stride) *stride = it->second.stride;
    return it->second.memory;
}
const uint8_t* heif_image_get_plane_readonly(const heif_image* image, heif_channel channel, int* stride)
{
    return heif_image_get_plane(const_cast<heif_image*>(image), channel, stride);
}
int heif_image_get_bits_per_pixel_range(const heif_image* image, heif_channel channel)
{
    auto it = image->planes.find(channel);
    return it == image->planes.end() ? -1 : it->second.depth;
}
heif_chroma heif_image_get_chroma_format(const heif_image* image) { return image->chroma; }
heif_colorspace heif_image_get_colorspace(const heif_image* image) { return image->colorspace; }
void heif_image_release(const heif_image* image)
{
    for (auto& kv : image->planes)
    {
        if (!(kv.second.memory >= g_arena && kv.second.memory < g_arena + g_arenaBytes)) std::free(kv.second.memory);
    }
    g_arenaUsed = 0; // one image at a time
    delete image;
}
}

namespace
{
    struct Host
    {
        FormatRecord record{};
        BufferProcs procs{};
        const uint8_t* source = nullptr;
        uint8_t* sink = nullptr;
        int64_t stride = 0;
        bool copy = true;
    };
    Host* g_host = nullptr;
    std::vector<void*> g_filled; // resident mode: staging buffers that already hold rows (the shuttle keeps its buffers between calls)

    OSErr Advance()
    {
        Host& h = *g_host;
        FormatRecord& r = h.record;
        const int top = r.theRect32.top, bottom = r.theRect32.bottom;
        if (!h.copy)
        {
            for (void* p : g_filled)
                if (p == r.data) return noErr;
            g_filled.push_back(r.data);
        }
        for (int y = top; y < bottom; ++y)
        {
            std::memcpy(static_cast<uint8_t*>(r.data) + static_cast<int64_t>(y - top) * r.rowBytes, h.source + static_cast<int64_t>(y) * h.stride,
                        static_cast<size_t>(h.stride));
        }
        return noErr;
    }

    // the read side: the plug-in hands rows [top, bottom) to the host
    OSErr Take()
    {
        Host& h = *g_host;
        FormatRecord& r = h.record;
        if (!h.copy || r.data == nullptr)
        {
            return noErr;
        }
        const int top = r.theRect32.top, bottom = r.theRect32.bottom;
        for (int y = top; y < bottom; ++y)
        {
            std::memcpy(h.sink + static_cast<int64_t>(y) * h.stride, static_cast<const uint8_t*>(r.data) + static_cast<int64_t>(y - top) * r.rowBytes,
                        static_cast<size_t>(h.stride));
        }
        return noErr;
    }
    Boolean Abort() { return 0; }
    void Progress(int32, int32) {}
}

int main(int argc, char** argv)
{
    if (argc < 7)
    {
        std::fprintf(stderr, "usage: shuttle_bench <c2|c3|c4> <width> <height> <steps> <copy|resident> <fresh|warm> [device ...]\n");
        return 2;
    }
    const std::string workload = argv[1];
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), steps = std::atoi(argv[4]);
    const bool copy = std::string(argv[5]) == "copy";
    g_warm = std::string(argv[6]) == "warm";
    std::vector<int32_t> devices;
    for (int i = 7; i < argc; ++i) devices.push_back(std::atoi(argv[i]));
    const bool c2 = workload == "c2";
    const bool c3 = workload == "c3";
    const int channels = (c2 || c3) ? 3 : 4;
    const int depth = (c2 || c3) ? 32 : 16;
    const int64_t stride = static_cast<int64_t>(w) * channels * (depth / 8);

    if (g_warm)
    {
        g_arenaBytes = static_cast<size_t>(w) * h * 8 + (64u << 20); // every plane of any layout, with padding
        g_arena = static_cast<uint8_t*>(std::malloc(g_arenaBytes));
        std::memset(g_arena, 1, g_arenaBytes);
    }

    // synthetic frame: finite, in range, different everywhere (a 64-bit LCG), pageable memory like a host's tiles
    std::vector<uint8_t> frame(static_cast<size_t>(stride) * h);
    uint64_t state = 0x9e3779b97f4a7c15ull;
    if (c3)
    {
        std::memset(frame.data(), 0, frame.size()); // the host's frame: touched once, then written by Take()
    }
    else if (c2)
    {
        float* v = reinterpret_cast<float*>(frame.data());
        for (size_t i = 0; i < frame.size() / 4; ++i)
        {
            state = state * 6364136223846793005ull + 1442695040888963407ull;
            v[i] = static_cast<float>(state >> 40) * (1.0f / 16777216.0f);
        }
    }
    else
    {
        uint16_t* v = reinterpret_cast<uint16_t*>(frame.data());
        for (size_t i = 0; i < frame.size() / 2; ++i)
        {
            state = state * 6364136223846793005ull + 1442695040888963407ull;
            v[i] = static_cast<uint16_t>((state >> 33) % 32769u);
        }
    }

    try
    {
        if (devices.size() > 1)
        {
            avifgpu_host::UseDevices(devices.data(), static_cast<int32_t>(devices.size()));
        }
        SaveUIOptions options{};
        options.chromaSubsampling = c2 ? ChromaSubsampling::Yuv420 : ChromaSubsampling::Yuv422;
        options.imageBitDepth = c2 ? ImageBitDepth::Twelve : ImageBitDepth::Ten;
        options.hdrTransferFunction = c2 ? ColorTransferFunction::PQ : ColorTransferFunction::Clip;
        options.pq.nominalPeakBrightness = 80;
        options.keepColorProfile = true;
        const AlphaState alpha = c2 ? AlphaState::None : AlphaState::Straight;

        double best = 1e30, total = 0.0, hostSeconds = 0.0;
        avifgpu_host::ShuttleTimes times{};
        // c3: the decoded image libheif would hand over -- 10-bit Y, Cb, Cr planes (4:2:0) in pageable memory, padded strides
        heif_image* decoded = nullptr;
        heif_color_profile_nclx nclx{};
        LoadUIOptions load{};
        if (c3)
        {
            const bool arena = g_warm;
            g_warm = false; // the source planes are libheif's own allocation, never the recycled arena
            heif_image_create(w, h, heif_colorspace_YCbCr, heif_chroma_420, &decoded);
            const int cw = (w + 1) / 2, ch = (h + 1) / 2;
            heif_image_add_plane(decoded, heif_channel_Y, w, h, 10);
            heif_image_add_plane(decoded, heif_channel_Cb, cw, ch, 10);
            heif_image_add_plane(decoded, heif_channel_Cr, cw, ch, 10);
            g_warm = arena;
            const heif_channel planeChannels[3] = { heif_channel_Y, heif_channel_Cb, heif_channel_Cr };
            for (int k = 0; k < 3; ++k)
            {
                int planeStride = 0;
                uint8_t* p = heif_image_get_plane(decoded, planeChannels[k], &planeStride);
                const int pw = k ? cw : w, ph = k ? ch : h;
                for (int y = 0; y < ph; ++y)
                {
                    uint16_t* row = reinterpret_cast<uint16_t*>(p + static_cast<int64_t>(y) * planeStride);
                    for (int x = 0; x < pw; ++x)
                    {
                        state = state * 6364136223846793005ull + 1442695040888963407ull;
                        row[x] = static_cast<uint16_t>((state >> 33) & 1023u);
                    }
                }
            }
            nclx.color_primaries = 9; nclx.transfer_characteristics = 18; nclx.matrix_coefficients = 9; nclx.full_range_flag = 1; // BT.2020 / HLG / NCL / full
            load.hlg.applyOOTF = true; load.hlg.displayGamma = 1.2f; load.hlg.nominalPeakBrightness = 1000; load.pq.nominalPeakBrightness = 80;
        }
        for (int step = -1; step < steps; ++step) // step -1 = warm-up (allocations, table state)
        {
            Host host;
            host.record.planes = static_cast<int16>(channels);
            host.record.depth = static_cast<int16>(depth);
            host.record.imageSize32.h = w;
            host.record.imageSize32.v = h;
            host.record.HostSupports32BitCoordinates = 1;
            host.record.PluginUsing32BitCoordinates = 1;
            host.record.advanceState = c3 ? Take : Advance;
            host.record.abortProc = Abort;
            host.record.progressProc = Progress;
            host.record.bufferProcs = &host.procs;
            host.source = frame.data();
            host.sink = frame.data();
            host.stride = stride;
            host.copy = copy;
            g_host = &host;
            const auto start = std::chrono::steady_clock::now();
            ScopedHeifImage image;
            if (c3)
            {
                ReadHeifImageRGBThirtyTwoBit(decoded, AlphaState::None, &nclx, load, &host.record);
            }
            else
            {
                image = c2 ? CreateHeifImageRGBThirtyTwoBit(&host.record, alpha, VPoint{ h, w }, options)
                           : CreateHeifImageRGBSixteenBit(&host.record, alpha, VPoint{ h, w }, options);
            }
            const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
            if (step >= 0)
            {
                total += seconds;
                best = seconds < best ? seconds : best;
                times = avifgpu_host::LastShuttleTimes();
                hostSeconds += times.host;
            }
        }
        if (decoded != nullptr)
        {
            heif_image_release(decoded);
        }
        const double pixels = static_cast<double>(w) * h;
        std::printf("{\"workload\": \"%s\", \"width\": %d, \"height\": %d, \"steps\": %d, \"host\": \"%s\", \"gpus\": %d, "
                    "\"seconds_per_image\": %.6f, \"best_seconds\": %.6f, \"host_seconds_per_image\": %.6f, \"gpx_s\": %.4f, \"best_gpx_s\": %.4f, "
                    "\"blocks\": %d, \"rows_per_block\": %d, \"planes\": \"pageable (malloc), padded strides, %s\"}\n",
                    workload.c_str(), w, h, steps, copy ? "copy" : "resident", devices.size() > 1 ? static_cast<int>(devices.size()) : 1,
                    total / steps, best, hostSeconds / steps, pixels * steps / total / 1e9, pixels / best / 1e9, times.blocks, times.rowsPerBlock,
                    g_warm ? "pages already touched (recycling allocator)" : "fresh pages per image (first-touch faults inside the timed call)");
    }
    catch (const OSErrException& e)
    {
        std::printf("{\"error\": \"OSErr %d\"}\n", static_cast<int>(e.GetErrorCode()));
        return 1;
    }
    catch (const std::exception& e)
    {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
    avifgpu_host::ReleaseSharedContext();
    return 0;
}
