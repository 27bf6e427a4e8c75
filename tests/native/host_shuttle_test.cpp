// tests/native/host_shuttle_test.cpp -- drives the C++ host mirror (avif-format_b200/host/GpuRowShuttle.cpp) the way
// the plug-in's Write.cpp / Read.cpp would: a mock Photoshop host serves / collects row blocks through
// FormatRecord::advanceState, a mock libheif (oracle/shim/mock_heif.cpp) owns the planes, and every result is
// compared with the CPU oracle (liboracle.so).  Needs a B200.  Exit code 0 = all checks passed.
#include "GpuRowShuttle.h"
#include "../../include/avifgpu.h"
#include "../../oracle/avif_oracle.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

namespace
{
    struct MockHost
    {
        FormatRecord record{};
        BufferProcs procs{};
        const uint8_t* source = nullptr; // encode: rows the host hands out
        uint8_t* sink = nullptr;         // decode: rows the host receives
        int64_t stride = 0;
        int64_t payload = 0;
        int advanceCalls = 0;
        int abortAfterCalls = -1;
        int maxRowsSeen = 0;
    };
    MockHost* g_host = nullptr;

    OSErr Advance()
    {
        MockHost& h = *g_host;
        FormatRecord& r = h.record;
        h.advanceCalls++;
        const int top = r.theRect32.top, bottom = r.theRect32.bottom;
        h.maxRowsSeen = std::max(h.maxRowsSeen, bottom - top);
        for (int y = top; y < bottom; ++y)
        {
            uint8_t* hostRow = static_cast<uint8_t*>(r.data) + static_cast<int64_t>(y - top) * r.rowBytes;
            if (h.source) std::memcpy(hostRow, h.source + y * h.stride, static_cast<size_t>(h.payload));
            if (h.sink) std::memcpy(h.sink + y * h.stride, hostRow, static_cast<size_t>(h.payload));
        }
        return noErr;
    }
    Boolean Abort() { return g_host->abortAfterCalls >= 0 && g_host->advanceCalls >= g_host->abortAfterCalls; }
    void Progress(int32, int32) {}

    void InitHost(MockHost& h, int w, int hgt, int channels, int depth)
    {
        FormatRecord& r = h.record;
        r.planes = static_cast<int16>(channels);
        r.depth = static_cast<int16>(depth);
        r.imageSize32.h = w;
        r.imageSize32.v = hgt;
        r.HostSupports32BitCoordinates = 1;
        r.PluginUsing32BitCoordinates = 1;
        r.advanceState = Advance;
        r.abortProc = Abort;
        r.progressProc = Progress;
        r.bufferProcs = &h.procs;
        g_host = &h;
    }

    int g_failures = 0;
    void Expect(bool ok, const char* what)
    {
        std::printf("%s %s\n", ok ? "PASS" : "FAIL", what);
        if (!ok) ++g_failures;
    }

    bool PlaneEquals(const heif_image* image, heif_channel channel, const std::vector<uint16_t>& expected, int w, int h)
    {
        int stride = 0;
        const uint8_t* p = heif_image_get_plane_readonly(image, channel, &stride);
        if (!p) return false;
        for (int y = 0; y < h; ++y)
        {
            if (std::memcmp(p + static_cast<int64_t>(y) * stride, expected.data() + static_cast<size_t>(y) * w, static_cast<size_t>(w) * 2) != 0) return false;
        }
        return true;
    }

    void TestEncodeRgb32(int w, int h, ChromaSubsampling sub, bool withAlpha)
    {
        const int channels = withAlpha ? 4 : 3;
        std::mt19937 rng(1234 + w);
        std::uniform_real_distribution<float> dist(-0.1f, 1.3f);
        std::vector<float> rows(static_cast<size_t>(w) * h * channels);
        for (float& v : rows) v = dist(rng);

        MockHost host;
        InitHost(host, w, h, channels, 32);
        host.source = reinterpret_cast<const uint8_t*>(rows.data());
        host.stride = static_cast<int64_t>(w) * channels * 4;
        host.payload = host.stride;

        SaveUIOptions options{};
        options.chromaSubsampling = sub;
        options.imageBitDepth = ImageBitDepth::Twelve;
        options.hdrTransferFunction = ColorTransferFunction::PQ;
        options.pq.nominalPeakBrightness = 80;
        const AlphaState alpha = withAlpha ? AlphaState::Straight : AlphaState::None;
        const VPoint size{ h, w };
        avifgpu_host::SetRowsPerBlock(6);
        ScopedHeifImage image = CreateHeifImageRGBThirtyTwoBit(&host.record, alpha, size, options);

        avifgpu_encode_desc d{};
        d.struct_size = sizeof(d);
        d.width = w; d.height = h; d.host_depth = 32; d.host_channels = channels; d.alpha_state = static_cast<int>(alpha);
        d.image_bit_depth = 12; d.transfer = AVIFGPU_TRANSFER_PQ; d.pq_peak_nits = 80; d.layout = AVIFGPU_LAYOUT_PLANAR_YCBCR;
        d.chroma = sub == ChromaSubsampling::Yuv420 ? AVIFGPU_CHROMA_420 : sub == ChromaSubsampling::Yuv422 ? AVIFGPU_CHROMA_422 : AVIFGPU_CHROMA_444;
        d.nclx.present = 1; d.nclx.color_primaries = 9; d.nclx.transfer_characteristics = 16; d.nclx.matrix_coefficients = 9; d.nclx.full_range_flag = 1;
        const int cw = d.chroma == AVIFGPU_CHROMA_444 ? w : (w + 1) / 2;
        const int ch = d.chroma == AVIFGPU_CHROMA_420 ? (h + 1) / 2 : h;
        std::vector<uint16_t> y(static_cast<size_t>(w) * h), cb(static_cast<size_t>(cw) * ch), cr(cb.size()), a(y.size());
        avifgpu_planes planes{};
        planes.data[0] = y.data(); planes.stride[0] = w * 2;
        planes.data[1] = cb.data(); planes.stride[1] = cw * 2;
        planes.data[2] = cr.data(); planes.stride[2] = cw * 2;
        if (withAlpha) { planes.data[3] = a.data(); planes.stride[3] = w * 2; }
        const int status = avif_oracle_encode_image(&d, rows.data(), host.stride, &planes);
        char label[160];
        std::snprintf(label, sizeof(label), "encode RGB%s32 %dx%d chroma %d: oracle status, Y, Cb, Cr%s, multi-row blocks (%d rows/advance)",
                      withAlpha ? "A" : "", w, h, d.chroma, withAlpha ? ", A" : "", host.maxRowsSeen);
        bool ok = status == 0 && heif_image_get_colorspace(image.get()) == heif_colorspace_YCbCr &&
                  PlaneEquals(image.get(), heif_channel_Y, y, w, h) && PlaneEquals(image.get(), heif_channel_Cb, cb, cw, ch) &&
                  PlaneEquals(image.get(), heif_channel_Cr, cr, cw, ch) && (!withAlpha || PlaneEquals(image.get(), heif_channel_Alpha, a, w, h)) &&
                  host.maxRowsSeen > 1 && host.advanceCalls == (h + 5) / 6;
        Expect(ok, label);
    }

    void TestDecodeHlg(int w, int h)
    {
        std::mt19937 rng(77);
        heif_image* raw = nullptr;
        heif_image_create(w, h, heif_colorspace_YCbCr, heif_chroma_420, &raw);
        ScopedHeifImage image(raw);
        const int cw = (w + 1) / 2, ch = (h + 1) / 2;
        heif_image_add_plane(raw, heif_channel_Y, w, h, 10);
        heif_image_add_plane(raw, heif_channel_Cb, cw, ch, 10);
        heif_image_add_plane(raw, heif_channel_Cr, cw, ch, 10);
        avifgpu_planes planes{};
        const heif_channel channels[3] = { heif_channel_Y, heif_channel_Cb, heif_channel_Cr };
        for (int k = 0; k < 3; ++k)
        {
            int stride = 0;
            uint8_t* p = heif_image_get_plane(raw, channels[k], &stride);
            const int pw = k ? cw : w, ph = k ? ch : h;
            for (int y = 0; y < ph; ++y)
                for (int x = 0; x < pw; ++x)
                    reinterpret_cast<uint16_t*>(p + static_cast<int64_t>(y) * stride)[x] = static_cast<uint16_t>(rng() % 1024);
            planes.data[k] = p;
            planes.stride[k] = stride;
        }
        heif_color_profile_nclx nclx{};
        nclx.color_primaries = 9; nclx.transfer_characteristics = 18; nclx.matrix_coefficients = 9; nclx.full_range_flag = 1;
        LoadUIOptions load{};
        load.hlg.applyOOTF = true; load.hlg.displayGamma = 1.2f; load.hlg.nominalPeakBrightness = 1000; load.pq.nominalPeakBrightness = 80;

        std::vector<float> got(static_cast<size_t>(w) * h * 3), expected(got.size());
        MockHost host;
        InitHost(host, w, h, 3, 32);
        host.sink = reinterpret_cast<uint8_t*>(got.data());
        host.stride = static_cast<int64_t>(w) * 12;
        host.payload = host.stride;
        avifgpu_host::SetRowsPerBlock(10);
        ReadHeifImageRGBThirtyTwoBit(image.get(), AlphaState::None, &nclx, load, &host.record);

        avifgpu_decode_desc d{};
        d.struct_size = sizeof(d);
        d.width = w; d.height = h; d.colorspace = AVIFGPU_COLORSPACE_YCBCR; d.chroma = AVIFGPU_CHROMA_420; d.bit_depth = 10; d.host_depth = 32;
        d.nclx.present = 1; d.nclx.color_primaries = 9; d.nclx.transfer_characteristics = 18; d.nclx.matrix_coefficients = 9; d.nclx.full_range_flag = 1;
        d.hlg_apply_ootf = 1; d.hlg_display_gamma = 1.2f; d.hlg_peak_nits = 1000; d.pq_peak_nits = 80;
        const int status = avif_oracle_decode_image(&d, &planes, expected.data(), host.stride);
        Expect(status == 0 && std::memcmp(got.data(), expected.data(), got.size() * 4) == 0 && host.record.rowBytes == 0,
               "decode 10-bit HLG 4:2:0 -> RGB32f bit-identical to the oracle; formatRecord->data/rowBytes restored");
    }

    // A stand-in for the plug-in's lcms2 transform: something observable and exactly reproducible (halve every colour
    // sample; powers of two are exact in float32).
    struct HalvingTransform final : avifgpu_host::RowTransform
    {
        int channels;
        int rowsConverted = 0;
        explicit HalvingTransform(int c) : channels(c) {}
        void ConvertRow(void* row, uint32_t pixelsPerLine, uint32_t) override
        {
            float* v = static_cast<float*>(row);
            for (uint32_t x = 0; x < pixelsPerLine; ++x)
                for (int c = 0; c < 3; ++c) v[x * channels + c] *= 0.5f;
            ++rowsConverted;
        }
    };
    HalvingTransform* g_lastTransform = nullptr;
    int g_factoryCalls = 0;
    avifgpu_host::RowTransform* HalvingFactory(FormatRecordPtr, bool hasAlpha, int hostBits, ColorTransferFunction, bool, void*)
    {
        ++g_factoryCalls;
        g_lastTransform = hostBits == 32 ? new HalvingTransform(hasAlpha ? 4 : 3) : nullptr; // nullptr = "profile already matches"
        return g_lastTransform;
    }

    // A minimal ICC profile: RGB / XYZ, linear Rec.709 primaries (D65) adapted to D50 (the sRGB profile's colorants), identity
    // tone curves -- the profile Photoshop attaches to a 32-bit document in the "sRGB (linear)" working space.
    std::vector<unsigned char> MakeLinearRec709Profile()
    {
        const double colorants[3][3] = { { 0.43607, 0.22249, 0.01392 }, { 0.38515, 0.71687, 0.09708 }, { 0.14307, 0.06061, 0.71410 } }; // r, g, b columns as (X, Y, Z)
        std::vector<unsigned char> p(128, 0);
        auto put32 = [&](size_t at, uint32_t v) { p[at] = v >> 24; p[at + 1] = (v >> 16) & 0xff; p[at + 2] = (v >> 8) & 0xff; p[at + 3] = v & 0xff; };
        auto append32 = [&](uint32_t v) { p.push_back(v >> 24); p.push_back((v >> 16) & 0xff); p.push_back((v >> 8) & 0xff); p.push_back(v & 0xff); };
        std::memcpy(&p[12], "mntr", 4); std::memcpy(&p[16], "RGB ", 4); std::memcpy(&p[20], "XYZ ", 4); std::memcpy(&p[36], "acsp", 4);
        const char* names[6] = { "rXYZ", "gXYZ", "bXYZ", "rTRC", "gTRC", "bTRC" };
        append32(6);
        const uint32_t dataStart = 128 + 4 + 6 * 12;
        for (int i = 0; i < 6; ++i)
        {
            for (int c = 0; c < 4; ++c) p.push_back(static_cast<unsigned char>(names[i][c]));
            append32(i < 3 ? dataStart + 20 * i : dataStart + 60 + 12 * (i - 3));
            append32(i < 3 ? 20 : 12);
        }
        for (int i = 0; i < 3; ++i)
        {
            for (int c = 0; c < 4; ++c) p.push_back(static_cast<unsigned char>("XYZ "[c]));
            append32(0);
            for (int k = 0; k < 3; ++k) append32(static_cast<uint32_t>(static_cast<int32_t>(colorants[i][k] * 65536.0 + 0.5)));
        }
        for (int i = 0; i < 3; ++i)
        {
            for (int c = 0; c < 4; ++c) p.push_back(static_cast<unsigned char>("curv"[c]));
            append32(0);
            append32(0); // no entries: the identity curve
        }
        put32(0, static_cast<uint32_t>(p.size()));
        return p;
    }

    // The ICC seam (WriteHeifImage.cpp:1015-1036, ColorProfileConversion.cpp:107-120, HostMetadata.cpp:63-69).
    void TestColorProfileStep()
    {
        const int w = 40, h = 22;
        std::mt19937 rng(4321);
        std::uniform_real_distribution<float> dist(0.0f, 1.2f);
        std::vector<float> rows(static_cast<size_t>(w) * h * 3);
        for (float& v : rows) v = dist(rng);
        static const unsigned char fakeProfile[16] = { 1 };

        auto save = [&](bool withProfile, bool keepProfile, ColorTransferFunction transfer, MockHost& host) -> ScopedHeifImage
        {
            InitHost(host, w, h, 3, 32);
            host.source = reinterpret_cast<const uint8_t*>(rows.data());
            host.stride = host.payload = static_cast<int64_t>(w) * 12;
            if (withProfile)
            {
                host.record.canUseICCProfiles = 1;
                host.record.iCCprofileData = const_cast<unsigned char*>(fakeProfile);
                host.record.iCCprofileSize = sizeof(fakeProfile);
            }
            SaveUIOptions options{};
            options.chromaSubsampling = ChromaSubsampling::Yuv444;
            options.imageBitDepth = ImageBitDepth::Twelve;
            options.hdrTransferFunction = transfer;
            options.pq.nominalPeakBrightness = 80;
            options.keepColorProfile = keepProfile;
            avifgpu_host::SetRowsPerBlock(8);
            return CreateHeifImageRGBThirtyTwoBit(&host.record, AlphaState::None, VPoint{ h, w }, options);
        };

        // 1. a profile that may need converting, and nobody to do it: an error -- never unconverted pixels
        avifgpu_host::SetRowTransformFactory(nullptr, nullptr);
        {
            MockHost host;
            bool refused = false;
            try { save(true, false, ColorTransferFunction::PQ, host); }
            catch (const OSErrException& e) { refused = e.GetErrorCode() == formatBadParameters; }
            Expect(refused && host.advanceCalls == 0, "document profile + PQ save without a row transform -> OSErrException(formatBadParameters) before any row is read");
        }
        // 2. the reference's "no conversion" cases need no transform: no profile; keepColorProfile with the clip transfer
        {
            MockHost host;
            bool ok = true;
            try { save(false, false, ColorTransferFunction::PQ, host); save(true, true, ColorTransferFunction::Clip, host); }
            catch (...) { ok = false; }
            Expect(ok, "no document profile, or keepColorProfile with the clip transfer: no transform asked for (ColorProfileConversion.cpp:107)");
        }
        // 3. with a factory the transform runs over every staged row before the conversion
        avifgpu_host::SetRowTransformFactory(HalvingFactory, nullptr);
        {
            MockHost host;
            g_factoryCalls = 0;
            ScopedHeifImage image = save(true, true, ColorTransferFunction::PQ, host); // PQ: may require conversion even when keeping the profile
            std::vector<float> halved(rows);
            for (float& v : halved) v *= 0.5f;
            avifgpu_encode_desc d{};
            d.struct_size = sizeof(d);
            d.width = w; d.height = h; d.host_depth = 32; d.host_channels = 3; d.image_bit_depth = 12; d.transfer = AVIFGPU_TRANSFER_PQ;
            d.pq_peak_nits = 80; d.layout = AVIFGPU_LAYOUT_PLANAR_YCBCR; d.chroma = AVIFGPU_CHROMA_444;
            d.nclx.present = 1; d.nclx.color_primaries = 9; d.nclx.transfer_characteristics = 16; d.nclx.matrix_coefficients = 9; d.nclx.full_range_flag = 1;
            std::vector<uint16_t> y(static_cast<size_t>(w) * h), cb(y.size()), cr(y.size());
            avifgpu_planes planes{};
            planes.data[0] = y.data(); planes.data[1] = cb.data(); planes.data[2] = cr.data();
            planes.stride[0] = planes.stride[1] = planes.stride[2] = w * 2;
            const int status = avif_oracle_encode_image(&d, halved.data(), static_cast<int64_t>(w) * 12, &planes);
            const bool same = status == 0 && PlaneEquals(image.get(), heif_channel_Y, y, w, h) && PlaneEquals(image.get(), heif_channel_Cb, cb, w, h) &&
                              PlaneEquals(image.get(), heif_channel_Cr, cr, w, h);
            Expect(same && g_factoryCalls == 1, "row transform applied to every staged row, once, before the conversion (planes == oracle of the transformed rows)");
        }
        avifgpu_host::SetRowTransformFactory(nullptr, nullptr);
        // 4. a linear-light document in a matrix profile (linear Rec.709 primaries): no host transform needed at all -- the
        //    shuttle derives the 3x3 from the profile and the GPU applies it ahead of the conversion
        {
            const std::vector<unsigned char> profile = MakeLinearRec709Profile();
            MockHost host;
            InitHost(host, w, h, 3, 32);
            host.source = reinterpret_cast<const uint8_t*>(rows.data());
            host.stride = host.payload = static_cast<int64_t>(w) * 12;
            host.record.canUseICCProfiles = 1;
            host.record.iCCprofileData = const_cast<unsigned char*>(profile.data());
            host.record.iCCprofileSize = static_cast<int32>(profile.size());
            SaveUIOptions options{};
            options.chromaSubsampling = ChromaSubsampling::Yuv444;
            options.imageBitDepth = ImageBitDepth::Twelve;
            options.hdrTransferFunction = ColorTransferFunction::PQ;
            options.pq.nominalPeakBrightness = 80;
            avifgpu_host::SetRowsPerBlock(8);
            ScopedHeifImage image = CreateHeifImageRGBThirtyTwoBit(&host.record, AlphaState::None, VPoint{ h, w }, options);
            avifgpu_encode_desc d{};
            d.struct_size = sizeof(d);
            d.width = w; d.height = h; d.host_depth = 32; d.host_channels = 3; d.image_bit_depth = 12; d.transfer = AVIFGPU_TRANSFER_PQ;
            d.pq_peak_nits = 80; d.layout = AVIFGPU_LAYOUT_PLANAR_YCBCR; d.chroma = AVIFGPU_CHROMA_444;
            d.nclx.present = 1; d.nclx.color_primaries = 9; d.nclx.transfer_characteristics = 16; d.nclx.matrix_coefficients = 9; d.nclx.full_range_flag = 1;
            int32_t same = 1;
            const int parsed = avifgpu_icc_to_rec2020_linear_matrix(profile.data(), profile.size(), d.row_matrix, &same);
            d.row_matrix_enabled = 1;
            std::vector<uint16_t> y(static_cast<size_t>(w) * h), cb(y.size()), cr(y.size());
            avifgpu_planes planes{};
            planes.data[0] = y.data(); planes.data[1] = cb.data(); planes.data[2] = cr.data();
            planes.stride[0] = planes.stride[1] = planes.stride[2] = w * 2;
            const int status = avif_oracle_encode_image(&d, rows.data(), static_cast<int64_t>(w) * 12, &planes);
            const bool matrixLooksRight = d.row_matrix[0] > 0.62f && d.row_matrix[0] < 0.635f && d.row_matrix[4] > 0.91f; // BT.2087: 0.6274 / 0.9195
            Expect(parsed == 0 && same == 0 && matrixLooksRight && status == 0 && PlaneEquals(image.get(), heif_channel_Y, y, w, h) &&
                       PlaneEquals(image.get(), heif_channel_Cb, cb, w, h) && PlaneEquals(image.get(), heif_channel_Cr, cr, w, h),
                   "matrix profile (linear Rec.709) on an HDR save: 3x3 derived from the ICC bytes, applied on the GPU, planes == oracle with that matrix");
        }
    }

    // Staging is bounded in bytes, not rows (a 300 000-pixel RGBA32f row is 4.8 MB).
    void TestStagingBudget()
    {
        const int w = 4096, h = 64;
        std::vector<uint8_t> rows(static_cast<size_t>(w) * h * 4, 9);
        MockHost host;
        InitHost(host, w, h, 4, 8);
        host.source = rows.data();
        host.stride = host.payload = static_cast<int64_t>(w) * 4;
        SaveUIOptions options{};
        options.chromaSubsampling = ChromaSubsampling::Yuv420;
        options.imageBitDepth = ImageBitDepth::Eight;
        options.hdrTransferFunction = ColorTransferFunction::Clip;
        avifgpu_host::SetRowsPerBlock(4096);
        avifgpu_host::SetStagingBudgetBytes(5 * 16384); // five rows' worth -> blocks of four rows
        CreateHeifImageRGBEightBit(&host.record, AlphaState::Straight, VPoint{ h, w }, options);
        const avifgpu_host::ShuttleTimes times = avifgpu_host::LastShuttleTimes();
        Expect(host.maxRowsSeen == 4 && times.rowsPerBlock == 4 && times.blocks == 16 && times.total > 0.0,
               "staging budget in bytes bounds the block (4 rows of 16 KiB under an 80 KiB budget), 16 blocks for 64 rows");
        avifgpu_host::SetStagingBudgetBytes(64ll << 20);
    }

    void TestErrors()
    {
        // user cancel between blocks -> OSErrException(userCanceledErr), as WriteHeifImage.cpp:208-211
        {
            const int w = 32, h = 40;
            std::vector<uint8_t> rows(static_cast<size_t>(w) * h * 3, 7);
            MockHost host;
            InitHost(host, w, h, 3, 8);
            host.source = rows.data();
            host.stride = host.payload = w * 3;
            host.abortAfterCalls = 2;
            SaveUIOptions options{};
            options.chromaSubsampling = ChromaSubsampling::Yuv420;
            options.imageBitDepth = ImageBitDepth::Eight;
            options.hdrTransferFunction = ColorTransferFunction::Clip;
            bool canceled = false;
            avifgpu_host::SetRowsPerBlock(8);
            try { CreateHeifImageRGBEightBit(&host.record, AlphaState::None, VPoint{ h, w }, options); }
            catch (const OSErrException& e) { canceled = e.GetErrorCode() == userCanceledErr; }
            Expect(canceled && host.advanceCalls == 2, "abortProc between row blocks -> OSErrException(userCanceledErr)");
        }
        // nclx == nullptr on a 32-bit read -> std::runtime_error("The nclxProfile is null."), ReadHeifImage.cpp:956-959
        {
            MockHost host;
            InitHost(host, 4, 4, 3, 32);
            LoadUIOptions load{};
            bool thrown = false;
            try { ReadHeifImageRGBThirtyTwoBit(nullptr, AlphaState::None, nullptr, load, &host.record); }
            catch (const std::runtime_error& e) { thrown = std::strcmp(e.what(), "The nclxProfile is null.") == 0; }
            Expect(thrown, "null nclx -> std::runtime_error(\"The nclxProfile is null.\")");
        }
        // unsupported depth enum -> OSErrException(formatCannotRead), WriteHeifImage.cpp:56-57
        {
            MockHost host;
            InitHost(host, 4, 4, 1, 8);
            SaveUIOptions options{};
            options.imageBitDepth = static_cast<ImageBitDepth>(9);
            bool thrown = false;
            try { CreateHeifImageGrayEightBit(&host.record, AlphaState::None, VPoint{ 4, 4 }, options); }
            catch (const OSErrException& e) { thrown = e.GetErrorCode() == formatCannotRead; }
            Expect(thrown, "bad ImageBitDepth -> OSErrException(formatCannotRead)");
        }
    }
}

int main()
{
    try
    {
        TestEncodeRgb32(70, 21, ChromaSubsampling::Yuv420, false);
        TestEncodeRgb32(64, 16, ChromaSubsampling::Yuv422, true);
        TestEncodeRgb32(33, 9, ChromaSubsampling::Yuv444, false);
        TestDecodeHlg(75, 33);
        TestColorProfileStep();
        TestStagingBudget();
        TestErrors();
    }
    catch (const std::exception& e)
    {
        std::printf("FAIL unexpected exception: %s\n", e.what());
        ++g_failures;
    }
    avifgpu_host::ReleaseSharedContext();
    std::printf("%s\n", g_failures == 0 ? "ALL PASSED" : "SOME FAILED");
    return g_failures == 0 ? 0 : 1;
}
