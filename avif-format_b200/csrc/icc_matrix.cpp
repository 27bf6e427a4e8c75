// icc_matrix.cpp -- the colour-profile step of the float save path for the profiles that are nothing but a matrix
// (SURVEY.md 8f-3).  The reference runs every host row through lcms2 before the transfer curve
// (ColorProfileConversion::ConvertRow, ColorProfileConversion.cpp:159-186; built at WriteHeifImage.cpp:1015 for 32-bit
// documents): document profile -> linear Rec.2020 (InitializeForRec2020Conversion, :240-266, perceptual intent + black
// point compensation).  Photoshop's 32-bit documents are linear-light, so for a matrix / TRC display profile with
// identity tone curves that whole transform is ONE 3x3 matrix:
//     M = inverse(Rec.2020 colorants adapted to the PCS white) x (document colorants: rXYZ gXYZ bXYZ, already PCS-adapted)
// (both profiles have a zero black point, so black point compensation is the identity).  This file parses just enough of
// an ICC profile to recognise that case and builds M in binary64; the GPU applies it per pixel in binary32 as a prologue
// of the conversion (avifgpu_encode_desc.row_matrix).  PARITY UNPINNED: lcms2 is not in the reference tree; its float
// pipeline evaluates the same chain with its own staging and rounding, so outputs can differ in the last float bits.
// Anything else -- LUT profiles, non-linear tone curves, non-RGB -- is reported as unsupported and stays with the host's
// lcms2 (GpuRowShuttle's RowTransform).
#include "../../include/avifgpu.h"

#include <cmath>
#include <cstring>

namespace
{

uint32_t ReadU32(const uint8_t* p) { return (static_cast<uint32_t>(p[0]) << 24) | (static_cast<uint32_t>(p[1]) << 16) | (static_cast<uint32_t>(p[2]) << 8) | p[3]; }
double ReadS15Fixed16(const uint8_t* p) { return static_cast<double>(static_cast<int32_t>(ReadU32(p))) / 65536.0; }
constexpr uint32_t Tag(char a, char b, char c, char d)
{
    return (static_cast<uint32_t>(static_cast<uint8_t>(a)) << 24) | (static_cast<uint32_t>(static_cast<uint8_t>(b)) << 16) |
           (static_cast<uint32_t>(static_cast<uint8_t>(c)) << 8) | static_cast<uint32_t>(static_cast<uint8_t>(d));
}

struct Profile
{
    const uint8_t* data;
    size_t size;
    const uint8_t* Find(uint32_t signature, uint32_t* outSize) const
    {
        const uint32_t count = ReadU32(data + 128);
        if (count > 1024 || 132 + static_cast<size_t>(count) * 12 > size)
        {
            return nullptr;
        }
        for (uint32_t i = 0; i < count; ++i)
        {
            const uint8_t* entry = data + 132 + static_cast<size_t>(i) * 12;
            if (ReadU32(entry) == signature)
            {
                const uint32_t offset = ReadU32(entry + 4), bytes = ReadU32(entry + 8);
                if (static_cast<uint64_t>(offset) + bytes > size || bytes < 8)
                {
                    return nullptr;
                }
                *outSize = bytes;
                return data + offset;
            }
        }
        return nullptr;
    }
};

bool ReadXyz(const Profile& profile, uint32_t signature, double out[3])
{
    uint32_t bytes = 0;
    const uint8_t* tag = profile.Find(signature, &bytes);
    if (tag == nullptr || bytes < 20 || ReadU32(tag) != Tag('X', 'Y', 'Z', ' '))
    {
        return false;
    }
    for (int i = 0; i < 3; ++i)
    {
        out[i] = ReadS15Fixed16(tag + 8 + 4 * i);
    }
    return true;
}

// True when the tone curve is the identity: 'curv' with no entries, 'curv' with one entry = gamma 1.0, a two-point 'curv'
// from 0 to 65535, or 'para' type 0 with g = 1.
bool CurveIsLinear(const Profile& profile, uint32_t signature)
{
    uint32_t bytes = 0;
    const uint8_t* tag = profile.Find(signature, &bytes);
    if (tag == nullptr || bytes < 12)
    {
        return false;
    }
    const uint32_t type = ReadU32(tag);
    if (type == Tag('c', 'u', 'r', 'v'))
    {
        const uint32_t count = ReadU32(tag + 8);
        if (count == 0)
        {
            return true;
        }
        if (count == 1 && bytes >= 14)
        {
            return ((static_cast<uint32_t>(tag[12]) << 8) | tag[13]) == 0x0100; // u8Fixed8 1.0
        }
        if (count == 2 && bytes >= 16)
        {
            return tag[12] == 0 && tag[13] == 0 && tag[14] == 0xff && tag[15] == 0xff;
        }
        return false;
    }
    if (type == Tag('p', 'a', 'r', 'a') && bytes >= 16)
    {
        const uint32_t function = (static_cast<uint32_t>(tag[8]) << 8) | tag[9];
        return function == 0 && ReadU32(tag + 12) == 0x00010000u; // Y = X^g, g = 1.0
    }
    return false;
}

void Multiply(const double a[9], const double b[9], double out[9])
{
    for (int r = 0; r < 3; ++r)
    {
        for (int c = 0; c < 3; ++c)
        {
            out[3 * r + c] = a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c];
        }
    }
}

bool Invert(const double m[9], double out[9])
{
    const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (std::fabs(det) < 1e-12)
    {
        return false;
    }
    const double inv = 1.0 / det;
    out[0] = (m[4] * m[8] - m[5] * m[7]) * inv;
    out[1] = (m[2] * m[7] - m[1] * m[8]) * inv;
    out[2] = (m[1] * m[5] - m[2] * m[4]) * inv;
    out[3] = (m[5] * m[6] - m[3] * m[8]) * inv;
    out[4] = (m[0] * m[8] - m[2] * m[6]) * inv;
    out[5] = (m[2] * m[3] - m[0] * m[5]) * inv;
    out[6] = (m[3] * m[7] - m[4] * m[6]) * inv;
    out[7] = (m[1] * m[6] - m[0] * m[7]) * inv;
    out[8] = (m[0] * m[4] - m[1] * m[3]) * inv;
    return true;
}

// RGB -> XYZ for primaries (x, y) and a white point (x, y), Y of white = 1 (the construction every colour text gives).
bool PrimariesToXyz(const double primaries[6], const double white[2], double out[9])
{
    double xyz[9];
    for (int c = 0; c < 3; ++c)
    {
        const double x = primaries[2 * c], y = primaries[2 * c + 1];
        xyz[c] = x / y;
        xyz[3 + c] = 1.0;
        xyz[6 + c] = (1.0 - x - y) / y;
    }
    double inverse[9];
    if (!Invert(xyz, inverse))
    {
        return false;
    }
    const double w[3] = { white[0] / white[1], 1.0, (1.0 - white[0] - white[1]) / white[1] };
    double scale[3];
    for (int c = 0; c < 3; ++c)
    {
        scale[c] = inverse[3 * c] * w[0] + inverse[3 * c + 1] * w[1] + inverse[3 * c + 2] * w[2];
    }
    for (int r = 0; r < 3; ++r)
    {
        for (int c = 0; c < 3; ++c)
        {
            out[3 * r + c] = xyz[3 * r + c] * scale[c];
        }
    }
    return true;
}

// Bradford chromatic adaptation from `source` white to `target` white (XYZ, Y = 1).
void Bradford(const double source[3], const double target[3], double out[9])
{
    static const double cone[9] = { 0.8951, 0.2664, -0.1614, -0.7502, 1.7135, 0.0367, 0.0389, -0.0685, 1.0296 };
    double coneInverse[9];
    Invert(cone, coneInverse);
    double s[3], t[3];
    for (int r = 0; r < 3; ++r)
    {
        s[r] = cone[3 * r] * source[0] + cone[3 * r + 1] * source[1] + cone[3 * r + 2] * source[2];
        t[r] = cone[3 * r] * target[0] + cone[3 * r + 1] * target[1] + cone[3 * r + 2] * target[2];
    }
    const double diagonal[9] = { t[0] / s[0], 0, 0, 0, t[1] / s[1], 0, 0, 0, t[2] / s[2] };
    double scaled[9];
    Multiply(diagonal, cone, scaled);
    Multiply(coneInverse, scaled, out);
}

} // namespace

extern "C" {

AVIFGPU_EXPORT int avifgpu_icc_to_rec2020_linear_matrix(const void* icc_profile, size_t size, float* out_matrix9, int32_t* out_is_rec2020)
{
    if (icc_profile == nullptr || out_matrix9 == nullptr)
    {
        return AVIFGPU_ERR_BAD_PARAM;
    }
    const Profile profile{ static_cast<const uint8_t*>(icc_profile), size };
    if (size < 132 || ReadU32(profile.data + 36) != Tag('a', 'c', 's', 'p'))
    {
        return AVIFGPU_ERR_BAD_PARAM; // not an ICC profile
    }
    if (ReadU32(profile.data + 16) != Tag('R', 'G', 'B', ' ') || ReadU32(profile.data + 20) != Tag('X', 'Y', 'Z', ' '))
    {
        return AVIFGPU_ERR_UNSUPPORTED; // not an RGB profile over the XYZ connection space
    }
    uint32_t ignored = 0;
    if (profile.Find(Tag('A', '2', 'B', '0'), &ignored) != nullptr)
    {
        return AVIFGPU_ERR_UNSUPPORTED; // a LUT profile: lcms2 would use the table for the perceptual intent
    }
    double colorants[9];
    double column[3];
    const uint32_t xyzTags[3] = { Tag('r', 'X', 'Y', 'Z'), Tag('g', 'X', 'Y', 'Z'), Tag('b', 'X', 'Y', 'Z') };
    const uint32_t trcTags[3] = { Tag('r', 'T', 'R', 'C'), Tag('g', 'T', 'R', 'C'), Tag('b', 'T', 'R', 'C') };
    for (int c = 0; c < 3; ++c)
    {
        if (!ReadXyz(profile, xyzTags[c], column) || !CurveIsLinear(profile, trcTags[c]))
        {
            return AVIFGPU_ERR_UNSUPPORTED;
        }
        colorants[c] = column[0];
        colorants[3 + c] = column[1];
        colorants[6 + c] = column[2];
    }
    // Rec.2020 (CreateRec2020LinearRGBProfile, ColorProfileGeneration.cpp:141-177: these primaries, D65) adapted to the
    // profile connection space white the way lcms2 builds an RGB profile (Bradford, D50 = 0.9642 / 1 / 0.8249).
    static const double rec2020Primaries[6] = { 0.708, 0.292, 0.170, 0.797, 0.131, 0.046 };
    static const double d65xy[2] = { 0.3127, 0.3290 };
    static const double d50[3] = { 0.9642, 1.0, 0.8249 };
    double rec2020[9], adaptation[9], rec2020Pcs[9], rec2020PcsInverse[9], matrix[9];
    const double d65[3] = { d65xy[0] / d65xy[1], 1.0, (1.0 - d65xy[0] - d65xy[1]) / d65xy[1] };
    if (!PrimariesToXyz(rec2020Primaries, d65xy, rec2020))
    {
        return AVIFGPU_ERR_UNSUPPORTED;
    }
    Bradford(d65, d50, adaptation);
    Multiply(adaptation, rec2020, rec2020Pcs);
    if (!Invert(rec2020Pcs, rec2020PcsInverse))
    {
        return AVIFGPU_ERR_UNSUPPORTED;
    }
    Multiply(rec2020PcsInverse, colorants, matrix);
    double deviation = 0.0;
    for (int i = 0; i < 9; ++i)
    {
        out_matrix9[i] = static_cast<float>(matrix[i]);
        deviation = std::fmax(deviation, std::fabs(matrix[i] - ((i % 4 == 0) ? 1.0 : 0.0)));
    }
    if (out_is_rec2020 != nullptr)
    {
        // s15Fixed16 colorants resolve 1.5e-5; a profile whose matrix is the identity to that precision IS Rec.2020
        // (the reference then skips the conversion: IsRec2020ColorProfile, ColorProfileConversion.cpp:128-131)
        *out_is_rec2020 = deviation < 2e-4 ? 1 : 0;
    }
    return AVIFGPU_OK;
}

} // extern "C"
