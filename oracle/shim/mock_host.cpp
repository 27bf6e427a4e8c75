/*
 * oracle/shim/mock_host.cpp -- TEST INFRASTRUCTURE ONLY.
 * The three host-side helpers the reference's row shuttle links against but whose own translation units
 * (Utilities.cpp, ColorProfileConversion.cpp) need the real Photoshop SDK / Little-CMS:
 *   GetImageSize / SetRect   semantics of Utilities.cpp:382-416 for a 32-bit-coordinate host,
 *   ColorProfileConversion    the "no ICC transform" state (ColorProfileConversion.cpp:107-131,143-156 return
 *                             early when there is no document profile / keepColorProfile is set), which is the
 *                             only state the accelerated path supports.
 */
#include "Utilities.h"
#include "ColorProfileConversion.h"

VPoint GetImageSize(const FormatRecordPtr formatRecord)
{
    VPoint size;
    if (formatRecord->HostSupports32BitCoordinates && formatRecord->PluginUsing32BitCoordinates)
    {
        size.h = formatRecord->imageSize32.h;
        size.v = formatRecord->imageSize32.v;
    }
    else
    {
        size.h = formatRecord->imageSize.h;
        size.v = formatRecord->imageSize.v;
    }
    return size;
}

void SetRect(FormatRecordPtr formatRecord, int32 top, int32 left, int32 bottom, int32 right)
{
    if (formatRecord->HostSupports32BitCoordinates && formatRecord->PluginUsing32BitCoordinates)
    {
        formatRecord->theRect32 = VRect{ top, left, bottom, right };
    }
    else
    {
        formatRecord->theRect = Rect{ static_cast<int16>(top), static_cast<int16>(left),
                                      static_cast<int16>(bottom), static_cast<int16>(right) };
    }
}

ColorProfileConversion::ColorProfileConversion(const FormatRecordPtr, bool hasAlpha, ColorTransferFunction, bool)
    : context(nullptr), documentProfile(), outputImageProfile(), transform(),
      numberOfChannels(hasAlpha ? 4 : 3), isSixteenBitMode(false)
{
}

ColorProfileConversion::ColorProfileConversion(const FormatRecordPtr, bool hasAlpha, int hostBitsPerChannel, bool)
    : context(nullptr), documentProfile(), outputImageProfile(), transform(),
      numberOfChannels(hasAlpha ? 4 : 3), isSixteenBitMode(hostBitsPerChannel == 16)
{
}

void ColorProfileConversion::ConvertRow(void*, cmsUInt32Number, cmsUInt32Number)
{
    // No transform: the row is left untouched.
}
