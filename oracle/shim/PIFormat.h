/*
 * oracle/shim/PIFormat.h -- TEST INFRASTRUCTURE ONLY.
 * Mock of the Photoshop file-format parameter block, reduced to the fields the reference's row shuttle
 * reads or writes (grep "formatRecord->" over WriteHeifImage.cpp / ReadHeifImage.cpp / Utilities.cpp:375-446).
 * Field names are the SDK's; the layout is NOT the SDK's (this is a mock host, not Photoshop).
 */
#ifndef ORACLE_SHIM_PIFORMAT_H
#define ORACLE_SHIM_PIFORMAT_H

#include "PITypes.h"

typedef struct PSBufferID_* BufferID;

typedef OSErr (*AllocateBufferProc)(int32 size, BufferID* bufferID);
typedef Ptr (*LockBufferProc)(BufferID bufferID, Boolean moveHigh);
typedef void (*UnlockBufferProc)(BufferID bufferID);
typedef void (*FreeBufferProc)(BufferID bufferID);

typedef struct BufferProcs
{
    AllocateBufferProc allocateProc;
    LockBufferProc lockProc;
    UnlockBufferProc unlockProc;
    FreeBufferProc freeProc;
} BufferProcs;

typedef Boolean (*TestAbortProc)(void);
typedef void (*ProgressProc)(int32 done, int32 total);
typedef OSErr (*AdvanceStateProc)(void);

enum
{
    plugInModeGrayScale = 1,
    plugInModeRGBColor = 3,
    plugInModeGray16 = 10,
    plugInModeRGB48 = 11,
    plugInModeGray32 = 13,
    plugInModeRGB96 = 16
};

enum
{
    formatBadParameters = -30500,
    formatCannotRead = -30501,
    errPlugInHostInsufficient = -30900
};

typedef struct FormatRecord
{
    void* data;
    int32 rowBytes;
    int16 colBytes;
    int16 planeBytes;
    int16 loPlane;
    int16 hiPlane;
    int16 planes;
    int16 depth;
    int16 imageMode;
    Point imageSize;
    VPoint imageSize32;
    Rect theRect;
    VRect theRect32;
    Boolean HostSupports32BitCoordinates;
    Boolean PluginUsing32BitCoordinates;
    int32 maxData;
    int32 maxValue;
    int16 transparencyPlane;
    AdvanceStateProc advanceState;
    TestAbortProc abortProc;
    ProgressProc progressProc;
    BufferProcs* bufferProcs;
} FormatRecord;

typedef FormatRecord* FormatRecordPtr;

#endif
