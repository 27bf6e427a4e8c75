/*
 * oracle/shim/PITypes.h -- TEST INFRASTRUCTURE ONLY.
 * Independently written stand-in for the handful of Photoshop SDK scalar / geometry types the
 * reference's pixel-path sources use.  Names and error numbers are the public SDK / classic Mac OS ABI.
 */
#ifndef ORACLE_SHIM_PITYPES_H
#define ORACLE_SHIM_PITYPES_H

#include <stdint.h>

typedef int8_t int8;
typedef int16_t int16;
typedef int32_t int32;
typedef int64_t int64;
typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef uint64_t unsigned64;

typedef int16 OSErr;
typedef char* Ptr;
typedef Ptr* Handle;
typedef unsigned char Boolean;

typedef struct Point { int16 v; int16 h; } Point;
typedef struct Rect { int16 top; int16 left; int16 bottom; int16 right; } Rect;
typedef struct VPoint { int32 v; int32 h; } VPoint;
typedef struct VRect { int32 top; int32 left; int32 bottom; int32 right; } VRect;

enum
{
    noErr = 0,
    readErr = -19,
    writErr = -20,
    memFullErr = -108,
    userCanceledErr = -128
};

#endif
