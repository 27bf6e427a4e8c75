// group_walk.cuh -- a grid-stride loop over (row, group-of-pixels) pairs without a division per step.
//
// The streaming kernels give every thread one group of adjacent pixels per step and stride over the image by the size
// of the grid.  Turning the flat index into (row, column) costs a 64-bit division -- as many instructions as the whole
// conversion of eight 8-bit samples.  GroupWalk divides twice per thread (first index, stride) and then steps both
// coordinates: row += stride / groupsPerRow, column += stride % groupsPerRow, one conditional carry.
#ifndef AVIFGPU_GROUP_WALK_CUH
#define AVIFGPU_GROUP_WALK_CUH

#include <stdint.h>

namespace avifgpu
{

struct GroupWalk
{
    int32_t row;    // < rowCount while the walk is inside the image
    int32_t column; // group index inside the row, 0 .. groupsPerRow - 1
    int32_t stepRows;
    int32_t stepColumns;
    int32_t groupsPerRow;

    // first = the thread's first flat index, stride = threads in the grid (both < 2^63, rows < 2^31)
    __device__ __forceinline__ GroupWalk(long long first, long long stride, int32_t groupsPerRow_, int32_t rowCount) : groupsPerRow(groupsPerRow_)
    {
        const long long lastRow = static_cast<long long>(rowCount);
        const long long firstRow = first / groupsPerRow;
        // a thread that starts past the image parks on rowCount (the loop condition) instead of overflowing int32
        row = static_cast<int32_t>(firstRow < lastRow ? firstRow : lastRow);
        column = static_cast<int32_t>(first - firstRow * groupsPerRow);
        const long long strideRows = stride / groupsPerRow;
        stepRows = static_cast<int32_t>(strideRows < lastRow ? strideRows : lastRow);
        stepColumns = static_cast<int32_t>(stride - strideRows * groupsPerRow);
    }

    __device__ __forceinline__ bool Inside(int32_t rowCount) const { return row < rowCount; }

    __device__ __forceinline__ void Advance(int32_t rowCount)
    {
        column += stepColumns;
        // rows saturate at rowCount: row + stepRows + 1 <= 2 * rowCount + 1 cannot wrap for rowCount < 2^30
        row += stepRows;
        if (column >= groupsPerRow)
        {
            column -= groupsPerRow;
            ++row;
        }
        (void)rowCount;
    }
};

} // namespace avifgpu

#endif
