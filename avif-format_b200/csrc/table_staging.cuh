// table_staging.cuh -- the compact step table's shared-memory image, staged by the copy engine.
//
// CurveTableView::compact / firstBits are one allocation laid out like the kernels' shared-memory copy (curve_tables.h),
// so a single thread can hand it to cp.async.bulk in a few 16 KB pieces and every thread of the CTA then waits on one
// mbarrier.  Staging the same 67 KB with ordinary loads cost every persistent CTA ~7 us before its first tile (148 CTAs
// walk the same lines in step, 8 loads per thread in flight); the copy engine needs no registers, no issue slots, and the
// CTA's own first tile fetch proceeds underneath it.
#ifndef AVIFGPU_TABLE_STAGING_CUH
#define AVIFGPU_TABLE_STAGING_CUH

#include "curve_tables.h"

#include <stdint.h>

namespace avifgpu
{
namespace staging
{

constexpr uint32_t kTableCopyChunk = 16384; // bytes per bulk copy (a multiple of 16)

__device__ __forceinline__ uint32_t SharedAddress(const void* pointer) { return static_cast<uint32_t>(__cvta_generic_to_shared(pointer)); }

__device__ __forceinline__ void BarrierInit(uint32_t barrier, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barrier), "r"(arrivals) : "memory");
}

__device__ __forceinline__ void BarrierInitFence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void BarrierExpect(uint32_t barrier, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barrier), "r"(bytes) : "memory");
}

__device__ __forceinline__ void BulkCopyToShared(uint32_t target, const void* source, uint32_t bytes, uint32_t barrier)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(target), "l"(source), "r"(bytes),
                 "r"(barrier)
                 : "memory");
}

__device__ __forceinline__ void BarrierWait(uint32_t barrier, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred done;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n"
        "@done bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(barrier),
        "r"(parity)
        : "memory");
}

// ONE thread of the CTA: initialises `barrier` (8 bytes of shared memory, 8-byte aligned) and starts the copy of the table
// image to `sharedImage` (16-byte aligned, table.compactImageBytes long).  The CTA must pass a __syncthreads() before any
// other thread calls WaitTableImage (the initialisation has to be visible to the waiters).
__device__ __forceinline__ void BeginTableImageCopy(const CurveTableView& table, void* sharedImage, uint64_t* barrierStorage)
{
    const uint32_t barrier = SharedAddress(barrierStorage);
    BarrierInit(barrier, 1);
    BarrierInitFence();
    const uint32_t imageBytes = table.compactImageBytes;
    BarrierExpect(barrier, imageBytes);
    const uint32_t target = SharedAddress(sharedImage);
    const uint8_t* source = reinterpret_cast<const uint8_t*>(table.compact);
    for (uint32_t offset = 0; offset < imageBytes; offset += kTableCopyChunk)
    {
        BulkCopyToShared(target + offset, source + offset, min(kTableCopyChunk, imageBytes - offset), barrier);
    }
}

// Every thread that reads the table: returns once the whole image has landed.
__device__ __forceinline__ void WaitTableImage(uint64_t* barrierStorage) { BarrierWait(SharedAddress(barrierStorage), 0); }

} // namespace staging
} // namespace avifgpu

#endif
