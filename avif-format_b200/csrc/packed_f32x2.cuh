// packed_f32x2.cuh -- Blackwell's two-lane FP32 instructions (FMUL2 / FADD2 / FFMA2; PTX mul/add/sub/fma.rn.f32x2,
// sm_100+): one issue slot for two IEEE binary32 operations, each rounded exactly as its scalar sibling.  The kernels
// that use them are bound by instruction issue, not by the FMA pipe, so halving the count of the plain float
// arithmetic is a direct win.
//
// One rule decides where they may be used.  ptxas CONTRACTS a packed multiply whose result feeds a packed add / sub
// into FFMA2 -- for the .f32x2 forms it does so even with explicit .rn modifiers and with -fmad=false (nvcc 12.9;
// the scalar forms honour both).  A contraction skips the rounding of the product, and this library's results are
// defined with that rounding (pixel_math.cuh), so:
//     * a Mul2 result must never be an operand of Add2 / Sub2 -- add the halves with scalar __fadd_rn instead
//       (ptxas does not split a packed multiply to fuse it with a scalar add; the parity tests would catch it);
//     * Add2 / Sub2 results may feed Mul2 (add -> multiply cannot contract);
//     * Fma2 is used only where the fused and the two-step result are the same number: a multiplication by a power
//       of two is exact, so fma(x, 0.25f, c) == (x * 0.25f) + c.
#ifndef AVIF_PACKED_F32X2_CUH
#define AVIF_PACKED_F32X2_CUH

#include <stdint.h>

namespace avifx2
{

typedef unsigned long long F32x2; // .lo = first operand of Pack

__device__ __forceinline__ F32x2 Pack(float lo, float hi)
{
    F32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}

__device__ __forceinline__ F32x2 Splat(float v) { return Pack(v, v); }

__device__ __forceinline__ void Unpack(F32x2 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

__device__ __forceinline__ F32x2 Add2(F32x2 a, F32x2 b)
{
    F32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// a + b rounded toward zero in both lanes (FADD2.RZ): the 2^23-bias truncation of the integer kernels
__device__ __forceinline__ F32x2 AddRz2(F32x2 a, F32x2 b)
{
    F32x2 r;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ F32x2 Sub2(F32x2 a, F32x2 b)
{
    F32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ F32x2 Mul2(F32x2 a, F32x2 b)
{
    F32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__device__ __forceinline__ F32x2 Fma2(F32x2 a, F32x2 b, F32x2 c)
{
    F32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

} // namespace avifx2

#endif
