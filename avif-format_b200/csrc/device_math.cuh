// device_math.cuh -- float powf / expf / logf whose results are bit-identical to glibc's (>= 2.28).
//
// Why this exists: the reference's transfer functions (ColorTransfer.cpp:69-220) are a handful of IEEE float
// operations around powf / expf / logf, so "identical to the reference's CPU loop" means "identical libm
// results".  CUDA's own powf (a few ULP) and even a correctly rounded powf would differ from the host on
// 0.06-0.17 % of inputs, and the PQ curve amplifies a 1-ULP inner difference to hundreds of ULP (SURVEY.md
// section 7.4).  glibc computes these three functions in binary64 with a small table and a short polynomial
// and rounds once to binary32; that is cheap on B200 (full-rate FP64 FMA pipe) and reproducible, so the device
// evaluates the very same operation sequence: same tables (libm_tables.inc, recovered from the installed
// libm and cross-checked), same association, fused multiply-adds exactly where glibc's x86-64 FMA build has
// them (the ifunc variant every FMA-capable host selects; the non-FMA variant gives the same float in all
// ~1.4e9 cases compared).  tests/test_device_math_host.py compiles this header for the HOST and compares with
// the system libm; tests/test_gpu_primitives.py does the same on the device.
//
// The functions are __host__ __device__ so the identical source is what both tests exercise.
//
// Compile with -fmad=false: plain a*b+c below must NOT be contracted; every fused operation is an explicit
// fma().
#ifndef AVIF_DEVICE_MATH_CUH
#define AVIF_DEVICE_MATH_CUH

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "libm_tables.inc"

#if defined(__CUDACC__)
#define AVIF_HD __host__ __device__ __forceinline__
#define AVIF_CONSTEXPR_HD __host__ __device__ constexpr
#else
#define AVIF_HD inline
#define AVIF_CONSTEXPR_HD constexpr
#endif

namespace avifmath
{

#if defined(__CUDACC__)
// Device copies live in constant memory and are staged into shared memory by the kernels that index them with
// a per-lane (divergent) index; see StageLibmTables below.
__device__ __constant__ const uint64_t kExp2fTableConst[32] = { AVIF_LIBM_EXP2F_TABLE };
__device__ __constant__ const double kPowfLog2TableConst[32] = { AVIF_LIBM_POWF_LOG2_TABLE };
__device__ __constant__ const double kLogfTableConst[32] = { AVIF_LIBM_LOGF_TABLE };
#endif

#if defined(__CUDACC__)
// The polynomial coefficients, as NON-const __constant__ data: an FP64 instruction takes a constant-bank operand
// directly, whereas a binary64 literal costs two register moves at every use (measured: ~16 % of the instructions
// of the float decode kernel).  Not const on purpose -- a const initialiser would be folded back into literals.
static __device__ __constant__ double kLibmExp2fPolyConst[3] = AVIF_LIBM_EXP2F_POLY;
static __device__ __constant__ double kLibmExp2fPolyScaledConst[3] = AVIF_LIBM_EXP2F_POLY_SCALED;
static __device__ __constant__ double kLibmExp2fInvLn2ScaledConst[1] = { AVIF_LIBM_EXP2F_INVLN2_SCALED };
static __device__ __constant__ double kLibmPowfLog2PolyConst[5] = AVIF_LIBM_POWF_LOG2_POLY;
static __device__ __constant__ double kLibmLogfPolyConst[3] = AVIF_LIBM_LOGF_POLY;
static __device__ __constant__ double kLibmLogfLn2Const[1] = { AVIF_LIBM_LOGF_LN2 };
#endif
#if defined(__CUDA_ARCH__)
#define AVIF_LIBM_COEFFICIENTS(name, count, deviceArray, literal) const double* const name = deviceArray
#define AVIF_LIBM_SCALAR(deviceArray, literal) (deviceArray[0])
#else
#define AVIF_LIBM_COEFFICIENTS(name, count, deviceArray, literal) const double name[count] = literal
#define AVIF_LIBM_SCALAR(deviceArray, literal) (literal)
#endif

static const uint64_t kExp2fTableHost[32] = { AVIF_LIBM_EXP2F_TABLE };
static const double kPowfLog2TableHost[32] = { AVIF_LIBM_POWF_LOG2_TABLE };
static const double kLogfTableHost[32] = { AVIF_LIBM_LOGF_TABLE };

// Where the three 256-byte tables are read from.  Kernels pass pointers to shared-memory copies (a divergent
// index into __constant__ memory serialises; shared memory does not); host code passes the static arrays.
struct LibmTables
{
    const uint64_t* exp2f;   // 32 entries
    const double* powfLog2;  // 16 x { invc, log2 c }
    const double* logf;      // 16 x { invc, ln c }
};

inline LibmTables HostLibmTables()
{
    LibmTables t;
    t.exp2f = kExp2fTableHost;
    t.powfLog2 = kPowfLog2TableHost;
    t.logf = kLogfTableHost;
    return t;
}

#if defined(__CUDACC__)
// Cooperative copy of the tables into shared memory: `storage` must hold 96 eight-byte words.
__device__ __forceinline__ LibmTables StageLibmTables(uint64_t* storage, int threadIndex, int threadCount)
{
    for (int i = threadIndex; i < 96; i += threadCount)
    {
        uint64_t word;
        if (i < 32)
        {
            word = kExp2fTableConst[i];
        }
        else if (i < 64)
        {
            word = static_cast<uint64_t>(__double_as_longlong(kPowfLog2TableConst[i - 32]));
        }
        else
        {
            word = static_cast<uint64_t>(__double_as_longlong(kLogfTableConst[i - 64]));
        }
        storage[i] = word;
    }
    LibmTables t;
    t.exp2f = storage;
    t.powfLog2 = reinterpret_cast<const double*>(storage + 32);
    t.logf = reinterpret_cast<const double*>(storage + 64);
    return t;
}
#endif

// Table reads go through these accessors so that a kernel can hand the functions below a table set of its own kind.
AVIF_HD uint64_t TableExp2f(const LibmTables& t, uint32_t index) { return t.exp2f[index]; }
AVIF_HD void TablePowfLog2(const LibmTables& t, uint32_t index, double& invc, double& logc)
{
    invc = t.powfLog2[2 * index];
    logc = t.powfLog2[2 * index + 1];
}

// powf's log2 table with the exponent folded in.  glibc's log2_inline splits x = 2^k z, z in [OFF, 2 OFF), reads
// { invc, logc } for z's top four fraction bits and forms y0 = logc + (double)k before the polynomial.  Both k and the
// table index come from the same 13 bits of ix - OFF, so { invc, logc + (double)k } can be tabulated over (k, index): the
// very same binary64 addition, done once per entry instead of once per call (two FP64 instructions and three integer ones
// fewer per powf, a sixth of its cost).  An entry is 16 bytes; a table for k in [lowestExponent, 1] has
// (2 - lowestExponent) * 16 entries (PQ / SMPTE 428 bases are +0 or at least 2^-77: 25 KB for -96; every float down to the
// smallest subnormal after glibc's normalisation: 39 KB for -152).
struct PowfLog2Wide
{
    static constexpr uint32_t kOff = 0x3f330000u;
    // ix - BiasedOffset has the table index (in entries) in bits 19..31, counted from the entry of lowestExponent
    static AVIF_CONSTEXPR_HD uint32_t BiasedOffset(int lowestExponent) { return kOff - (static_cast<uint32_t>(-lowestExponent) << 23); }
    static AVIF_CONSTEXPR_HD uint32_t Entries(int lowestExponent) { return static_cast<uint32_t>(2 - lowestExponent) * 16u; }
};

// Entry `entry` of the wide table for `lowestExponent`: { invc, logc + (double)k } -- e_powf.c's `y0 = logc + (double) k`.
AVIF_HD void PowfLog2WideEntry(const double* narrowTable, int lowestExponent, uint32_t entry, double& invc, double& y0)
{
    const uint32_t i = entry % 16u;
    const int32_t k = static_cast<int32_t>(entry / 16u) + lowestExponent;
    invc = narrowTable[2 * i];
    y0 = narrowTable[2 * i + 1] + static_cast<double>(k);
}

// Host-side form of a staged wide table (the tests run the identical function bodies on the CPU).
struct LibmTablesWideHost
{
    LibmTables narrow;
    const double* wide; // Entries(lowestExponent) x { invc, y0 }
    uint32_t wideBiasedOffset;
    uint32_t wideLastEntry; // byte offset
};

AVIF_HD uint64_t TableExp2f(const LibmTablesWideHost& t, uint32_t index) { return t.narrow.exp2f[index]; }
AVIF_HD void TablePowfLog2(const LibmTablesWideHost& t, uint32_t index, double& invc, double& logc) { TablePowfLog2(t.narrow, index, invc, logc); }
AVIF_HD uint32_t WideBiasedOffset(const LibmTablesWideHost& t) { return t.wideBiasedOffset; }
AVIF_HD uint32_t WideLastEntry(const LibmTablesWideHost& t) { return t.wideLastEntry; }
AVIF_HD void TablePowfLog2Wide(const LibmTablesWideHost& t, uint32_t byteOffset, double& invc, double& y0)
{
    invc = t.wide[byteOffset / 8u];
    y0 = t.wide[byteOffset / 8u + 1u];
}

#if defined(__CUDACC__)
// The same tables named by their shared-STATE-SPACE addresses.  A generic pointer into shared memory costs an extra
// add per look-up on sm_100 (the shared window's base, rebuilt from the CTA's rank in its cluster, is added to every
// index before the LDS); an ld.shared on a 32-bit address lets the base sit in the instruction's uniform-register
// operand.  One add per look-up is 4 % of the float decode kernel's instructions.
struct LibmTablesShared
{
    uint32_t exp2f;
    uint32_t powfLog2;
    uint32_t logf;
    // the exponent-folded log2 table (PowfLog2Wide below), when a kernel has staged one
    uint32_t powfLog2Wide;     // address of its first entry
    uint32_t wideBiasedOffset; // PowfLog2Wide::BiasedOffset(lowestExponent)
    uint32_t wideLastEntry;    // byte offset of its last entry
};

__device__ __forceinline__ LibmTablesShared SharedSpace(const LibmTables& t)
{
    LibmTablesShared s;
    s.exp2f = static_cast<uint32_t>(__cvta_generic_to_shared(t.exp2f));
    s.powfLog2 = static_cast<uint32_t>(__cvta_generic_to_shared(t.powfLog2));
    s.logf = static_cast<uint32_t>(__cvta_generic_to_shared(t.logf));
    return s;
}

__device__ __forceinline__ uint64_t TableExp2f(const LibmTablesShared& t, uint32_t index)
{
    unsigned long long bits;
    asm("ld.shared.b64 %0, [%1];" : "=l"(bits) : "r"(t.exp2f + index * 8u)); // the tables never change once staged
    return bits;
}

__device__ __forceinline__ uint32_t WideBiasedOffset(const LibmTablesShared& t) { return t.wideBiasedOffset; }
__device__ __forceinline__ uint32_t WideLastEntry(const LibmTablesShared& t) { return t.wideLastEntry; }
__device__ __forceinline__ void TablePowfLog2Wide(const LibmTablesShared& t, uint32_t byteOffset, double& invc, double& y0)
{
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc), "=d"(y0) : "r"(t.powfLog2Wide + byteOffset));
}

// Cooperative fill of a wide table in shared memory (`storage`: Entries(lowestExponent) * 2 doubles) from the staged narrow one.
__device__ __forceinline__ void StagePowfLog2Wide(double* storage, const LibmTables& narrow, int lowestExponent, int threadIndex, int threadCount,
                                                  LibmTablesShared* shared)
{
    const uint32_t entries = PowfLog2Wide::Entries(lowestExponent);
    for (uint32_t entry = threadIndex; entry < entries; entry += threadCount)
    {
        double invc, y0;
        PowfLog2WideEntry(narrow.powfLog2, lowestExponent, entry, invc, y0);
        storage[2 * entry] = invc;
        storage[2 * entry + 1] = y0;
    }
    shared->powfLog2Wide = static_cast<uint32_t>(__cvta_generic_to_shared(storage));
    shared->wideBiasedOffset = PowfLog2Wide::BiasedOffset(lowestExponent);
    shared->wideLastEntry = (entries - 1u) * 16u;
}

__device__ __forceinline__ void TablePowfLog2(const LibmTablesShared& t, uint32_t index, double& invc, double& logc)
{
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc), "=d"(logc) : "r"(t.powfLog2 + index * 16u));
}
#endif

AVIF_HD uint32_t AsUint(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, sizeof(u));
    return u;
#endif
}

AVIF_HD float AsFloat(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, sizeof(f));
    return f;
#endif
}

AVIF_HD uint64_t AsUint64(double f)
{
#if defined(__CUDA_ARCH__)
    return static_cast<uint64_t>(__double_as_longlong(f));
#else
    uint64_t u;
    memcpy(&u, &f, sizeof(u));
    return u;
#endif
}

AVIF_HD double AsDouble(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(static_cast<long long>(u));
#else
    double f;
    memcpy(&f, &u, sizeof(f));
    return f;
#endif
}

// binary32 -> binary64 for a positive NORMAL float given as bits, with integer ops only (the conversion
// instruction runs on the quarter-rate pipe; this is exact for normal inputs: rebias the exponent by
// 1023 - 127 = 896 and widen the fraction by 29 zero bits).
AVIF_HD double NormalFloatBitsToDouble(uint32_t bits)
{
#if defined(__CUDA_ARCH__)
    const uint32_t hi = (bits >> 3) + 0x38000000u;
    const uint32_t lo = bits << 29;
    return __hiloint2double(static_cast<int>(hi), static_cast<int>(lo));
#else
    return static_cast<double>(AsFloat(bits));
#endif
}

// (double)k for an int32 k without the integer->double conversion instruction (quarter-rate pipe).
AVIF_HD double SmallIntToDouble(int32_t k)
{
#if defined(__CUDA_ARCH__)
    // bits(2^52) | (k + 2^31) has the value 2^52 + 2^31 + k exactly.
    return __hiloint2double(0x43300000, k ^ static_cast<int32_t>(0x80000000u)) - 0x1.000008p+52;
#else
    return static_cast<double>(k);
#endif
}

// AsDouble(tableBits + (ki << 47)): the table entry with k / 32 added to its exponent (glibc: `t += ki << (52 - 5)`).
// ki << 47 has no bits below bit 47, so on the device the addition is a single 32-bit add on the high word.
AVIF_HD double ScaleTableEntry(uint64_t tableBits, uint64_t ki)
{
#if defined(__CUDA_ARCH__)
    const uint32_t high = static_cast<uint32_t>(tableBits >> 32) + (static_cast<uint32_t>(ki) << 15);
    return __hiloint2double(static_cast<int>(high), static_cast<int>(static_cast<uint32_t>(tableBits)));
#else
    return AsDouble(tableBits + (ki << (52 - 5)));
#endif
}

// ---- exp2 core shared by powf (glibc e_powf.c exp2_inline) ------------------------------------------------

// 2^xd rounded once to binary32, for |xd| < 126 (callers check).  sign_bias is 0 on every path this library
// takes (no negative bases), so it is omitted.
template <typename Tables>
AVIF_HD float Exp2Inline(double xd, const Tables& t)
{
    AVIF_LIBM_COEFFICIENTS(C, 3, kLibmExp2fPolyConst, AVIF_LIBM_EXP2F_POLY);
    // x = k/N + r with r in [-1/(2N), 1/(2N)], N = 32
    double kd = xd + AVIF_LIBM_EXP2F_SHIFT_SCALED;
    const uint64_t ki = AsUint64(kd);
    kd -= AVIF_LIBM_EXP2F_SHIFT_SCALED;
    const double r = xd - kd;
    // exp2(x) = 2^(k/N) * 2^r ~= s * (C0*r^3 + C1*r^2 + C2*r + 1)
    uint64_t bits = TableExp2f(t, static_cast<uint32_t>(ki) % 32u);
    const double s = ScaleTableEntry(bits, ki);
    const double z = fma(C[0], r, C[1]);
    const double r2 = r * r;
    double y = fma(C[2], r, 1.0);
    y = fma(z, r2, y);
    y = y * s;
    return static_cast<float>(y);
}

// log2 of the positive normal float whose bits are ix (glibc e_powf.c log2_inline), in binary64.
template <typename Tables>
AVIF_HD double Log2Inline(uint32_t ix, const Tables& t)
{
    AVIF_LIBM_COEFFICIENTS(A, 5, kLibmPowfLog2PolyConst, AVIF_LIBM_POWF_LOG2_POLY);
    // x = 2^k z; z in [OFF, 2*OFF) with OFF = 0x3f330000; 16 sub-intervals
    const uint32_t tmp = ix - 0x3f330000u;
    const uint32_t i = (tmp >> (23 - 4)) % 16u;
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int32_t k = static_cast<int32_t>(top) >> 23; // arithmetic shift
    double invc, logc;
    TablePowfLog2(t, i, invc, logc);
    const double z = NormalFloatBitsToDouble(iz);

    // log2(x) = log1p(z/c-1)/ln2 + log2(c) + k
    const double r = fma(z, invc, -1.0);
    const double y0 = logc + SmallIntToDouble(k);

    const double r2 = r * r;
    double y = fma(A[0], r, A[1]);
    const double p = fma(A[2], r, A[3]);
    const double r4 = r2 * r2;
    double q = fma(A[4], r, y0);
    q = fma(p, r2, q);
    y = fma(y, r4, q);
    return y;
}

// Returns 2 if the float with bits iy is an even integer, 1 if odd, 0 if not an integer (glibc checkint).
AVIF_HD int CheckInt(uint32_t iy)
{
    const int e = static_cast<int>((iy >> 23) & 0xff);
    if (e < 0x7f)
    {
        return 0;
    }
    if (e > 0x7f + 23)
    {
        return 2;
    }
    if (iy & ((1u << (0x7f + 23 - e)) - 1))
    {
        return 0;
    }
    if (iy & (1u << (0x7f + 23 - e)))
    {
        return 1;
    }
    return 2;
}

AVIF_HD bool ZeroInfNan(uint32_t ix) { return 2 * ix - 1 >= 2u * 0x7f800000u - 1; }

AVIF_HD bool IsSignaling(uint32_t ix) { return 2 * (ix ^ 0x00400000u) > 2u * 0x7fc00000u; }

// powf(x, y) as glibc computes it (sysdeps/ieee754/flt-32/e_powf.c), errno / exception flags aside.
// kBaseNotNegative = true is a promise by the caller that x's sign bit is clear (x is +0, positive, +inf or a
// positive-signed NaN); the negative-base handling and the sign of the result then drop out at compile time.
template <bool kBaseNotNegative>
AVIF_HD float PowfImpl(float x, float y, const LibmTables& t)
{
    uint32_t signBias = 0;
    uint32_t ix = AsUint(x);
    const uint32_t iy = AsUint(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || ZeroInfNan(iy))
    {
        // Either (x < 0x1p-126 or inf or nan) or (y is 0 or inf or nan).
        if (ZeroInfNan(iy))
        {
            if (2 * iy == 0)
            {
                return IsSignaling(ix) ? x + y : 1.0f;
            }
            if (ix == 0x3f800000u)
            {
                return IsSignaling(iy) ? x + y : 1.0f;
            }
            if (2 * ix > 2u * 0x7f800000u || 2 * iy > 2u * 0x7f800000u)
            {
                return x + y;
            }
            if (2 * ix == 2 * 0x3f800000u)
            {
                return 1.0f;
            }
            if ((2 * ix < 2 * 0x3f800000u) == !(iy & 0x80000000u))
            {
                return 0.0f; // |x|<1 && y==inf or |x|>1 && y==-inf
            }
            return y * y;
        }
        if (ZeroInfNan(ix))
        {
            float x2 = x * x;
            if (!kBaseNotNegative && (ix & 0x80000000u) && CheckInt(iy) == 1)
            {
                x2 = -x2;
            }
            return (iy & 0x80000000u) ? 1 / x2 : x2;
        }
        // x and y are non-zero finite.
        if (!kBaseNotNegative && (ix & 0x80000000u))
        {
            // Finite x < 0.
            const int yint = CheckInt(iy);
            if (yint == 0)
            {
                return AsFloat(0x7fc00000u); // invalid: NaN
            }
            if (yint == 1)
            {
                signBias = 1;
            }
            ix &= 0x7fffffffu;
        }
        if (ix < 0x00800000u)
        {
            // Normalize subnormal x so exponent becomes negative.
            ix = AsUint(AsFloat(ix) * 0x1p23f);
            ix &= 0x7fffffffu;
            ix -= 23u << 23;
        }
    }
    const double logx = Log2Inline(ix, t);
    const double ylogx = static_cast<double>(y) * logx; // cannot overflow, y is single precision
    if (((AsUint64(ylogx) >> 47) & 0xffff) >= (AsUint64(126.0) >> 47))
    {
        // |y*log(x)| >= 126.
        if (ylogx > 0x1.fffffffd1d571p+6)
        {
            const float inf = AsFloat(0x7f800000u);
            return signBias ? -inf : inf;
        }
        if (ylogx <= -150.0)
        {
            return signBias ? -0.0f : 0.0f;
        }
    }
    const float result = Exp2Inline(ylogx, t);
    return signBias ? -result : result;
}

// True when powf(x, y) may go through PowfModerateExponent for every x whose sign bit is clear.
inline bool PowfExponentIsModerate(float y) { return y == y && y != 0.0f && (y < 0.0f ? -y : y) < 0.8f; }

// powf(x, y) for an exponent the caller has checked once (PowfExponentIsModerate: finite, non-zero, |y| < 0.8) and a
// base whose sign bit is clear -- the HLG OOTF's powf(luma, gamma - 1), gamma = 1.2 at 1000 nit.  For a normal x,
// |log2 x| <= 128, so |y log2 x| < 102.4 and glibc's |y log2 x| >= 126 screen cannot fire; y needs no screening at all.
// What is left is the main path of PowfImpl, operation for operation; x = 0, subnormal, inf and NaN take PowfImpl itself.
AVIF_HD float PowfModerateExponent(float x, float y, const LibmTables& t)
{
    const uint32_t ix = AsUint(x);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u)
    {
        return PowfImpl<true>(x, y, t);
    }
    const double ylogx = static_cast<double>(y) * Log2Inline(ix, t);
    return Exp2Inline(ylogx, t);
}

// True when PowfStraightLine(x, y) is glibc's powf(x, y) for every x in [+0, limit]: y finite and not zero, and
// y log2 x can neither overflow (>= 126) nor leave the range the exp2 core handles by itself.  Two cases cover the
// path: a positive exponent with bases up to 1 (PQ's 1 / m2 and 1 / m1, SMPTE 428's 2.6: y log2 x <= 0, and down to
// -150 * 2.6 the double result simply rounds to the +0 / subnormal glibc's underflow screen returns), and a moderate
// exponent with any finite base (the OOTF's gamma - 1).
inline bool PowfStraightLineCovers(float y, bool basesUpToOneOnly)
{
    if (!(y == y) || y == 0.0f) return false;
    const float magnitude = y < 0.0f ? -y : y;
    if (magnitude < 0.8f) return true;                        // |y log2 x| < 0.8 * 150 for every finite x > 0
    return basesUpToOneOnly && y > 0.0f && magnitude <= 6.5f; // y log2 x in [-975, 0]
}

// powf(x, yd) for a base whose sign bit is clear and that is finite (+0, subnormal or normal), and an exponent the caller
// has checked with PowfStraightLineCovers -- without a single branch, so that several evaluations interleave (the
// special-case branches of PowfImpl keep the compiler from overlapping two calls: measured as the float decode kernel's
// main stall).  x = 0 runs the main path on whatever its bits give (finite garbage, discarded) and selects `zeroResult`
// (+0 for y > 0, +inf for y < 0: what glibc returns) at the end; a subnormal x is normalised the way glibc does it
// (x * 2^23, exponent - 23) with two selects instead of a branch when kMaybeSubnormal, and must not occur otherwise.
// For -150 < y log2 x <= -126 glibc also falls through to the exp2 core; below that its screen returns +0, which is what
// the core's double result (< 2^-150) rounds to.
template <bool kMaybeSubnormal, typename Tables>
AVIF_HD float PowfStraightLine(float x, double yd, float zeroResult, const Tables& t)
{
    uint32_t ix = AsUint(x);
    const bool zero = ix == 0u;
    if (kMaybeSubnormal)
    {
        const bool small = ix < 0x00800000u;
        const float scaled = x * (small ? 0x1p23f : 1.0f);
        ix = AsUint(scaled) - (small ? (23u << 23) : 0u);
    }
    const double ylogx = yd * Log2Inline(ix, t);
    const float result = Exp2Inline(ylogx, t);
    return zero ? zeroResult : result;
}

// Log2Inline through the exponent-folded table: the same values in every operation (y0 comes out of the table as the sum
// glibc forms; z's bits are assembled straight from the fraction: iz = (tmp & 0x007fffff) + OFF is ix - top, and OFF has its
// low three bits clear, so NormalFloatBitsToDouble(iz) is { (fraction >> 3) + ((OFF >> 3) + 0x38000000), fraction << 29 }).
// An ix outside the table (a base that breaks the caller's promise, or the +0 whose result is selected away) reads the
// nearest end: finite garbage, never an access outside the table.
template <typename Tables>
AVIF_HD double Log2InlineWide(uint32_t ix, const Tables& t)
{
    AVIF_LIBM_COEFFICIENTS(A, 5, kLibmPowfLog2PolyConst, AVIF_LIBM_POWF_LOG2_POLY);
    const uint32_t tmp = ix - WideBiasedOffset(t);
    uint32_t at = (tmp >> 15) & 0x1fff0u;
    const uint32_t last = WideLastEntry(t);
    at = at < last ? at : last;
    double invc, y0;
    TablePowfLog2Wide(t, at, invc, y0);
    const uint32_t fraction = tmp & 0x007fffffu;
#if defined(__CUDA_ARCH__)
    const double z = __hiloint2double(static_cast<int>((fraction >> 3) + ((PowfLog2Wide::kOff >> 3) + 0x38000000u)), static_cast<int>(tmp << 29));
#else
    const double z = static_cast<double>(AsFloat(fraction + PowfLog2Wide::kOff));
#endif
    const double r = fma(z, invc, -1.0);
    const double r2 = r * r;
    double y = fma(A[0], r, A[1]);
    const double p = fma(A[2], r, A[3]);
    const double r4 = r2 * r2;
    double q = fma(A[4], r, y0);
    q = fma(p, r2, q);
    y = fma(y, r4, q);
    return y;
}

// PowfStraightLine on a table set that carries the wide log2 table; the bases must lie inside it (2^lowestExponent * OFF
// up to 2.8) or be +0.
template <bool kMaybeSubnormal, typename Tables>
AVIF_HD float PowfStraightLineWide(float x, double yd, float zeroResult, const Tables& t)
{
    uint32_t ix = AsUint(x);
    const bool zero = ix == 0u;
    if (kMaybeSubnormal)
    {
        const bool small = ix < 0x00800000u;
        const float scaled = x * (small ? 0x1p23f : 1.0f);
        ix = AsUint(scaled) - (small ? (23u << 23) : 0u);
    }
    const double ylogx = yd * Log2InlineWide(ix, t);
    const float result = Exp2Inline(ylogx, t);
    return zero ? zeroResult : result;
}

AVIF_HD float Powf(float x, float y, const LibmTables& t) { return PowfImpl<false>(x, y, t); }

// Powf for a base whose sign bit is known to be clear.
AVIF_HD float PowfOfNonNegative(float x, float y, const LibmTables& t) { return PowfImpl<true>(x, y, t); }

// expf(x) as glibc computes it (sysdeps/ieee754/flt-32/e_expf.c).
AVIF_HD float Expf(float x, const LibmTables& t)
{
    AVIF_LIBM_COEFFICIENTS(C, 3, kLibmExp2fPolyScaledConst, AVIF_LIBM_EXP2F_POLY_SCALED);
    const uint32_t abstop = (AsUint(x) >> 20) & 0x7ff;
    if (abstop >= (0x42b00000u >> 20)) // |x| >= 88 or x is nan
    {
        if (AsUint(x) == 0xff800000u)
        {
            return 0.0f;
        }
        if (abstop >= (0x7f800000u >> 20))
        {
            return x + x;
        }
        if (x > 0x1.62e42ep6f)
        {
            return AsFloat(0x7f800000u);
        }
        if (x < -0x1.9fe368p6f)
        {
            return 0.0f;
        }
    }
    const double xd = static_cast<double>(x);
    // x*N/Ln2 = k + r with r in [-1/2, 1/2] and int k.
    const double z = AVIF_LIBM_SCALAR(kLibmExp2fInvLn2ScaledConst, AVIF_LIBM_EXP2F_INVLN2_SCALED) * xd;
    double kd = z + AVIF_LIBM_EXP2F_SHIFT;
    const uint64_t ki = AsUint64(kd);
    kd -= AVIF_LIBM_EXP2F_SHIFT;
    const double r = z - kd;
    // exp(x) = 2^(k/N) * 2^(r/N) ~= s * (C0*r^3 + C1*r^2 + C2*r + 1)
    uint64_t bits = t.exp2f[ki % 32];
    const double s = ScaleTableEntry(bits, ki);
    const double zz = fma(C[0], r, C[1]);
    const double r2 = r * r;
    double y = fma(C[2], r, 1.0);
    y = fma(zz, r2, y);
    y = y * s;
    return static_cast<float>(y);
}

// The body of Expf for arguments known to be finite with |x| < 88 (no overflow / underflow / NaN screening).
template <typename Tables>
AVIF_HD float ExpfNoScreen(float x, const Tables& t)
{
    AVIF_LIBM_COEFFICIENTS(C, 3, kLibmExp2fPolyScaledConst, AVIF_LIBM_EXP2F_POLY_SCALED);
    const double xd = static_cast<double>(x);
    const double z = AVIF_LIBM_SCALAR(kLibmExp2fInvLn2ScaledConst, AVIF_LIBM_EXP2F_INVLN2_SCALED) * xd;
    double kd = z + AVIF_LIBM_EXP2F_SHIFT;
    const uint64_t ki = AsUint64(kd);
    kd -= AVIF_LIBM_EXP2F_SHIFT;
    const double r = z - kd;
    uint64_t bits = TableExp2f(t, static_cast<uint32_t>(ki) % 32u);
    const double s = ScaleTableEntry(bits, ki);
    const double zz = fma(C[0], r, C[1]);
    const double r2 = r * r;
    double y = fma(C[2], r, 1.0);
    y = fma(zz, r2, y);
    y = y * s;
    return static_cast<float>(y);
}

// logf(x) as glibc computes it (sysdeps/ieee754/flt-32/e_logf.c).
AVIF_HD float Logf(float x, const LibmTables& t)
{
    AVIF_LIBM_COEFFICIENTS(A, 3, kLibmLogfPolyConst, AVIF_LIBM_LOGF_POLY);
    uint32_t ix = AsUint(x);
    if (ix == 0x3f800000u)
    {
        return 0.0f;
    }
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u)
    {
        // x < 0x1p-126 or inf or nan.
        if (ix * 2 == 0)
        {
            return AsFloat(0xff800000u); // -inf
        }
        if (ix == 0x7f800000u)
        {
            return x;
        }
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u)
        {
            return AsFloat(0x7fc00000u);
        }
        // x is subnormal, normalize it.
        ix = AsUint(x * 0x1p23f);
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = static_cast<int>((tmp >> (23 - 4)) % 16);
    const int32_t k = static_cast<int32_t>(tmp) >> 23; // arithmetic shift
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    const double invc = t.logf[2 * i];
    const double logc = t.logf[2 * i + 1];
    const double z = NormalFloatBitsToDouble(iz);

    // log(x) = log1p(z/c-1) + log(c) + k*Ln2
    const double r = fma(z, invc, -1.0);
    const double y0 = fma(SmallIntToDouble(k), AVIF_LIBM_SCALAR(kLibmLogfLn2Const, AVIF_LIBM_LOGF_LN2), logc);

    const double r2 = r * r;
    double y = fma(A[1], r, A[2]);
    y = fma(A[0], r2, y);
    y = fma(y, r2, y0 + r);
    return static_cast<float>(y);
}

} // namespace avifmath

#endif // AVIF_DEVICE_MATH_CUH
