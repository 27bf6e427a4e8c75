// tests/native/libm_replica_check.cpp -- host-side gate for csrc/device_math.cuh.
// Compiles the SAME header the CUDA kernels use for the host and compares it with the system libm.
// Usage: libm_replica_check <stride> [threads] [straight]   (stride 1 = every positive float; `straight` = only the
// branch-free powf section, for exhaustive runs)
// Prints one line per function: name, comparisons, mismatches, first mismatching input bits.  Exit code 1 on
// any mismatch.
#include "device_math.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

using namespace avifmath;

struct Tally
{
    unsigned long long compared = 0;
    unsigned long long mismatched = 0;
    uint32_t first = 0;
};

static bool SameFloat(float a, float b)
{
    const uint32_t ua = AsUint(a), ub = AsUint(b);
    if (ua == ub) return true;
    return (a != a) && (b != b); // any NaN equals any NaN
}

template <typename F, typename G>
static Tally Sweep(uint32_t begin, uint32_t end, uint32_t stride, int threads, F mine, G theirs)
{
    std::vector<Tally> parts(threads);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
    {
        pool.emplace_back([&, t]
        {
            Tally tally;
            const LibmTables tables = HostLibmTables();
            for (uint64_t u = static_cast<uint64_t>(begin) + static_cast<uint64_t>(t) * stride; u < end;
                 u += static_cast<uint64_t>(stride) * threads)
            {
                const float x = AsFloat(static_cast<uint32_t>(u));
                const float a = mine(x, tables);
                const float b = theirs(x);
                tally.compared++;
                if (!SameFloat(a, b))
                {
                    if (tally.mismatched == 0) tally.first = static_cast<uint32_t>(u);
                    tally.mismatched++;
                }
            }
            parts[t] = tally;
        });
    }
    for (auto& th : pool) th.join();
    Tally total;
    for (const Tally& p : parts)
    {
        total.compared += p.compared;
        if (p.mismatched && !total.mismatched) total.first = p.first;
        total.mismatched += p.mismatched;
    }
    return total;
}

int main(int argc, char** argv)
{
    const uint32_t stride = argc > 1 ? static_cast<uint32_t>(std::strtoul(argv[1], nullptr, 10)) : 101;
    const int threads = argc > 2 ? std::atoi(argv[2]) : 4;
    const bool straightOnly = argc > 3 && std::string(argv[3]) == "straight"; // only the PowfStraightLine section (exhaustive runs)
    bool failed = false;

    // The exponents the path uses: ColorTransfer.cpp:73-77,100-104 (PQ), :126,138 (SMPTE 428), :201 (OOTF,
    // gamma 1.2), plus generic ones for the OOTF gamma range [1,3] -> exponent [0,2].
    const float m1 = 2610.0f / 16384.0f;
    const float m2 = 2523.0f / 4096.0f * 128.0f;
    const float exponents[] = { m1, m2, 1.0f / m2, 1.0f / m1, 1.0f / 2.6f, 2.6f, 1.2f - 1.0f,
                                0.0f, 1.0f, 2.0f, 0.5f, 1.7f, -0.3f, -2.0f, 3.0f };
    for (float y : exponents)
    {
        if (straightOnly) break;
        // all of [0, +inf] and NaNs, positive half; negative bases behave per glibc as well
        const Tally pos = Sweep(0x00000000u, 0x7fc00001u, stride, threads,
                                [y](float x, const LibmTables& t) { return Powf(x, y, t); },
                                [y](float x) { return powf(x, y); });
        const Tally neg = Sweep(0x80000000u, 0xff800001u, stride * 16u + 1u, threads,
                                [y](float x, const LibmTables& t) { return Powf(x, y, t); },
                                [y](float x) { return powf(x, y); });
        // the variant for bases whose sign bit is known to be clear (HLG OOTF luma) over its whole domain
        const Tally unsignedBase = Sweep(0x00000000u, 0x7fc00001u, stride, threads,
                                         [y](float x, const LibmTables& t) { return PowfOfNonNegative(x, y, t); },
                                         [y](float x) { return powf(x, y); });
        std::printf("powf y=%-14a compared=%llu mismatched=%llu first=%08x | negative-x compared=%llu mismatched=%llu first=%08x | "
                    "non-negative variant mismatched=%llu\n",
                    y, pos.compared, pos.mismatched, pos.first, neg.compared, neg.mismatched, neg.first, unsignedBase.mismatched);
        failed |= pos.mismatched != 0 || neg.mismatched != 0 || unsignedBase.mismatched != 0;
    }
    // The branch-free powf of the tuned float decode kernel (PowfStraightLine), on the domains its callers promise:
    // bases in [+0, 1] for the positive exponents of PQ / SMPTE 428, every finite base for a moderate exponent.
    {
        const float straightExponents[] = { 1.0f / m2, 1.0f / m1, 2.6f, 1.2f - 1.0f, 0.5f, -0.3f, 0.79f, 6.5f };
        for (float y : straightExponents)
        {
            const bool upToOne = !PowfStraightLineCovers(y, false);
            if (!PowfStraightLineCovers(y, upToOne))
            {
                std::printf("straight-line powf y=%a not covered\n", y);
                failed = true;
                continue;
            }
            const uint32_t end = upToOne ? 0x3f800001u : 0x7f800000u;
            const float zeroResult = y < 0.0f ? AsFloat(0x7f800000u) : 0.0f;
            const double yd = static_cast<double>(y);
            Tally all = Sweep(0x00000000u, end, stride, threads,
                                    [yd, zeroResult](float x, const LibmTables& t) { return PowfStraightLine<true>(x, yd, zeroResult, t); },
                                    [y](float x) { return powf(x, y); });
            // without the subnormal normalisation: +0 and the normal bases
            Tally normal = Sweep(0x00800000u, end, stride, threads,
                                 [yd, zeroResult](float x, const LibmTables& t) { return PowfStraightLine<false>(x, yd, zeroResult, t); },
                                 [y](float x) { return powf(x, y); });
            {
                const LibmTables tables = HostLibmTables();
                normal.compared++;
                if (!SameFloat(PowfStraightLine<false>(0.0f, yd, zeroResult, tables), powf(0.0f, y))) normal.mismatched++;
            }
            // the subnormal bases, every one of them when the stride is small enough to afford it
            const Tally subnormal = Sweep(0x00000000u, 0x00800000u, stride > 16u ? 16u : stride, threads,
                                          [yd, zeroResult](float x, const LibmTables& t) { return PowfStraightLine<true>(x, yd, zeroResult, t); },
                                          [y](float x) { return powf(x, y); });
            // the same through the exponent-folded log2 table (PowfStraightLineWide), built here as the kernels stage it
            {
                const int lowestExponent = upToOne ? -96 : -152;
                std::vector<double> wide(PowfLog2Wide::Entries(lowestExponent) * 2u);
                for (uint32_t entry = 0; entry < PowfLog2Wide::Entries(lowestExponent); ++entry)
                {
                    PowfLog2WideEntry(kPowfLog2TableHost, lowestExponent, entry, wide[2 * entry], wide[2 * entry + 1]);
                }
                LibmTablesWideHost wideTables;
                wideTables.narrow = HostLibmTables();
                wideTables.wide = wide.data();
                wideTables.wideBiasedOffset = PowfLog2Wide::BiasedOffset(lowestExponent);
                wideTables.wideLastEntry = (PowfLog2Wide::Entries(lowestExponent) - 1u) * 16u;
                // bases inside the table: from 2^-90 (the kernels promise at least 2^-77) for -96, every positive float for -152
                const uint32_t begin = upToOne ? 0x12800000u : 0x00000000u;
                const Tally viaWide = upToOne ? Sweep(begin, end, stride, threads,
                                                      [yd, zeroResult, &wideTables](float x, const LibmTables&) { return PowfStraightLineWide<false>(x, yd, zeroResult, wideTables); },
                                                      [y](float x) { return powf(x, y); })
                                              : Sweep(begin, 0x40300000u, stride, threads, // up to 2.75: the table's top
                                                      [yd, zeroResult, &wideTables](float x, const LibmTables&) { return PowfStraightLineWide<true>(x, yd, zeroResult, wideTables); },
                                                      [y](float x) { return powf(x, y); });
                all.compared += viaWide.compared;
                if (viaWide.mismatched && !all.mismatched) all.first = viaWide.first;
                all.mismatched += viaWide.mismatched;
                if (!SameFloat(PowfStraightLineWide<false>(0.0f, yd, zeroResult, wideTables), powf(0.0f, y))) all.mismatched++;
            }
            std::printf("straight-line powf y=%-14a compared=%llu mismatched=%llu first=%08x | normal-only variant compared=%llu mismatched=%llu first=%08x | "
                        "subnormal bases compared=%llu mismatched=%llu\n",
                        y, all.compared, all.mismatched, all.first, normal.compared, normal.mismatched, normal.first, subnormal.compared,
                        subnormal.mismatched);
            failed |= all.mismatched != 0 || normal.mismatched != 0 || subnormal.mismatched != 0;
        }
    }
    // The two constant divisions of HLGToLinear as the tuned float decode kernel performs them (kernels_fast_decode.cu
    // DivideBySplit: fma(x, hi, x * lo) with hi + lo = 1 / d), against the IEEE division, for EVERY numerator: value - c for
    // value in (0.5, 1] and expf(..) + b in [1, 16).  Plain IEEE operations: the CPU's answer is the GPU's.
    {
        const float a = 0.17883277f, c = 0.55991073f;
        const float hiA = static_cast<float>(1.0 / static_cast<double>(a)), loA = static_cast<float>(1.0 / static_cast<double>(a) - static_cast<double>(hiA));
        const float hi12 = static_cast<float>(1.0 / 12.0), lo12 = static_cast<float>(1.0 / 12.0 - static_cast<double>(hi12));
        unsigned long long compared = 0, mismatched = 0;
        for (uint32_t bits = 0x3f000001u; bits <= 0x3f800000u; ++bits, ++compared)
        {
            const float numerator = AsFloat(bits) - c;
            if (AsUint(fmaf(numerator, hiA, numerator * loA)) != AsUint(numerator / a)) ++mismatched;
        }
        for (uint32_t bits = 0x3f800000u; bits < 0x41800000u; ++bits, ++compared)
        {
            const float x = AsFloat(bits);
            if (AsUint(fmaf(x, hi12, x * lo12)) != AsUint(x / 12.0f)) ++mismatched;
        }
        std::printf("split-reciprocal divisions (HLG) compared=%llu mismatched=%llu\n", compared, mismatched);
        failed |= mismatched != 0;
    }
    if (straightOnly) return failed ? 1 : 0;
    {
        const Tally a = Sweep(0x00000000u, 0x7fc00001u, stride, threads,
                              [](float x, const LibmTables& t) { return Expf(x, t); }, [](float x) { return expf(x); });
        const Tally b = Sweep(0x80000000u, 0xffc00001u, stride, threads,
                              [](float x, const LibmTables& t) { return Expf(x, t); }, [](float x) { return expf(x); });
        std::printf("expf compared=%llu mismatched=%llu first=%08x\n", a.compared + b.compared,
                    a.mismatched + b.mismatched, a.mismatched ? a.first : b.first);
        failed |= a.mismatched != 0 || b.mismatched != 0;
    }
    {
        const Tally a = Sweep(0x00000000u, 0x7fc00001u, stride, threads,
                              [](float x, const LibmTables& t) { return Logf(x, t); }, [](float x) { return logf(x); });
        const Tally b = Sweep(0x80000000u, 0xffc00001u, stride * 16u + 1u, threads,
                              [](float x, const LibmTables& t) { return Logf(x, t); }, [](float x) { return logf(x); });
        std::printf("logf compared=%llu mismatched=%llu first=%08x\n", a.compared + b.compared,
                    a.mismatched + b.mismatched, a.mismatched ? a.first : b.first);
        failed |= a.mismatched != 0 || b.mismatched != 0;
    }
    return failed ? 1 : 0;
}
