// tests/native/libm_replica_check.cpp -- host-side gate for csrc/device_math.cuh.
// Compiles the SAME header the CUDA kernels use for the host and compares it with the system libm.
// Usage: libm_replica_check <stride> [threads]    (stride 1 = every positive float)
// Prints one line per function: name, comparisons, mismatches, first mismatching input bits.  Exit code 1 on
// any mismatch.
#include "device_math.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
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
    bool failed = false;

    // The exponents the path uses: ColorTransfer.cpp:73-77,100-104 (PQ), :126,138 (SMPTE 428), :201 (OOTF,
    // gamma 1.2), plus generic ones for the OOTF gamma range [1,3] -> exponent [0,2].
    const float m1 = 2610.0f / 16384.0f;
    const float m2 = 2523.0f / 4096.0f * 128.0f;
    const float exponents[] = { m1, m2, 1.0f / m2, 1.0f / m1, 1.0f / 2.6f, 2.6f, 1.2f - 1.0f,
                                0.0f, 1.0f, 2.0f, 0.5f, 1.7f, -0.3f, -2.0f, 3.0f };
    for (float y : exponents)
    {
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
