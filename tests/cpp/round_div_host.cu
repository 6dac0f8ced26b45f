// Host-side check of round_div (csrc/common.cuh): rint(p * (1/h)) with the IEEE division only near half-integers must
// equal rint(p / h) for EVERY float - checked here on all ties k + 0.5 and their +-4 ulp neighbours for |k| < 2^20 and
// several cell widths, on 2e8 random quotients, and on the special values. Prints the number of mismatches (0).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../wgsparkl_b200/csrc/common.cuh"

static long check(float p, float h, float inv_h) {
    const float a = b2::round_div(p, h, inv_h), b = rintf(p / h);
    return std::memcmp(&a, &b, 4) != 0 && !(a != a && b != b);
}

int main() {
    long bad = 0, n = 0;
    const float widths[] = {1.0f, 0.25f, 0.1f, 0.3f, 3.0f, 0.0123f, 7.7f};
    for (float h : widths) {
        const float inv_h = 1.0f / h;
        for (int k = -(1 << 20); k < (1 << 20); k += 1) {
            float p = ((float)k + 0.5f) * h; // lands on or next to a tie of p / h
            for (int u = -4; u <= 4; ++u) {
                float q = p;
                for (int s = 0; s < (u < 0 ? -u : u); ++s) q = nextafterf(q, u < 0 ? -INFINITY : INFINITY);
                bad += check(q, h, inv_h), ++n;
            }
        }
        uint64_t state = 88172645463325252ull;
        for (long i = 0; i < 30000000; ++i) {
            state ^= state << 13, state ^= state >> 7, state ^= state << 17;
            const float q = ((float)(state >> 40) / 16777216.0f - 0.5f) * ((i & 1) ? 4096.0f : 1.0e7f);
            bad += check(q * h, h, inv_h), ++n;
        }
        const float special[] = {0.0f, -0.0f, INFINITY, -INFINITY, NAN, 1e30f, -1e30f, 4194304.0f * h, 8388608.0f * h, 1e-38f};
        for (float p : special) bad += check(p, h, inv_h), ++n;
    }
    printf("%ld %ld\n", n, bad);
    return 0;
}
