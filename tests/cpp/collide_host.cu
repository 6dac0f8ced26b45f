// Host-side property test of the block-level collider culling (csrc/collide.cuh is __host__ __device__): for random
// balls, cuboids and capsules in random poses and random blocks around them, every node's collide() result must be
// IDENTICAL whether it looks at all bodies or only at those body_may_touch_block() keeps. Prints the number of
// blocks tested, how many body/block pairs were culled, and the number of mismatching nodes (must be 0).
// Compiled with nvcc, runs without a GPU (tests/test_abi.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../wgsparkl_b200/csrc/collide.cuh"

using namespace b2;

static float frand() { return (float)rand() / (float)RAND_MAX; }
static float srnd(float a) { return (2.0f * frand() - 1.0f) * a; }

template <int D>
static void random_body(BodyDev& b, float h) {
    std::memset(&b, 0, sizeof(b));
    const int kind = rand() % 3;
    b.shape_type = kind == 0 ? B200MPM_SHAPE_BALL : kind == 1 ? B200MPM_SHAPE_CUBOID : B200MPM_SHAPE_CAPSULE;
    b.radius = (0.3f + 3.0f * frand()) * h;
    for (int k = 0; k < D; ++k) {
        b.shape_a[k] = (0.2f + 6.0f * frand()) * h; // cuboid half extents / capsule end point a
        b.shape_b[k] = srnd(5.0f) * h; // capsule end point b
        b.trans[k] = srnd(9.0f) * h;
    }
    if (D == 2) {
        const float a = srnd(3.14159f);
        b.rot[0] = cosf(a), b.rot[1] = sinf(a), b.rot[2] = -sinf(a), b.rot[3] = cosf(a); // column-major 2x2
    } else {
        // random rotation from a random unit quaternion, column-major 3x3
        float q[4] = {srnd(1.f), srnd(1.f), srnd(1.f), srnd(1.f)};
        const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + 1e-9f;
        const float i = q[0] / n, j = q[1] / n, k = q[2] / n, w = q[3] / n;
        b.rot[0] = 1 - 2 * (j * j + k * k), b.rot[1] = 2 * (i * j + k * w), b.rot[2] = 2 * (i * k - j * w);
        b.rot[3] = 2 * (i * j - k * w), b.rot[4] = 1 - 2 * (i * i + k * k), b.rot[5] = 2 * (j * k + i * w);
        b.rot[6] = 2 * (i * k + j * w), b.rot[7] = 2 * (j * k - i * w), b.rot[8] = 1 - 2 * (i * i + j * j);
    }
}

template <int D>
static void run(int blocks, long& culled, long& pairs, long& mismatches) {
    constexpr int B = Dim<D>::BLOCK;
    for (int t = 0; t < blocks; ++t) {
        const float h = (t % 3 == 0) ? 1.0f : (t % 3 == 1) ? 0.25f : 3.0f;
        const int nb = 1 + rand() % 6;
        BodyDev bodies[B200MPM_MAX_BODIES];
        for (int i = 0; i < nb; ++i) random_body<D>(bodies[i], h);
        int vid[3] = {rand() % 5 - 2, rand() % 5 - 2, D == 3 ? rand() % 5 - 2 : 0};
        const float origin[3] = {(float)(vid[0] * B) * h, (float)(vid[1] * B) * h, (float)(vid[2] * B) * h};
        uint32_t mask = 0;
        for (int i = 0; i < nb; ++i)
            if (body_may_touch_block<D>(bodies[i], h, origin)) mask |= 1u << i;
        pairs += nb;
        culled += nb - __builtin_popcount(mask);
        const int nz = D == 3 ? B : 1;
        for (int z = 0; z < nz; ++z)
            for (int y = 0; y < B; ++y)
                for (int x = 0; x < B; ++x) {
                    const float pt[3] = {(float)(vid[0] * B + x) * h, (float)(vid[1] * B + y) * h, (float)(vid[2] * B + z) * h};
                    const NodeCdf all = collide<D>(bodies, (1u << nb) - 1u, h, pt);
                    const NodeCdf few = collide<D>(bodies, mask, h, pt);
                    if (all.affinities != few.affinities || all.closest_id != few.closest_id ||
                        std::memcmp(&all.distance, &few.distance, 4) != 0)
                        ++mismatches;
                }
    }
}

int main() {
    srand(20260117);
    long culled = 0, pairs = 0, mismatches = 0;
    run<3>(20000, culled, pairs, mismatches);
    printf("3 %d %ld %ld %ld\n", 20000, pairs, culled, mismatches);
    culled = pairs = mismatches = 0;
    run<2>(20000, culled, pairs, mismatches);
    printf("2 %d %ld %ld %ld\n", 20000, pairs, culled, mismatches);
    return 0;
}
