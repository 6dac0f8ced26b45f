// Host-side run of the SVD code the kernels share (csrc/svd.cuh is __host__ __device__): prints F, U, S, V for a set
// of matrices; tests/test_abi.py checks them against numpy in float64. Compiled with nvcc, runs without a GPU.
#include <cstdio>
#include <cstdlib>

#include "../../wgsparkl_b200/csrc/svd.cuh"

static float frand() { return (float)rand() / (float)RAND_MAX * 2.0f - 1.0f; }

int main() {
    srand(12345);
    for (int n = 0; n < 300; ++n) {
        float F[9], U[9], S[3], V[9];
        const int kind = n % 3;
        const float amp = (kind == 0) ? 1e-4f : (kind == 1) ? 0.05f : 0.6f;
        for (int i = 0; i < 9; ++i) F[i] = ((i % 4 == 0) ? 1.0f : 0.0f) + amp * frand();
        if (n % 50 == 7) F[0] = -F[0], F[1] = -F[1], F[2] = -F[2]; // inverted: det < 0
        b2::svd3<4>(F, U, S, V);
        printf("3");
        for (int i = 0; i < 9; ++i) printf(" %.9g", F[i]);
        for (int i = 0; i < 9; ++i) printf(" %.9g", U[i]);
        for (int i = 0; i < 3; ++i) printf(" %.9g", S[i]);
        for (int i = 0; i < 9; ++i) printf(" %.9g", V[i]);
        printf("\n");
    }
    for (int n = 0; n < 100; ++n) {
        float F[4], U[4], S[2], V[4];
        const float amp = (n % 2) ? 1e-4f : 0.5f;
        for (int i = 0; i < 4; ++i) F[i] = ((i % 3 == 0) ? 1.0f : 0.0f) + amp * frand();
        b2::svd2(F, U, S, V);
        printf("2");
        for (int i = 0; i < 4; ++i) printf(" %.9g", F[i]);
        for (int i = 0; i < 4; ++i) printf(" %.9g", U[i]);
        for (int i = 0; i < 2; ++i) printf(" %.9g", S[i]);
        for (int i = 0; i < 4; ++i) printf(" %.9g", V[i]);
        printf("\n");
    }
    return 0;
}
