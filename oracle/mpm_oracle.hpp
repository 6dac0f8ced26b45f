// ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
//
// A CPU restatement of wgsparkl's MPM substep, following the reference's WGSL shaders
// kernel by kernel (citations are `file:line` relative to the wgsparkl source tree).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may build, load or call this code. The shipped library (wgsparkl_b200/csrc) never does.
//
// PARITY STATUS
//   * pinned by a reference test: the exclusive prefix sum only (src/grid/prefix_sum.rs:180-230,
//     inputs ones / iota / random%10000, LEN 15071; expected = eval_cpu, prefix_sum.rs:71-83).
//   * everything else is PARITY UNPINNED: the reference holds no golden vectors for any MPM
//     stage and cannot be built here (no Rust / Vulkan / WGSL toolchain). The arithmetic that
//     lives in un-vendored dependencies is restated from its published definition:
//       dimforge/wgmath @ 6d17942bd841efdfcc696d8455b22be3a8ddfe8d (Cargo.toml:26-32)
//         wgebra  svd2/svd3 (svd, recompose), inv::inv3/inv4, sim2/sim3::mulPt
//         wgparry Shape::projectPointOnBoundary (ball, cuboid, capsule)
//         wgrapier Body::{velocity_at_point, applyImpulse, integrateVelocity, updateMprops}
//     Every use of the SVD on this path is invariant to the decomposition's ordering / sign
//     convention, so the SVD here is computed by a double-precision Jacobi
//     iteration and rounded to f32 (validated against numpy in tests/test_oracle.py).
//
// Sequential semantics: wherever the reference's result depends on atomic ordering
// (block header ids, intra-block sorted order, per-node linked-list order) this restatement
// executes the invocations in increasing global invocation id. Everything else is plain f32
// arithmetic compiled with -ffp-contract=off so that no FMA is formed behind WGSL's back.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/b200mpm.h"

namespace oracle {

constexpr uint32_t NONE = 0xffffffffu; // grid.wgsl:80
constexpr uint32_t NUM_CELL_PER_BLOCK = 64; // grid.wgsl:43

// ---------------------------------------------------------------------------------------
// small fixed-size linear algebra, column-major like WGSL / nalgebra
// ---------------------------------------------------------------------------------------
template <int D>
struct Vec {
    float v[D];
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    static Vec zero() {
        Vec r;
        for (int i = 0; i < D; ++i) r.v[i] = 0.0f;
        return r;
    }
    static Vec splat(float x) {
        Vec r;
        for (int i = 0; i < D; ++i) r.v[i] = x;
        return r;
    }
};
template <int D>
inline Vec<D> operator+(Vec<D> a, Vec<D> b) {
    Vec<D> r;
    for (int i = 0; i < D; ++i) r.v[i] = a.v[i] + b.v[i];
    return r;
}
template <int D>
inline Vec<D> operator-(Vec<D> a, Vec<D> b) {
    Vec<D> r;
    for (int i = 0; i < D; ++i) r.v[i] = a.v[i] - b.v[i];
    return r;
}
template <int D>
inline Vec<D> operator-(Vec<D> a) {
    Vec<D> r;
    for (int i = 0; i < D; ++i) r.v[i] = -a.v[i];
    return r;
}
template <int D>
inline Vec<D> operator*(Vec<D> a, float s) {
    Vec<D> r;
    for (int i = 0; i < D; ++i) r.v[i] = a.v[i] * s;
    return r;
}
template <int D>
inline Vec<D> operator*(float s, Vec<D> a) {
    return a * s;
}
template <int D>
inline Vec<D> operator/(Vec<D> a, float s) {
    Vec<D> r;
    for (int i = 0; i < D; ++i) r.v[i] = a.v[i] / s;
    return r;
}
template <int D>
inline Vec<D> mul_comp(Vec<D> a, Vec<D> b) {
    Vec<D> r;
    for (int i = 0; i < D; ++i) r.v[i] = a.v[i] * b.v[i];
    return r;
}
template <int D>
inline float dot(Vec<D> a, Vec<D> b) {
    // WGSL dot(): a.x*b.x + a.y*b.y (+ a.z*b.z), left to right.
    float s = a.v[0] * b.v[0];
    for (int i = 1; i < D; ++i) s = s + a.v[i] * b.v[i];
    return s;
}
template <int D>
inline float length(Vec<D> a) {
    return std::sqrt(dot(a, a));
}
inline Vec<3> cross(Vec<3> a, Vec<3> b) {
    return Vec<3>{{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
}

// Column-major square matrix: c[col][row].
template <int D>
struct Mat {
    Vec<D> c[D];
    static Mat zero() {
        Mat m;
        for (int i = 0; i < D; ++i) m.c[i] = Vec<D>::zero();
        return m;
    }
    static Mat identity() {
        Mat m = zero();
        for (int i = 0; i < D; ++i) m.c[i].v[i] = 1.0f;
        return m;
    }
    float& at(int row, int col) { return c[col].v[row]; }
    float at(int row, int col) const { return c[col].v[row]; }
};
template <int D>
inline Mat<D> operator+(Mat<D> a, Mat<D> b) {
    Mat<D> r;
    for (int i = 0; i < D; ++i) r.c[i] = a.c[i] + b.c[i];
    return r;
}
template <int D>
inline Mat<D> operator-(Mat<D> a, Mat<D> b) {
    Mat<D> r;
    for (int i = 0; i < D; ++i) r.c[i] = a.c[i] - b.c[i];
    return r;
}
template <int D>
inline Mat<D> operator*(Mat<D> a, float s) {
    Mat<D> r;
    for (int i = 0; i < D; ++i) r.c[i] = a.c[i] * s;
    return r;
}
template <int D>
inline Vec<D> operator*(Mat<D> m, Vec<D> x) {
    // WGSL mat * vec = sum_k col_k * x_k.
    Vec<D> r = m.c[0] * x.v[0];
    for (int k = 1; k < D; ++k) r = r + m.c[k] * x.v[k];
    return r;
}
template <int D>
inline Mat<D> operator*(Mat<D> a, Mat<D> b) {
    Mat<D> r;
    for (int j = 0; j < D; ++j) r.c[j] = a * b.c[j];
    return r;
}
template <int D>
inline Mat<D> transpose(Mat<D> a) {
    Mat<D> r;
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) r.c[i].v[j] = a.c[j].v[i];
    return r;
}
// outer_product(a, b) = a b^T built by columns a*b.x, a*b.y, ... (g2p.wgsl:242-261).
template <int D>
inline Mat<D> outer_product(Vec<D> a, Vec<D> b) {
    Mat<D> r;
    for (int j = 0; j < D; ++j) r.c[j] = a * b.v[j];
    return r;
}
inline float determinant(const Mat<2>& m) { return m.at(0, 0) * m.at(1, 1) - m.at(0, 1) * m.at(1, 0); }
inline float determinant(const Mat<3>& m) {
    return m.at(0, 0) * (m.at(1, 1) * m.at(2, 2) - m.at(1, 2) * m.at(2, 1)) -
           m.at(0, 1) * (m.at(1, 0) * m.at(2, 2) - m.at(1, 2) * m.at(2, 0)) +
           m.at(0, 2) * (m.at(1, 0) * m.at(2, 1) - m.at(1, 1) * m.at(2, 0));
}

// ---------------------------------------------------------------------------------------
// wgebra::svd2 / svd3 (not vendored; contract in SURVEY Appendix B): F = U diag(S) Vt with
// U, Vt proper rotations, |S| sorted descending, the sign of det(F) carried by the last S.
// Computed in double by cyclic Jacobi on F^T F, then rounded to f32.
// ---------------------------------------------------------------------------------------
template <int D>
struct Svd {
    Mat<D> U;
    Vec<D> S;
    Mat<D> Vt;
    double Sd[D]; // the same singular values before rounding to f32 (used by the exact-sigma mode only)
};

// Exact-sigma mode (test diagnostic, OFF by default = reference semantics). The WGSL rounds each singular
// value to f32 and then forms sigma - 1, J - 1 and log(sigma) from the rounded value, which injects an absolute
// error of ~6e-8 into strains that stiff materials multiply by 1e7..1e9. With the mode ON those three
// quantities are formed from the unrounded (f64) singular values instead - the exact-arithmetic value of the
// same formulas on the same f32 state. Tests use it to tell the reference's own rounding noise apart from
// errors of the CUDA path (tests/test_gpu_step.py::test_sand_against_exact_sigma_oracle).
inline bool& exact_sigma_mode() {
    static bool on = false;
    return on;
}

template <int D>
inline Svd<D> svd(const Mat<D>& Ff) {
    double F[D][D]; // F[row][col]
    for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) F[r][c] = (double)Ff.at(r, c);
    double A[D][D]; // F^T F
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            double s = 0;
            for (int k = 0; k < D; ++k) s += F[k][i] * F[k][j];
            A[i][j] = s;
        }
    double V[D][D];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0;
        for (int p = 0; p < D; ++p)
            for (int q = p + 1; q < D; ++q) off += A[p][q] * A[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < D; ++p)
            for (int q = p + 1; q < D; ++q) {
                if (std::fabs(A[p][q]) < 1e-300) continue;
                double tau = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < D; ++k) { // A <- A J
                    double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < D; ++k) { // A <- J^T A
                    double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < D; ++k) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    // Sort eigenpairs descending.
    int order[D];
    for (int i = 0; i < D; ++i) order[i] = i;
    std::sort(order, order + D, [&](int a, int b) { return A[a][a] > A[b][b]; });
    double Vs[D][D];
    for (int j = 0; j < D; ++j)
        for (int i = 0; i < D; ++i) Vs[i][j] = V[i][order[j]];
    auto det = [](double M[D][D]) -> double {
        if constexpr (D == 2) {
            return M[0][0] * M[1][1] - M[0][1] * M[1][0];
        } else {
            return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) -
                   M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                   M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
        }
    };
    if (det(Vs) < 0)
        for (int i = 0; i < D; ++i) Vs[i][D - 1] = -Vs[i][D - 1];
    // B = F V ; sigma_j = |B_j| ; U_j = B_j / sigma_j with Gram-Schmidt repair for tiny sigma.
    double B[D][D], Ud[D][D] = {}, sig[D];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            double s = 0;
            for (int k = 0; k < D; ++k) s += F[i][k] * Vs[k][j];
            B[i][j] = s;
        }
    for (int j = 0; j < D; ++j) {
        double n = 0;
        for (int i = 0; i < D; ++i) n += B[i][j] * B[i][j];
        sig[j] = std::sqrt(n);
    }
    double scale = std::max(sig[0], 1e-300);
    for (int j = 0; j < D; ++j) {
        double u[D];
        for (int i = 0; i < D; ++i) u[i] = B[i][j];
        for (int pass = 0; pass < 2; ++pass)
            for (int k = 0; k < j; ++k) {
                double d = 0;
                for (int i = 0; i < D; ++i) d += u[i] * Ud[i][k];
                for (int i = 0; i < D; ++i) u[i] -= d * Ud[i][k];
            }
        double n = 0;
        for (int i = 0; i < D; ++i) n += u[i] * u[i];
        n = std::sqrt(n);
        if (n > 1e-12 * scale && sig[j] > 1e-150) {
            for (int i = 0; i < D; ++i) Ud[i][j] = u[i] / n;
        } else {
            // Rank-deficient: pick any unit vector orthogonal to the previous columns.
            if constexpr (D == 2) {
                if (j == 0) {
                    Ud[0][0] = 1;
                    Ud[1][0] = 0;
                } else {
                    Ud[0][1] = -Ud[1][0];
                    Ud[1][1] = Ud[0][0];
                }
            } else {
                if (j == 0) {
                    Ud[0][0] = 1;
                    Ud[1][0] = 0;
                    Ud[2][0] = 0;
                } else if (j == 1) {
                    int m = 0;
                    for (int i = 1; i < 3; ++i)
                        if (std::fabs(Ud[i][0]) < std::fabs(Ud[m][0])) m = i;
                    double e[3] = {0, 0, 0};
                    e[m] = 1;
                    double d = Ud[m][0];
                    double w[3], nn = 0;
                    for (int i = 0; i < 3; ++i) {
                        w[i] = e[i] - d * Ud[i][0];
                        nn += w[i] * w[i];
                    }
                    nn = std::sqrt(nn);
                    for (int i = 0; i < 3; ++i) Ud[i][1] = w[i] / nn;
                } else {
                    Ud[0][2] = Ud[1][0] * Ud[2][1] - Ud[2][0] * Ud[1][1];
                    Ud[1][2] = Ud[2][0] * Ud[0][1] - Ud[0][0] * Ud[2][1];
                    Ud[2][2] = Ud[0][0] * Ud[1][1] - Ud[1][0] * Ud[0][1];
                }
            }
        }
    }
    if (det(Ud) < 0) {
        for (int i = 0; i < D; ++i) Ud[i][D - 1] = -Ud[i][D - 1];
        sig[D - 1] = -sig[D - 1];
    }
    Svd<D> out;
    for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) {
            out.U.at(r, c) = (float)Ud[r][c];
            out.Vt.at(r, c) = (float)Vs[c][r];
        }
    for (int i = 0; i < D; ++i) {
        out.S[i] = (float)sig[i];
        out.Sd[i] = sig[i];
    }
    return out;
}

// Svd::recompose = U * diag(S) * Vt.
template <int D>
inline Mat<D> recompose(const Svd<D>& s) {
    Mat<D> US;
    for (int j = 0; j < D; ++j) US.c[j] = s.U.c[j] * s.S[j];
    return US * s.Vt;
}

// wgebra::inv::inv3 / inv4 — general inverse by cofactors (g2p_cdf.wgsl:236,242).
template <int N>
struct MatN {
    float m[N][N]; // m[row][col]
};
template <int N>
inline float detN(const MatN<N>& a);
template <>
inline float detN<3>(const MatN<3>& a) {
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
           a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
           a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}
inline float minor3(const MatN<4>& a, int r, int c) {
    MatN<3> s;
    int ii = 0;
    for (int i = 0; i < 4; ++i) {
        if (i == r) continue;
        int jj = 0;
        for (int j = 0; j < 4; ++j) {
            if (j == c) continue;
            s.m[ii][jj++] = a.m[i][j];
        }
        ++ii;
    }
    return detN<3>(s);
}
template <>
inline float detN<4>(const MatN<4>& a) {
    float d = 0.0f;
    for (int j = 0; j < 4; ++j) {
        float cof = minor3(a, 0, j);
        d += ((j & 1) ? -1.0f : 1.0f) * a.m[0][j] * cof;
    }
    return d;
}
inline MatN<3> invN(const MatN<3>& a) {
    float d = detN<3>(a);
    float id = 1.0f / d;
    MatN<3> r;
    r.m[0][0] = (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) * id;
    r.m[0][1] = (a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2]) * id;
    r.m[0][2] = (a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1]) * id;
    r.m[1][0] = (a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2]) * id;
    r.m[1][1] = (a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0]) * id;
    r.m[1][2] = (a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2]) * id;
    r.m[2][0] = (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]) * id;
    r.m[2][1] = (a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1]) * id;
    r.m[2][2] = (a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0]) * id;
    return r;
}
inline MatN<4> invN(const MatN<4>& a) {
    float d = detN<4>(a);
    float id = 1.0f / d;
    MatN<4> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float cof = minor3(a, j, i); // adjugate = transposed cofactors
            r.m[i][j] = (((i + j) & 1) ? -1.0f : 1.0f) * cof * id;
        }
    return r;
}

// ---------------------------------------------------------------------------------------
// Quadratic B-spline kernel (src/grid/kernel.wgsl)
// ---------------------------------------------------------------------------------------
template <int D>
struct Nbh;
template <>
struct Nbh<2> {
    static constexpr int LEN = 9; // kernel.wgsl:6
    static constexpr int SHIFTS[9][2] = {{2, 2}, {2, 0}, {2, 1}, {0, 2}, {0, 0},
                                         {0, 1}, {1, 2}, {1, 0}, {1, 1}}; // kernel.wgsl:7-17
};
template <>
struct Nbh<3> {
    static constexpr int LEN = 27; // kernel.wgsl:22
    static constexpr int SHIFTS[27][3] = { // kernel.wgsl:23-51
        {2, 2, 2}, {2, 0, 2}, {2, 1, 2}, {0, 2, 2}, {0, 0, 2}, {0, 1, 2}, {1, 2, 2}, {1, 0, 2}, {1, 1, 2},
        {2, 2, 0}, {2, 0, 0}, {2, 1, 0}, {0, 2, 0}, {0, 0, 0}, {0, 1, 0}, {1, 2, 0}, {1, 0, 0}, {1, 1, 0},
        {2, 2, 1}, {2, 0, 1}, {2, 1, 1}, {0, 2, 1}, {0, 0, 1}, {0, 1, 1}, {1, 2, 1}, {1, 0, 1}, {1, 1, 1}};
};

inline float inv_d(float cell_width) { return 4.0f / (cell_width * cell_width); } // kernel.wgsl:57-59

struct Weights3 {
    float w[3];
};
inline Weights3 eval_all(float x) { // kernel.wgsl:61-67
    return Weights3{{0.5f * (1.5f - x) * (1.5f - x), 0.75f - (x - 1.0f) * (x - 1.0f),
                     0.5f * (x - 0.5f) * (x - 0.5f)}};
}
// precompute_weights (kernel.wgsl:86-105): column k holds the three 1-D weights of axis k.
template <int D>
struct KernelWeights {
    Weights3 axis[D];
};
template <int D>
inline KernelWeights<D> precompute_weights(Vec<D> ref_elt_pos_minus_particle_pos, float h) {
    KernelWeights<D> w;
    for (int k = 0; k < D; ++k) w.axis[k] = eval_all(-ref_elt_pos_minus_particle_pos[k] / h);
    return w;
}
template <int D>
inline float weight_at(const KernelWeights<D>& w, const int* shift) {
    float r = w.axis[0].w[shift[0]] * w.axis[1].w[shift[1]];
    if constexpr (D == 3) r = r * w.axis[2].w[shift[2]];
    return r;
}

// ---------------------------------------------------------------------------------------
// Data (src/grid/grid.wgsl, src/solver/particle{2,3}d.wgsl, src/models/*.wgsl)
// ---------------------------------------------------------------------------------------
template <int D>
struct BlockVirtualId { // grid.wgsl:55-61
    int32_t id[D];
    bool operator==(const BlockVirtualId& o) const {
        for (int i = 0; i < D; ++i)
            if (id[i] != o.id[i]) return false;
        return true;
    }
};

struct NodeCdf { // grid.wgsl:233-240
    float distance;
    uint32_t affinities;
    uint32_t closest_id;
};
template <int D>
struct Node { // grid.wgsl:257-267
    Vec<D> momentum_velocity;
    float mass;
    NodeCdf cdf;
};
struct NodeLinkedList { // grid.wgsl:29-38
    uint32_t head, len;
};
template <int D>
struct HashMapEntry { // grid.wgsl:110-117
    uint32_t state;
    BlockVirtualId<D> key;
    uint32_t value;
};
template <int D>
struct ActiveBlockHeader { // grid.wgsl:215-219
    BlockVirtualId<D> virtual_id;
    uint32_t first_particle;
    uint32_t num_particles;
};

template <int D>
struct Cdf { // particle3d.wgsl:17-25
    Vec<D> normal;
    Vec<D> rigid_vel;
    float signed_distance;
    uint32_t affinity;
};
template <int D>
struct Dynamics { // particle3d.wgsl:7-15
    Vec<D> velocity;
    Mat<D> def_grad;
    Mat<D> affine;
    Cdf<D> cdf;
    float init_volume, init_radius, mass;
};
struct ElasticCoefficients { // linear_elasticity.wgsl:8-11
    float lambda, mu;
};
struct Plasticity { // drucker_prager.wgsl:8-16
    float ha, hb, hc, hd, lambda, mu;
};
struct PlasticState { // drucker_prager.wgsl:19-23
    float plastic_deformation_gradient_det, plastic_hardening, log_vol_gain;
};
struct Phase { // particle_update.wgsl:40-43
    float phase, max_stretch;
};

template <int D>
struct Pose { // wgebra sim2/sim3 with scale 1 (src_testbed/step.rs:88)
    Mat<D> R;
    Vec<D> t;
    Vec<D> mulPt(Vec<D> p) const { return R * p + t; }
    Vec<D> invMulPt(Vec<D> p) const { return transpose(R) * (p - t); }
};
template <int D>
struct AngVecT;
template <>
struct AngVecT<2> {
    using type = float;
};
template <>
struct AngVecT<3> {
    using type = Vec<3>;
};
template <int D>
struct Velocity { // wgrapier body.wgsl (not vendored)
    Vec<D> linear;
    typename AngVecT<D>::type angular;
};
template <int D>
struct MassProperties {
    Mat<D> inv_inertia; // 2D: scalar in at(0,0)
    Vec<D> inv_mass;
    Vec<D> com;
};
template <int D>
struct IntegerImpulse { // rigid_impulses.wgsl:12-47
    Vec<D> com;
    int32_t linear[D];
    int32_t angular[3]; // 2D uses [0]
};
template <int D>
struct Shape {
    uint32_t type;
    Vec<D> a, b;
    float radius;
};

// ---------------------------------------------------------------------------------------
// wgrapier body.wgsl contracts
// ---------------------------------------------------------------------------------------
inline Vec<2> velocity_at_point(Vec<2> com, const Velocity<2>& vel, Vec<2> pt) {
    Vec<2> d = pt - com;
    return vel.linear + Vec<2>{{-d[1], d[0]}} * vel.angular;
}
inline Vec<3> velocity_at_point(Vec<3> com, const Velocity<3>& vel, Vec<3> pt) {
    return vel.linear + cross(vel.angular, pt - com);
}

inline Mat<2> rotation_from(const float* rot, std::integral_constant<int, 2>) {
    Mat<2> R;
    R.at(0, 0) = rot[0];
    R.at(0, 1) = -rot[1];
    R.at(1, 0) = rot[1];
    R.at(1, 1) = rot[0];
    return R;
}
inline Mat<3> rotation_from(const float* q, std::integral_constant<int, 3>) {
    float i = q[0], j = q[1], k = q[2], w = q[3];
    Mat<3> R;
    R.at(0, 0) = 1.0f - 2.0f * (j * j + k * k);
    R.at(0, 1) = 2.0f * (i * j - k * w);
    R.at(0, 2) = 2.0f * (i * k + j * w);
    R.at(1, 0) = 2.0f * (i * j + k * w);
    R.at(1, 1) = 1.0f - 2.0f * (i * i + k * k);
    R.at(1, 2) = 2.0f * (j * k - i * w);
    R.at(2, 0) = 2.0f * (i * k - j * w);
    R.at(2, 1) = 2.0f * (j * k + i * w);
    R.at(2, 2) = 1.0f - 2.0f * (i * i + j * j);
    return R;
}

// ---------------------------------------------------------------------------------------
// wgparry Shape::projectPointOnBoundary (not vendored): parry's project_local_point with
// solid = false, mapped back to world space.
// ---------------------------------------------------------------------------------------
template <int D>
struct ProjectionResult {
    Vec<D> point;
    bool is_inside;
};

template <int D>
inline ProjectionResult<D> project_local_point_on_boundary(const Shape<D>& s, Vec<D> pt) {
    ProjectionResult<D> out;
    if (s.type == B200MPM_SHAPE_BALL) {
        float d2 = dot(pt, pt);
        out.is_inside = d2 <= s.radius * s.radius;
        if (d2 == 0.0f) {
            out.point = Vec<D>::zero();
            out.point[1] = s.radius;
        } else {
            out.point = pt * (s.radius / std::sqrt(d2));
        }
        return out;
    }
    if (s.type == B200MPM_SHAPE_CUBOID) {
        // parry Aabb::do_project_local_point(pt, solid = false) with mins = -he, maxs = he.
        Vec<D> mins_pt, pt_maxs, shift;
        bool inside = true;
        for (int i = 0; i < D; ++i) {
            mins_pt[i] = -s.a[i] - pt[i];
            pt_maxs[i] = pt[i] - s.a[i];
            shift[i] = std::max(mins_pt[i], 0.0f) - std::max(pt_maxs[i], 0.0f);
            if (shift[i] != 0.0f) inside = false;
        }
        out.is_inside = inside;
        if (!inside) {
            out.point = pt + shift;
            return out;
        }
        float best = -3.402823466e38f;
        bool is_mins = false;
        int best_id = 0;
        for (int i = 0; i < D; ++i) {
            if (mins_pt[i] < pt_maxs[i]) {
                if (pt_maxs[i] > best) {
                    best_id = i;
                    is_mins = false;
                    best = pt_maxs[i];
                }
            } else if (mins_pt[i] > best) {
                best_id = i;
                is_mins = true;
                best = mins_pt[i];
            }
        }
        out.point = pt;
        out.point[best_id] = pt[best_id] + (is_mins ? best : -best);
        return out;
    }
    // Capsule: project on the segment, then on the sphere of `radius` around that point.
    Vec<D> ab = s.b - s.a;
    Vec<D> ap = pt - s.a;
    float ab_ap = dot(ab, ap);
    float sqnab = dot(ab, ab);
    Vec<D> seg_pt;
    if (ab_ap <= 0.0f) {
        seg_pt = s.a;
    } else if (ab_ap >= sqnab) {
        seg_pt = s.b;
    } else {
        seg_pt = s.a + ab * (ab_ap / sqnab);
    }
    Vec<D> dproj = pt - seg_pt;
    float dist = length(dproj);
    if (dist > 1.1920929e-7f) {
        out.is_inside = dist <= s.radius;
        out.point = seg_pt + dproj * (s.radius / dist);
    } else {
        // Point on the segment: any direction orthogonal to the segment.
        Vec<D> dir = Vec<D>::zero();
        float n = std::sqrt(sqnab);
        if (n > 0.0f) {
            if constexpr (D == 2) {
                dir[0] = -ab[1] / n;
                dir[1] = ab[0] / n;
            } else {
                Vec<3> u = ab / n;
                Vec<3> e = (std::fabs(u[0]) < 0.9f) ? Vec<3>{{1, 0, 0}} : Vec<3>{{0, 1, 0}};
                Vec<3> o = cross(u, e);
                dir = o / length(o);
            }
        } else {
            dir[1] = 1.0f;
        }
        out.is_inside = true;
        out.point = seg_pt + dir * s.radius;
    }
    return out;
}

template <int D>
inline ProjectionResult<D> project_point_on_boundary(const Shape<D>& s, const Pose<D>& pose, Vec<D> pt) {
    ProjectionResult<D> loc = project_local_point_on_boundary(s, pose.invMulPt(pt));
    loc.point = pose.mulPt(loc.point);
    return loc;
}

// ---------------------------------------------------------------------------------------
// CPIC helpers (grid.wgsl:230-255, 390-404)
// ---------------------------------------------------------------------------------------
constexpr uint32_t AFFINITY_BITS_MASK = 0x0000ffffu;
constexpr uint32_t SIGN_BITS_SHIFT = 16;
inline bool affinity_bit(uint32_t i, uint32_t a) { return (a & (1u << i)) != 0; }
inline bool sign_bit(uint32_t i, uint32_t a) { return ((a >> SIGN_BITS_SHIFT) & (1u << i)) != 0; }
inline bool affinities_are_compatible(uint32_t a1, uint32_t a2) {
    uint32_t common = a1 & a2 & AFFINITY_BITS_MASK;
    uint32_t s1 = (a1 >> SIGN_BITS_SHIFT) & common;
    uint32_t s2 = (a2 >> SIGN_BITS_SHIFT) & common;
    return s1 == s2;
}
template <int D>
inline Vec<D> project_velocity(Vec<D> vel, Vec<D> n) { // grid.wgsl:390-404
    float normal_vel = dot(vel, n);
    if (normal_vel < 0.0f) {
        const float friction = 20.0f;
        Vec<D> tangent_vel = vel - n * normal_vel;
        float tangent_vel_len = length(tangent_vel);
        Vec<D> dir = (tangent_vel_len > 1.0e-8f) ? tangent_vel / tangent_vel_len : Vec<D>::zero();
        return dir * std::max(0.0f, tangent_vel_len + friction * normal_vel);
    }
    return vel;
}

// ---------------------------------------------------------------------------------------
// Hash map (grid.wgsl:83-184)
// ---------------------------------------------------------------------------------------
inline uint32_t pack_key(const BlockVirtualId<2>& k) { // grid.wgsl:83-86
    return ((uint32_t)(k.id[0] + 0x00007fff) & 0x0000ffffu) | (((uint32_t)(k.id[1] + 0x00007fff) & 0x0000ffffu) << 16);
}
inline uint32_t pack_key(const BlockVirtualId<3>& k) { // grid.wgsl:88-95
    return ((uint32_t)(k.id[0] + 0x000003ff) & 0x000007ffu) | (((uint32_t)(k.id[1] + 0x000001ff) & 0x000003ffu) << 11) |
           (((uint32_t)(k.id[2] + 0x000003ff) & 0x000007ffu) << 21);
}
inline uint32_t hash(uint32_t packed_key) { // grid.wgsl:98-105
    uint32_t key = packed_key;
    key *= 0xcc9e2d51u;
    key = (key << 15) | (key >> 17);
    key *= 0x1b873593u;
    return key;
}

// ---------------------------------------------------------------------------------------
// The simulation state + one method per reference kernel.
// ---------------------------------------------------------------------------------------
template <int D>
struct Sim {
    static constexpr int BLOCK = (D == 2) ? 8 : 4; // grid.wgsl:43, particle{2,3}d.wgsl
    static constexpr int TILE = BLOCK + 2; // p2g.wgsl:31,38
    static constexpr int NUM_ASSOC_BLOCKS = (D == 2) ? 4 : 8; // grid.wgsl:208-212
    static constexpr int NUM_SHARED_CELLS = (D == 2) ? TILE * TILE : TILE * TILE * TILE;

    // SimulationParams (params.wgsl)
    Vec<D> gravity;
    float dt;
    // Grid (grid.wgsl:222-228)
    uint32_t num_active_blocks = 0;
    float cell_width;
    uint32_t hmap_capacity;
    uint32_t capacity;
    bool overflowed = false;

    std::vector<HashMapEntry<D>> hmap_entries;
    std::vector<ActiveBlockHeader<D>> active_blocks;
    std::vector<Node<D>> nodes;
    std::vector<NodeLinkedList> nodes_linked_lists;
    std::vector<uint32_t> scan_values;

    // Particles
    std::vector<Vec<D>> particles_pos;
    std::vector<Dynamics<D>> particles_dyn;
    std::vector<ElasticCoefficients> constitutive_model;
    std::vector<Plasticity> plasticity;
    std::vector<PlasticState> plastic_state;
    std::vector<Phase> phases;
    std::vector<uint32_t> model_kind; // additive selector, 0 = reference behaviour
    std::vector<uint32_t> sorted_particle_ids;
    std::vector<uint32_t> particle_node_linked_lists;

    // Bodies (<= 16)
    std::vector<Shape<D>> collision_shapes;
    std::vector<Pose<D>> poses;
    std::vector<float> pose_rot_raw; // 4 per body, the representation handed back to the host
    std::vector<Velocity<D>> body_vels;
    std::vector<MassProperties<D>> local_mprops;
    std::vector<MassProperties<D>> mprops;
    std::vector<IntegerImpulse<D>> body_impulses;

    // Rigid particles: sample points of trimesh (3D) / polyline (2D) colliders (GpuRigidParticles, particle3d.rs:82-88,
    // particle2d.rs:62-68) and the collider vertices they refer to (ShapeBuffers of wgrapier).
    std::vector<Vec<D>> rigid_local_pts, rigid_world_pts; // sample points
    std::vector<std::array<uint32_t, 4>> rigid_ids; // (vertex a, b, c, collider); 2D: (a, b, -, collider)
    std::vector<uint32_t> rigid_needs_block; // one bit per sample point
    std::vector<uint32_t> rigid_node_linked_lists; // next pointers
    std::vector<NodeLinkedList> nodes_rigid_linked_lists;
    std::vector<Vec<D>> vertex_local_pts, vertex_world_pts;
    std::vector<uint32_t> vertex_collider_ids;

    size_t num_particles() const { return particles_pos.size(); }
    size_t num_bodies() const { return collision_shapes.size(); }

    // GpuGrid::with_capacity (grid.rs:281-331)
    void init_grid(uint32_t cap, float h) {
        uint32_t c = 1;
        while (c < cap) c <<= 1;
        capacity = c;
        hmap_capacity = c;
        cell_width = h;
        hmap_entries.assign(c, HashMapEntry<D>{NONE, {}, 0});
        active_blocks.assign(c, ActiveBlockHeader<D>{});
        nodes.assign((size_t)c * NUM_CELL_PER_BLOCK, Node<D>{Vec<D>::zero(), 0.0f, {0.0f, 0, NONE}});
        nodes_linked_lists.assign((size_t)c * NUM_CELL_PER_BLOCK, NodeLinkedList{NONE, 0});
        nodes_rigid_linked_lists.assign((size_t)c * NUM_CELL_PER_BLOCK, NodeLinkedList{NONE, 0});
        scan_values.assign(c, 0);
    }

    // ---- position -> cell -> block (grid.wgsl:269-298, particle3d.wgsl:37-57) ----------
    BlockVirtualId<D> block_associated_to_point(Vec<D> pt) const {
        BlockVirtualId<D> b;
        for (int i = 0; i < D; ++i) {
            float assoc_cell = std::nearbyintf(pt[i] / cell_width) - 1.0f; // round = ties-to-even
            float assoc_block = std::floor(assoc_cell / (float)BLOCK);
            b.id[i] = (int32_t)assoc_block;
        }
        return b;
    }
    void blocks_associated_to_block(const BlockVirtualId<D>& block, BlockVirtualId<D>* out) const { // grid.wgsl:300-320
        if constexpr (D == 2) {
            const int s[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
            for (int k = 0; k < 4; ++k)
                for (int i = 0; i < 2; ++i) out[k].id[i] = block.id[i] + s[k][i];
        } else {
            const int s[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 1}, {1, 0, 0}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1}};
            for (int k = 0; k < 8; ++k)
                for (int i = 0; i < 3; ++i) out[k].id[i] = block.id[i] + s[k][i];
        }
    }
    void associated_cell_index_in_block_off_by_one(Vec<D> pt, uint32_t* out) const { // particle3d.wgsl:41-45
        for (int i = 0; i < D; ++i) {
            float assoc_cell = std::nearbyintf(pt[i] / cell_width) - 1.0f;
            float assoc_block = std::floor(assoc_cell / (float)BLOCK) * (float)BLOCK;
            out[i] = (uint32_t)(assoc_cell - assoc_block);
        }
    }
    Vec<D> dir_to_associated_grid_node(Vec<D> pt) const { // particle3d.wgsl:47-57
        Vec<D> r;
        for (int i = 0; i < D; ++i) r[i] = (std::nearbyintf(pt[i] / cell_width) - 1.0f) * cell_width - pt[i];
        return r;
    }
    static uint32_t node_id(uint32_t block_physical_id, const uint32_t* shift) { // grid.wgsl:341-349
        if constexpr (D == 2) return block_physical_id + shift[0] + shift[1] * 8;
        else return block_physical_id + shift[0] + shift[1] * 4 + shift[2] * 16;
    }

    // ---- hash map ------------------------------------------------------------------
    uint32_t insertion_index(const BlockVirtualId<D>& key) { // grid.wgsl:121-164
        uint32_t packed_key = pack_key(key);
        uint32_t slot = hash(packed_key) & (hmap_capacity - 1u);
        for (uint32_t k = 0; k < hmap_capacity; ++k) {
            HashMapEntry<D>& e = hmap_entries[slot];
            if (e.state == NONE) {
                e.state = packed_key;
                e.key = key;
                return slot;
            } else if (e.state == packed_key) {
                return NONE;
            }
            slot = (slot + 1u) % hmap_capacity & (hmap_capacity - 1u);
        }
        overflowed = true; // table full: the block is dropped silently (grid.wgsl:126-128)
        return NONE;
    }
    uint32_t find_block_header_id(const BlockVirtualId<D>& key) const { // grid.wgsl:167-184
        uint32_t packed_key = pack_key(key);
        uint32_t slot = hash(packed_key) & (hmap_capacity - 1u);
        for (uint32_t k = 0; k < hmap_capacity; ++k) { // (the reference loops forever on a full table)
            const HashMapEntry<D>& e = hmap_entries[slot];
            if (e.state == packed_key) return e.value;
            if (e.state == NONE) return NONE;
            slot = (slot + 1u) & (hmap_capacity - 1u);
        }
        return NONE;
    }
    void mark_block_as_active(const BlockVirtualId<D>& block) { // grid.wgsl:323-334
        uint32_t slot = insertion_index(block);
        if (slot != NONE) {
            uint32_t block_header_id = num_active_blocks++;
            active_blocks[block_header_id] = ActiveBlockHeader<D>{block, 0u, 0u};
            hmap_entries[slot].value = block_header_id;
        }
    }

    // ---- "grid sort" pass: WgGrid::queue_sort (grid.rs:30-207) ---------------------------
    void reset_hmap() { // grid.wgsl:186-203
        for (uint32_t id = 0; id < hmap_capacity; ++id) hmap_entries[id] = HashMapEntry<D>{NONE, {}, 0};
        num_active_blocks = 0;
    }
    void touch_particle_blocks() { // sort.wgsl:26-36
        for (size_t id = 0; id < num_particles(); ++id) {
            BlockVirtualId<D> blocks[NUM_ASSOC_BLOCKS];
            blocks_associated_to_block(block_associated_to_point(particles_pos[id]), blocks);
            for (int i = 0; i < NUM_ASSOC_BLOCKS; ++i) mark_block_as_active(blocks[i]);
        }
    }
    void update_block_particle_count() { // sort.wgsl:89-99
        for (size_t id = 0; id < num_particles(); ++id) {
            uint32_t hid = find_block_header_id(block_associated_to_point(particles_pos[id]));
            if (hid == NONE) continue; // only after an overflow (undefined in the reference)
            active_blocks[hid].num_particles += 1u;
        }
    }
    void copy_particles_len_to_scan_value() { // sort.wgsl:101-107
        for (uint32_t id = 0; id < num_active_blocks; ++id) scan_values[id] = active_blocks[id].num_particles;
    }
    // WgPrefixSum::queue == eval_cpu semantics (prefix_sum.rs:71-83): exclusive scan over the
    // whole capacity-length buffer.
    static void prefix_sum(std::vector<uint32_t>& v) {
        if (v.empty()) return;
        for (size_t i = 0; i + 1 < v.size(); ++i) v[i + 1] += v[i];
        for (size_t i = v.size() - 1; i >= 1; --i) v[i] = v[i - 1];
        v[0] = 0;
    }
    void copy_scan_values_to_first_particles() { // sort.wgsl:109-115
        for (uint32_t id = 0; id < num_active_blocks; ++id) active_blocks[id].first_particle = scan_values[id];
    }
    void reset() { // grid.wgsl:362-379
        size_t num_nodes = (size_t)num_active_blocks * NUM_CELL_PER_BLOCK;
        for (size_t i = 0; i < num_nodes; ++i) {
            nodes[i].momentum_velocity = Vec<D>::zero();
            nodes[i].mass = 0.0f;
            nodes[i].cdf = NodeCdf{0.0f, 0, NONE};
            nodes_linked_lists[i] = NodeLinkedList{NONE, 0u};
        }
    }
    void finalize_particles_sort() { // sort.wgsl:117-137
        for (size_t id = 0; id < num_particles(); ++id) {
            Vec<D> pt = particles_pos[id];
            uint32_t hid = find_block_header_id(block_associated_to_point(pt));
            if (hid == NONE) continue;
            uint32_t target_index = scan_values[hid]++;
            sorted_particle_ids[target_index] = (uint32_t)id;
            uint32_t local[D];
            associated_cell_index_in_block_off_by_one(pt, local);
            uint32_t node = node_id(hid * NUM_CELL_PER_BLOCK, local);
            uint32_t prev_head = nodes_linked_lists[node].head;
            nodes_linked_lists[node].head = (uint32_t)id;
            nodes_linked_lists[node].len += 1u;
            particle_node_linked_lists[id] = prev_head;
        }
    }
    void queue_sort() {
        reset_hmap();
        touch_particle_blocks();
        mark_rigid_particles_needing_block();
        touch_rigid_particle_blocks();
        update_block_particle_count();
        copy_particles_len_to_scan_value();
        prefix_sum(scan_values);
        copy_scan_values_to_first_particles();
        reset();
        finalize_particles_sort();
        sort_rigid_particles(); // MpmPipeline::queue_step (pipeline.rs:220-221)
    }

    // ---- rigid particles -----------------------------------------------------------------------------------
    void set_rigid_particles(const float* vertices, const uint32_t* vertex_colliders, size_t nv, const float* samples,
                             const uint32_t* ids4, size_t ns) {
        vertex_local_pts.resize(nv);
        vertex_world_pts.resize(nv);
        vertex_collider_ids.assign(vertex_colliders, vertex_colliders + nv);
        for (size_t i = 0; i < nv; ++i)
            for (int k = 0; k < D; ++k) vertex_local_pts[i][k] = vertices[3 * i + k];
        rigid_local_pts.resize(ns);
        rigid_world_pts.resize(ns);
        rigid_ids.resize(ns);
        for (size_t i = 0; i < ns; ++i) {
            for (int k = 0; k < D; ++k) rigid_local_pts[i][k] = samples[3 * i + k];
            rigid_ids[i] = {ids4[4 * i], ids4[4 * i + 1], ids4[4 * i + 2], ids4[4 * i + 3]};
        }
        rigid_needs_block.assign((ns + 31) / 32, 0u);
        rigid_node_linked_lists.assign(ns, NONE);
        transform_rigid_points();
    }
    // transform_sample_points / transform_shape_points (rigid_particle_update.wgsl:26-50)
    void transform_rigid_points() {
        for (size_t i = 0; i < rigid_local_pts.size(); ++i) rigid_world_pts[i] = poses[rigid_ids[i][3]].mulPt(rigid_local_pts[i]);
        for (size_t i = 0; i < vertex_local_pts.size(); ++i)
            vertex_world_pts[i] = poses[vertex_collider_ids[i]].mulPt(vertex_local_pts[i]);
    }
    // sort.wgsl:54-86: a sample point asks for its own block iff that block is missing while one of the other
    // blocks its stencil reaches exists.
    void mark_rigid_particles_needing_block() {
        for (size_t id = 0; id < rigid_world_pts.size(); ++id) {
            BlockVirtualId<D> blocks[NUM_ASSOC_BLOCKS];
            blocks_associated_to_block(block_associated_to_point(rigid_world_pts[id]), blocks);
            int i = 0;
            for (; i < NUM_ASSOC_BLOCKS; ++i)
                if (find_block_header_id(blocks[i]) != NONE) break;
            const uint32_t bit = 1u << (id % 32);
            if (i > 0 && i < NUM_ASSOC_BLOCKS) rigid_needs_block[id / 32] |= bit;
            else rigid_needs_block[id / 32] &= ~bit;
        }
    }
    void touch_rigid_particle_blocks() { // sort.wgsl:38-52
        for (size_t id = 0; id < rigid_world_pts.size(); ++id)
            if (rigid_needs_block[id / 32] & (1u << (id % 32))) mark_block_as_active(block_associated_to_point(rigid_world_pts[id]));
    }
    void sort_rigid_particles() { // sort.wgsl:139-161 (the node lists are reset by `reset`, grid.wgsl:362-379)
        for (uint32_t b = 0; b < num_active_blocks; ++b)
            for (uint32_t n = 0; n < NUM_CELL_PER_BLOCK; ++n) nodes_rigid_linked_lists[(size_t)b * NUM_CELL_PER_BLOCK + n] = NodeLinkedList{NONE, 0};
        for (uint32_t id = 0; id < (uint32_t)rigid_world_pts.size(); ++id) {
            const Vec<D> pt = rigid_world_pts[id];
            uint32_t hid = find_block_header_id(block_associated_to_point(pt));
            if (hid == NONE) continue; // cannot affect the simulation
            uint32_t local[3] = {0, 0, 0};
            associated_cell_index_in_block_off_by_one(pt, local);
            NodeLinkedList& list = nodes_rigid_linked_lists[node_id(hid * NUM_CELL_PER_BLOCK, local)];
            rigid_node_linked_lists[id] = list.head;
            list.head = id;
            list.len += 1;
        }
    }
    // One candidate primitive against one node (p2g_cdf.wgsl:116-190): a projection that falls strictly inside
    // the segment / on the face interior colours the node.
    bool project_on_primitive(uint32_t rigid_id, Vec<D> cell_pos, float& distance, bool& sign) const {
        const auto& ids = rigid_ids[rigid_id];
        if constexpr (D == 2) {
            const Vec<2> a = vertex_world_pts[ids[0]], b = vertex_world_pts[ids[1]];
            // wgparry Segment::projectLocalPoint (not vendored): clamp to the end points
            const Vec<2> ab = b - a, ap = cell_pos - a;
            const float ab_ap = dot(ab, ap), sqnab = dot(ab, ab);
            Vec<2> proj;
            if (ab_ap <= 0.0f) proj = a;
            else if (ab_ap >= sqnab) proj = b;
            else proj = a + ab * (ab_ap / sqnab);
            const bool ne_a = proj[0] != a[0] || proj[1] != a[1], ne_b = proj[0] != b[0] || proj[1] != b[1];
            if (!(ne_a && ne_b)) return false;
            const Vec<2> dpt = cell_pos - proj;
            distance = length(dpt);
            sign = (dpt[0] * -ab[1] + dpt[1] * ab[0]) < 0.0f;
            return true;
        } else {
            const Vec<3> a = vertex_world_pts[ids[0]], b = vertex_world_pts[ids[1]], c = vertex_world_pts[ids[2]];
            const Vec<3> ap = cell_pos - a, bp = cell_pos - b, cp = cell_pos - c;
            const Vec<3> ab = b - a, ac = c - a, bc = c - b;
            const Vec<3> n = cross(ab, ac);
            const float n_length = length(n);
            if (n_length != 0.0f && dot(cross(ab, n), ap) <= 0.0f && dot(cross(bc, n), bp) <= 0.0f && dot(cross(ac, n), cp) >= 0.0f) {
                const float signed_dist = dot(n, ap) / n_length;
                sign = signed_dist < 0.0f;
                distance = std::fabs(signed_dist);
                return true;
            }
            return false;
        }
    }
    // p2g_cdf (p2g_cdf.wgsl:51-114): every node gathers the rigid particles listed in the 3^D cells whose stencil
    // contains it, one list element per iteration, and merges the per-iteration result into its cdf.
    void p2g_cdf() {
        if (rigid_world_pts.empty()) return;
        for (uint32_t bid = 0; bid < num_active_blocks; ++bid) {
            const BlockVirtualId<D> vid = active_blocks[bid].virtual_id;
            for_each_tid([&](const uint32_t* tid) {
                const uint32_t global_id = node_id(bid * NUM_CELL_PER_BLOCK, tid);
                const Vec<D> cell_pos = cell_pos_of(vid, tid);
                NodeCdf node_cdf = nodes[global_id].cdf;
                // heads of the lists of the contributing cells: global cell = 4 vid + tid - shift, shift in {0,1,2}^D
                uint32_t heads[27];
                uint32_t max_len = 0;
                const int nsh = (D == 2) ? 9 : 27;
                for (int sh = 0; sh < nsh; ++sh) {
                    int shift[3] = {sh % 3, (sh / 3) % 3, sh / 9};
                    BlockVirtualId<D> nb;
                    uint32_t local[3] = {0, 0, 0};
                    for (int k = 0; k < D; ++k) {
                        const int cell = vid.id[k] * BLOCK + (int)tid[k] - shift[k];
                        const int blk = (cell >= 0) ? cell / BLOCK : -((-cell + BLOCK - 1) / BLOCK);
                        nb.id[k] = blk;
                        local[k] = (uint32_t)(cell - blk * BLOCK);
                    }
                    heads[sh] = NONE;
                    const uint32_t hid = find_block_header_id(nb);
                    if (hid != NONE) {
                        const NodeLinkedList& l = nodes_rigid_linked_lists[node_id(hid * NUM_CELL_PER_BLOCK, local)];
                        heads[sh] = l.head;
                        max_len = std::max(max_len, l.len);
                    }
                }
                for (uint32_t it = 0; it < max_len; ++it) {
                    NodeCdf result{1.0e10f, 0u, NONE};
                    for (int sh = 0; sh < nsh; ++sh) {
                        const uint32_t rid = heads[sh];
                        if (rid == NONE) continue;
                        heads[sh] = rigid_node_linked_lists[rid];
                        float distance;
                        bool sign;
                        if (project_on_primitive(rid, cell_pos, distance, sign)) {
                            const uint32_t collider_id = rigid_ids[rid][3];
                            result.affinities |= (1u << collider_id) | ((uint32_t)sign << (collider_id + 16));
                            if (distance < result.distance) {
                                result.distance = distance;
                                result.closest_id = collider_id;
                            }
                        }
                    }
                    if (result.closest_id != NONE) {
                        node_cdf.affinities |= result.affinities;
                        if (result.distance < node_cdf.distance) {
                            node_cdf.distance = result.distance;
                            node_cdf.closest_id = result.closest_id;
                        }
                    }
                }
                nodes[global_id].cdf = node_cdf;
            });
        }
    }

    // ---- collide (collision/collide.wgsl:23-55) + grid_update_cdf (grid_update_cdf.wgsl:16-39) --
    NodeCdf collide(Vec<D> point) const {
        const float MAX_FLT = 1.0e10f;
        NodeCdf cdf{MAX_FLT, 0u, NONE};
        float dist_cap = cell_width * 1.5f;
        for (uint32_t i = 0; i < (uint32_t)num_bodies(); ++i) {
            if (collision_shapes[i].type == B200MPM_SHAPE_TRIMESH || collision_shapes[i].type == B200MPM_SHAPE_POLYLINE) continue;
            ProjectionResult<D> proj = project_point_on_boundary(collision_shapes[i], poses[i], point);
            Vec<D> dpt = proj.point - point;
            bool all_le = true;
            for (int k = 0; k < D; ++k) all_le = all_le && (std::fabs(dpt[k]) <= dist_cap);
            if (proj.is_inside || all_le) {
                float dist = length(dpt);
                cdf.closest_id = (dist < cdf.distance) ? i : cdf.closest_id;
                cdf.distance = std::min(cdf.distance, dist);
                cdf.affinities |= (proj.is_inside ? 0x00010001u : 0x00000001u) << i;
            }
        }
        return cdf;
    }
    Vec<D> cell_pos_of(const BlockVirtualId<D>& vid, const uint32_t* tid) const {
        Vec<D> p;
        for (int i = 0; i < D; ++i) p[i] = (float)(vid.id[i] * BLOCK + (int32_t)tid[i]) * cell_width;
        return p;
    }
    template <class F>
    static void for_each_tid(F&& f) { // one invocation per node of a 4x4x4 / 8x8 workgroup
        uint32_t tid[3] = {0, 0, 0};
        if constexpr (D == 2) {
            for (tid[1] = 0; tid[1] < 8; ++tid[1])
                for (tid[0] = 0; tid[0] < 8; ++tid[0]) f(tid);
        } else {
            for (tid[2] = 0; tid[2] < 4; ++tid[2])
                for (tid[1] = 0; tid[1] < 4; ++tid[1])
                    for (tid[0] = 0; tid[0] < 4; ++tid[0]) f(tid);
        }
    }
    void grid_update_cdf() {
#pragma omp parallel for schedule(static)
        for (int64_t bid = 0; bid < (int64_t)num_active_blocks; ++bid) {
            const BlockVirtualId<D> vid = active_blocks[bid].virtual_id;
            for_each_tid([&](const uint32_t* tid) {
                uint32_t global_id = node_id((uint32_t)bid * NUM_CELL_PER_BLOCK, tid);
                nodes[global_id].cdf = collide(cell_pos_of(vid, tid));
            });
        }
    }

    // ---- shared tile helper: G2P-style tile (g2p.wgsl:72-132, 251-268) -------------------
    static uint32_t flatten_g2p(const uint32_t* s) {
        if constexpr (D == 2) return s[0] + s[1] * 10;
        else return s[0] + s[1] * 6 + s[2] * 36;
    }
    // Fills tile[flat] with the global node id (or NONE) of every cell of the (BLOCK+2)^D tile
    // whose origin is block `vid` (octants with index 1 only hold their first two layers).
    void g2p_tile_node_ids(const BlockVirtualId<D>& vid, uint32_t* tile) const {
        for (int i = 0; i < NUM_SHARED_CELLS; ++i) tile[i] = NONE;
        const int noct = (D == 2) ? 4 : 8;
        for (int o = 0; o < noct; ++o) {
            int oc[3] = {0, 0, 0};
            if constexpr (D == 2) {
                oc[0] = o >> 1;
                oc[1] = o & 1;
            } else {
                oc[0] = o >> 2;
                oc[1] = (o >> 1) & 1;
                oc[2] = o & 1;
            }
            BlockVirtualId<D> nb = vid;
            for (int k = 0; k < D; ++k) nb.id[k] += oc[k];
            uint32_t hid = find_block_header_id(nb);
            if (hid == NONE) continue;
            for_each_tid([&](const uint32_t* tid) {
                for (int k = 0; k < D; ++k)
                    if (oc[k] == 1 && tid[k] > 1) return;
                uint32_t s[3] = {0, 0, 0};
                for (int k = 0; k < D; ++k) s[k] = (uint32_t)oc[k] * BLOCK + tid[k];
                tile[flatten_g2p(s)] = node_id(hid * NUM_CELL_PER_BLOCK, tid);
            });
        }
    }

    // ---- g2p_cdf (g2p_cdf.wgsl:39-63, 124-250) -------------------------------------------
    void particle_g2p_cdf(uint32_t particle_id, const NodeCdf* shared_nodes) {
        using NB = Nbh<D>;
        uint32_t particle_affinity = 0u;
        float affinity_signs[16];
        for (int i = 0; i < 16; ++i) affinity_signs[i] = 0.0f;
        const uint32_t prev_affinity = particles_dyn[particle_id].cdf.affinity;
        const Vec<D> particle_pos = particles_pos[particle_id];
        const Vec<D> ref = dir_to_associated_grid_node(particle_pos);
        const KernelWeights<D> w = precompute_weights(ref, cell_width);
        uint32_t local[3] = {0, 0, 0};
        associated_cell_index_in_block_off_by_one(particle_pos, local);
        const uint32_t packed = flatten_g2p(local);

        for (int i = 0; i < NB::LEN; ++i) {
            uint32_t sh[3] = {0, 0, 0};
            for (int k = 0; k < D; ++k) sh[k] = (uint32_t)NB::SHIFTS[i][k];
            const NodeCdf cell = shared_nodes[packed + flatten_g2p(sh)];
            particle_affinity |= cell.affinities & AFFINITY_BITS_MASK;
            const float weight = weight_at(w, NB::SHIFTS[i]);
            for (uint32_t ic = 0; ic < 16u; ++ic) {
                float compatible = affinity_bit(ic, cell.affinities) ? 1.0f : 0.0f;
                float sign = (sign_bit(ic, cell.affinities) /* && !shape_has_solid_interior */) ? -1.0f : 1.0f;
                affinity_signs[ic] += compatible * weight * sign * cell.distance;
            }
        }
        for (uint32_t ic = 0; ic < 16u; ++ic) {
            uint32_t mask = 1u << (ic + SIGN_BITS_SHIFT);
            if ((prev_affinity & (1u << ic)) == 0) {
                particle_affinity |= (affinity_signs[ic] < 0.0f) ? mask : 0u;
            } else {
                particle_affinity |= prev_affinity & mask;
            }
        }
        constexpr int Q = D + 1;
        MatN<Q> qtq;
        float qtu[Q];
        for (int a = 0; a < Q; ++a) {
            qtu[a] = 0.0f;
            for (int b = 0; b < Q; ++b) qtq.m[a][b] = 0.0f;
        }
        for (int i = 0; i < NB::LEN; ++i) {
            uint32_t sh[3] = {0, 0, 0};
            for (int k = 0; k < D; ++k) sh[k] = (uint32_t)NB::SHIFTS[i][k];
            const NodeCdf cell = shared_nodes[packed + flatten_g2p(sh)];
            float p[Q];
            for (int k = 0; k < D; ++k) p[k] = ref[k] + (float)NB::SHIFTS[i][k] * cell_width;
            p[D] = 1.0f;
            const float weight = weight_at(w, NB::SHIFTS[i]);
            uint32_t combined = cell.affinities & particle_affinity & AFFINITY_BITS_MASK;
            uint32_t sign_diff = ((cell.affinities >> SIGN_BITS_SHIFT) ^ (particle_affinity >> SIGN_BITS_SHIFT)) & combined;
            if (combined != 0u) {
                float dist = (sign_diff == 0u) ? cell.distance : -cell.distance;
                for (int b = 0; b < Q; ++b)
                    for (int a = 0; a < Q; ++a) qtq.m[a][b] += (p[a] * p[b]) * weight;
                for (int a = 0; a < Q; ++a) qtu[a] += p[a] * weight * dist;
            }
        }
        Cdf<D>& out = particles_dyn[particle_id].cdf;
        if (detN<Q>(qtq) > 1.0e-8f) {
            MatN<Q> inv = invN(qtq);
            float result[Q];
            for (int a = 0; a < Q; ++a) {
                float s = inv.m[a][0] * qtu[0];
                for (int b = 1; b < Q; ++b) s = s + inv.m[a][b] * qtu[b];
                result[a] = s;
            }
            Vec<D> n;
            for (int k = 0; k < D; ++k) n[k] = result[k];
            float len = length(n);
            if constexpr (D == 2) {
                n = (len > 1.0e-6f) ? n / len : Vec<D>::zero();
            } else {
                n = n / len;
            }
            out = Cdf<D>{n, Vec<D>::zero(), result[D], particle_affinity};
        } else {
            out = Cdf<D>{Vec<D>::zero(), Vec<D>::zero(), 0.0f, 0u}; // default_cdf()
        }
    }
    void g2p_cdf() {
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t bid = 0; bid < (int64_t)num_active_blocks; ++bid) {
            const ActiveBlockHeader<D>& ab = active_blocks[bid];
            if (ab.num_particles == 0) continue;
            uint32_t ids[NUM_SHARED_CELLS];
            NodeCdf shared_nodes[NUM_SHARED_CELLS];
            g2p_tile_node_ids(ab.virtual_id, ids);
            for (int i = 0; i < NUM_SHARED_CELLS; ++i)
                shared_nodes[i] = (ids[i] != NONE) ? nodes[ids[i]].cdf : NodeCdf{0.0f, 0, NONE};
            for (uint32_t s = ab.first_particle; s < ab.first_particle + ab.num_particles; ++s)
                particle_g2p_cdf(sorted_particle_ids[s], shared_nodes);
        }
    }

    // ---- p2g (p2g.wgsl:69-236) -------------------------------------------------------------
    static inline int32_t flt2int(float f) { // rigid_impulses.wgsl:52-54 (i32() saturates)
        float x = f * 1e5f;
        if (!(x == x)) return 0;
        if (x >= 2147483648.0f) return INT32_MAX;
        if (x <= -2147483648.0f) return INT32_MIN;
        return (int32_t)x;
    }
    static inline float int2flt(int32_t i) { return (float)i / 1e5f; }

    void p2g() {
        using NB = Nbh<D>;
        std::vector<int64_t> imp_lin((size_t)num_bodies() * 3, 0), imp_ang((size_t)num_bodies() * 3, 0);
#pragma omp parallel
        {
            std::vector<int64_t> my_lin((size_t)num_bodies() * 3, 0), my_ang((size_t)num_bodies() * 3, 0);
#pragma omp for schedule(dynamic, 16)
            for (int64_t bid = 0; bid < (int64_t)num_active_blocks; ++bid) {
                const BlockVirtualId<D> vid = active_blocks[bid].virtual_id;
                // fetch_nodes (p2g.wgsl:287-339): list heads of the tile whose origin is
                // (vid - 1 block) + 2 cells, i.e. global associated cells BLOCK*vid - 2 + s.
                uint32_t cur[NUM_SHARED_CELLS];
                uint32_t max_len = 0;
                for (int i = 0; i < NUM_SHARED_CELLS; ++i) cur[i] = NONE;
                const int noct = (D == 2) ? 4 : 8;
                for (int o = 0; o < noct; ++o) {
                    int oc[3] = {0, 0, 0};
                    if constexpr (D == 2) {
                        oc[0] = o >> 1;
                        oc[1] = o & 1;
                    } else {
                        oc[0] = o >> 2;
                        oc[1] = (o >> 1) & 1;
                        oc[2] = o & 1;
                    }
                    BlockVirtualId<D> nb = vid;
                    for (int k = 0; k < D; ++k) nb.id[k] += oc[k] - 1;
                    uint32_t hid = find_block_header_id(nb);
                    if (hid == NONE) continue;
                    for_each_tid([&](const uint32_t* tid) {
                        for (int k = 0; k < D; ++k)
                            if (oc[k] == 0 && tid[k] < (uint32_t)(BLOCK - 2)) return;
                        uint32_t s[3] = {0, 0, 0};
                        for (int k = 0; k < D; ++k) s[k] = (uint32_t)oc[k] * BLOCK + tid[k] - (uint32_t)(BLOCK - 2);
                        uint32_t gid = node_id(hid * NUM_CELL_PER_BLOCK, tid);
                        cur[flatten_g2p(s)] = nodes_linked_lists[gid].head;
                        max_len = std::max(max_len, nodes_linked_lists[gid].len);
                    });
                }
                Vec<D> acc_mv[NUM_CELL_PER_BLOCK];
                float acc_m[NUM_CELL_PER_BLOCK];
                Vec<D> acc_imp[NUM_CELL_PER_BLOCK];
                Vec<3> acc_ang[NUM_CELL_PER_BLOCK];
                for (uint32_t i = 0; i < NUM_CELL_PER_BLOCK; ++i) {
                    acc_mv[i] = Vec<D>::zero();
                    acc_m[i] = 0.0f;
                    acc_imp[i] = Vec<D>::zero();
                    acc_ang[i] = Vec<3>::zero();
                }
                // shared particle slots (fetch_next_particle, p2g.wgsl:341-396)
                struct Slot {
                    Vec<D> pos, vel, normal;
                    float mass;
                    Mat<D> affine;
                    uint32_t affinity;
                };
                std::vector<Slot> slots(NUM_SHARED_CELLS);
                for (uint32_t depth = 0; depth < max_len; ++depth) {
                    for (int i = 0; i < NUM_SHARED_CELLS; ++i) {
                        Slot& s = slots[i];
                        uint32_t pid = cur[i];
                        if (pid != NONE) {
                            const Dynamics<D>& dyn = particles_dyn[pid];
                            s.affinity = dyn.cdf.affinity;
                            s.normal = dyn.cdf.normal;
                            s.pos = particles_pos[pid];
                            s.affine = dyn.affine;
                            s.vel = dyn.velocity;
                            s.mass = dyn.mass;
                            cur[i] = particle_node_linked_lists[pid];
                        } else {
                            s.affinity = 0;
                            s.normal = Vec<D>::zero();
                            s.pos = Vec<D>::zero();
                            s.affine = Mat<D>::zero();
                            s.vel = Vec<D>::zero();
                            s.mass = 0.0f;
                        }
                    }
                    // p2g_step for each node thread (p2g.wgsl:158-236)
                    for_each_tid([&](const uint32_t* tid) {
                        uint32_t lin = node_id(0, tid);
                        uint32_t global_id = node_id((uint32_t)bid * NUM_CELL_PER_BLOCK, tid);
                        const uint32_t node_affinity = nodes[global_id].cdf.affinities;
                        const uint32_t collider_id = nodes[global_id].cdf.closest_id;
                        Vec<D> part_mv = Vec<D>::zero();
                        float part_m = 0.0f;
                        Vec<D> part_imp = Vec<D>::zero();
                        Vec<3> part_ang = Vec<3>::zero();
                        for (int i = 0; i < NB::LEN; ++i) {
                            uint32_t s[3] = {0, 0, 0};
                            for (int k = 0; k < D; ++k) s[k] = tid[k] + (uint32_t)NB::SHIFTS[i][k];
                            const Slot& sl = slots[flatten_g2p(s)];
                            Vec<D> ref = dir_to_associated_grid_node(sl.pos);
                            KernelWeights<D> w = precompute_weights(ref, cell_width);
                            int shift[3] = {0, 0, 0};
                            for (int k = 0; k < D; ++k) shift[k] = 2 - NB::SHIFTS[i][k];
                            Vec<D> momentum = sl.vel * sl.mass;
                            Vec<D> dpt;
                            for (int k = 0; k < D; ++k) dpt[k] = ref[k] + (float)shift[k] * cell_width;
                            float weight = weight_at(w, shift);
                            if (!affinities_are_compatible(node_affinity, sl.affinity)) {
                                if (collider_id != NONE) {
                                    Vec<D> body_com = body_impulses[collider_id].com;
                                    Vec<D> cell_center = dpt + sl.pos;
                                    Vec<D> body_pt_vel = velocity_at_point(body_com, body_vels[collider_id], cell_center);
                                    Vec<D> ghost = body_pt_vel + project_velocity(sl.vel - body_pt_vel, sl.normal);
                                    Vec<D> delta = (sl.vel - ghost) * (weight * sl.mass);
                                    Vec<D> lever = body_com - cell_center;
                                    if constexpr (D == 2) {
                                        part_ang[0] += dot(delta, Vec<2>{{lever[1], -lever[0]}});
                                    } else {
                                        part_ang = part_ang + cross(delta, lever);
                                    }
                                    part_imp = part_imp + delta;
                                    continue;
                                }
                            } else {
                                Vec<D> c = (sl.affine * dpt + momentum) * weight;
                                part_mv = part_mv + c;
                                part_m += sl.mass * weight;
                            }
                        }
                        acc_mv[lin] = acc_mv[lin] + part_mv;
                        acc_m[lin] += part_m;
                        acc_imp[lin] = acc_imp[lin] + part_imp;
                        acc_ang[lin] = acc_ang[lin] + part_ang;
                    });
                }
                for_each_tid([&](const uint32_t* tid) {
                    uint32_t lin = node_id(0, tid);
                    uint32_t global_id = node_id((uint32_t)bid * NUM_CELL_PER_BLOCK, tid);
                    nodes[global_id].momentum_velocity = acc_mv[lin];
                    nodes[global_id].mass = acc_m[lin];
                    uint32_t collider_id = nodes[global_id].cdf.closest_id;
                    if (collider_id != NONE) {
                        for (int k = 0; k < D; ++k) my_lin[collider_id * 3 + k] += flt2int(acc_imp[lin][k]);
                        if constexpr (D == 2) {
                            my_ang[collider_id * 3] += flt2int(acc_ang[lin][0]);
                        } else {
                            for (int k = 0; k < 3; ++k) my_ang[collider_id * 3 + k] += flt2int(acc_ang[lin][k]);
                        }
                    }
                });
            }
#pragma omp critical
            {
                for (size_t i = 0; i < my_lin.size(); ++i) {
                    imp_lin[i] += my_lin[i];
                    imp_ang[i] += my_ang[i];
                }
            }
        }
        for (size_t b = 0; b < num_bodies(); ++b) { // i32 atomicAdd wraps
            for (int k = 0; k < D; ++k)
                body_impulses[b].linear[k] = (int32_t)((uint32_t)body_impulses[b].linear[k] + (uint32_t)imp_lin[b * 3 + k]);
            for (int k = 0; k < 3; ++k)
                body_impulses[b].angular[k] = (int32_t)((uint32_t)body_impulses[b].angular[k] + (uint32_t)imp_ang[b * 3 + k]);
        }
    }

    // ---- grid_update (grid_update.wgsl:20-64) ----------------------------------------------
    void grid_update() {
        size_t num_nodes = (size_t)num_active_blocks * NUM_CELL_PER_BLOCK;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)num_nodes; ++i) {
            float mass = nodes[i].mass;
            float inv_mass = (mass > 0.0f) ? 1.0f / mass : 0.0f;
            Vec<D> velocity = (nodes[i].momentum_velocity + (mass * gravity) * dt) * inv_mass;
            float vel_limit = cell_width / dt;
            for (int k = 0; k < D; ++k) velocity[k] = std::min(std::max(velocity[k], -vel_limit), vel_limit);
            nodes[i].momentum_velocity = velocity;
        }
    }

    // ---- g2p (g2p.wgsl:44-238) ---------------------------------------------------------------
    void particle_g2p(uint32_t particle_id, const Node<D>* tile) {
        using NB = Nbh<D>;
        Vec<D> rigid_vel = Vec<D>::zero();
        Vec<D> mv = Vec<D>::zero();
        Mat<D> velocity_gradient = Mat<D>::zero();
        const Vec<D> particle_pos = particles_pos[particle_id];
        const Vec<D> particle_vel = particles_dyn[particle_id].velocity;
        const Cdf<D> particle_cdf = particles_dyn[particle_id].cdf;
        const float invd = inv_d(cell_width);
        const Vec<D> ref = dir_to_associated_grid_node(particle_pos);
        const KernelWeights<D> w = precompute_weights(ref, cell_width);
        uint32_t local[3] = {0, 0, 0};
        associated_cell_index_in_block_off_by_one(particle_pos, local);
        const uint32_t packed = flatten_g2p(local);
        for (int i = 0; i < NB::LEN; ++i) {
            uint32_t sh[3] = {0, 0, 0};
            for (int k = 0; k < D; ++k) sh[k] = (uint32_t)NB::SHIFTS[i][k];
            const Node<D>& cell = tile[packed + flatten_g2p(sh)];
            bool is_compatible = affinities_are_compatible(particle_cdf.affinity, cell.cdf.affinities);
            Vec<D> dpt;
            for (int k = 0; k < D; ++k) dpt[k] = ref[k] + (float)NB::SHIFTS[i][k] * cell_width;
            Vec<D> cpic_vel = cell.momentum_velocity;
            if (!is_compatible) {
                if (cell.cdf.closest_id != NONE) {
                    uint32_t cid = cell.cdf.closest_id;
                    Vec<D> cell_center = dpt + particle_pos;
                    Vec<D> body_pt_vel = velocity_at_point(mprops[cid].com, body_vels[cid], cell_center);
                    cpic_vel = body_pt_vel + project_velocity(particle_vel - body_pt_vel, particle_cdf.normal);
                } else {
                    cpic_vel = particle_vel;
                }
            }
            float weight = weight_at(w, NB::SHIFTS[i]);
            mv = mv + cpic_vel * weight;
            velocity_gradient = velocity_gradient + outer_product(cpic_vel, dpt) * (weight * invd);
        }
        for (uint32_t i = 0; i < 16u; ++i) {
            if (affinity_bit(i, particle_cdf.affinity) && i < (uint32_t)num_bodies())
                rigid_vel = rigid_vel + velocity_at_point(mprops[i].com, body_vels[i], particle_pos);
        }
        particles_dyn[particle_id].cdf.rigid_vel = rigid_vel;
        particles_dyn[particle_id].affine = velocity_gradient;
        particles_dyn[particle_id].velocity = mv;
    }
    void g2p() {
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t bid = 0; bid < (int64_t)num_active_blocks; ++bid) {
            const ActiveBlockHeader<D>& ab = active_blocks[bid];
            if (ab.num_particles == 0) continue;
            uint32_t ids[NUM_SHARED_CELLS];
            Node<D> tile[NUM_SHARED_CELLS];
            g2p_tile_node_ids(ab.virtual_id, ids);
            for (int i = 0; i < NUM_SHARED_CELLS; ++i)
                tile[i] = (ids[i] != NONE) ? nodes[ids[i]] : Node<D>{Vec<D>::zero(), 0.0f, {0.0f, 0, NONE}};
            for (uint32_t s = ab.first_particle; s < ab.first_particle + ab.num_particles; ++s)
                particle_g2p(sorted_particle_ids[s], tile);
        }
    }

    // ---- render hand-off (src_testbed/prep_vertex_buffer3d.wgsl:40-95, prep_vertex_buffer2d.wgsl:39-94) ------
    // inst: num_particles x 24 floats = {deformation 3 x vec4, position vec4, base_color vec4, color vec4};
    // writes xyz of the deformation columns and of the position, and color; reads base_color.
    void prep_vertex_buffer(uint32_t mode, float* inst) const {
        for (uint32_t id = 0; id < num_particles(); ++id) {
            float* o = inst + (size_t)id * 24;
            const Dynamics<D>& dyn = particles_dyn[id];
            if constexpr (D == 3) {
                for (int c = 0; c < 3; ++c)
                    for (int r = 0; r < 3; ++r) o[4 * c + r] = dyn.def_grad.at(r, c);
                for (int k = 0; k < 3; ++k) o[12 + k] = particles_pos[id][k];
            } else {
                o[0] = dyn.def_grad.at(0, 0), o[1] = dyn.def_grad.at(1, 0), o[2] = 0.0f;
                o[4] = dyn.def_grad.at(0, 1), o[5] = dyn.def_grad.at(1, 1), o[6] = 0.0f;
                o[8] = 0.0f, o[9] = 0.0f, o[10] = 1.0f;
                o[12] = particles_pos[id][0], o[13] = particles_pos[id][1], o[14] = 0.0f;
            }
            const float* base = o + 16;
            float* col = o + 20;
            for (int k = 0; k < 4; ++k) col[k] = base[k];
            if (mode == 2) { // VELOCITY
                for (int k = 0; k < D; ++k) col[k] = std::fabs(dyn.velocity[k]) * dt * 100.0f + 0.2f;
            } else if (mode == 1) { // VOLUME
                Svd<D> sv = svd(dyn.def_grad);
                for (int k = 0; k < D; ++k) col[k] = (1.0f - sv.S[k]) / 0.005f + 0.2f;
            } else if (mode == 3) { // CDF_NORMALS
                bool zero = true;
                for (int k = 0; k < D; ++k) zero = zero && dyn.cdf.normal[k] == 0.0f;
                for (int k = 0; k < 3; ++k) col[k] = (zero || k >= D) ? 0.0f : (dyn.cdf.normal[k] + 1.0f) / 2.0f;
            } else if (mode == 4) { // CDF_DISTANCES
                const float dd = dyn.cdf.signed_distance / (cell_width * 1.5f);
                col[0] = (dd > 0.0f) ? 0.0f : std::fabs(dd);
                col[1] = (dd > 0.0f) ? std::fabs(dd) : 0.0f;
                col[2] = 0.0f;
            } else if (mode == 5) { // CDF_SIGNS
                const uint32_t aff = dyn.cdf.affinity;
                const uint32_t a = (aff >> 16) & (aff & 0x0000ffffu);
                col[0] = (aff != 0u && a != 0u) ? 1.0f : 0.0f;
                col[1] = (aff != 0u && a == 0u) ? 1.0f : 0.0f;
                col[2] = 0.0f;
            }
        }
    }

    // ---- constitutive models (src/models/*.wgsl) -----------------------------------------------
    static Mat<D> kirchoff_stress_corotated(ElasticCoefficients model, const Mat<D>& F) { // linear_elasticity.wgsl:14-41
        Svd<D> s = svd(F);
        float j = s.S[0];
        for (int i = 1; i < D; ++i) j = j * s.S[i];
        float jm1 = j - 1.0f;
        if (exact_sigma_mode()) {
            double jd = s.Sd[0];
            for (int i = 1; i < D; ++i) jd *= s.Sd[i];
            jm1 = (float)(jd - 1.0);
            for (int i = 0; i < D; ++i) s.S[i] = (float)(s.Sd[i] - 1.0);
        } else {
            for (int i = 0; i < D; ++i) s.S[i] -= 1.0f;
        }
        float diag = model.lambda * jm1 * j;
        Mat<D> result = (recompose(s) * transpose(F)) * (2.0f * model.mu);
        for (int i = 0; i < D; ++i) result.at(i, i) += diag;
        return result;
    }
    static Mat<D> kirchoff_stress_neo_hookean(ElasticCoefficients model, const Mat<D>& F) { // neo_hookean_elasticity.wgsl:11-26
        float j = std::max(determinant(F), 1.0e-10f);
        float diag = model.lambda * std::log(j) - model.mu;
        Mat<D> stress = (F * transpose(F)) * model.mu;
        for (int i = 0; i < D; ++i) stress.at(i, i) += diag;
        return stress;
    }
    static float dp_alpha(const Plasticity& p, float q) { // drucker_prager.wgsl:25-29
        float angle = p.ha + (p.hb * q - p.hd) * std::exp(-p.hc * q);
        float s_angle = std::sin(angle);
        return std::sqrt(2.0f / 3.0f) * (2.0f * s_angle) / (3.0f - s_angle);
    }
    struct DpProjection {
        Vec<D> singular_values;
        float plastic_hardening;
        bool valid;
    };
    static DpProjection project_deformation_gradient(const Plasticity& p, Vec<D> sv, float log_vol_gain, float alpha,
                                                     const double* sv_exact = nullptr) {
        // drucker_prager.wgsl:43-64 (2D), 112-133 (3D)
        const float d = (float)D;
        Vec<D> strain;
        for (int i = 0; i < D; ++i) {
            float l = (sv_exact && exact_sigma_mode()) ? (float)std::log(sv_exact[i]) : std::log(sv[i]);
            strain[i] = l + log_vol_gain / d;
        }
        float strain_trace = strain[0];
        for (int i = 1; i < D; ++i) strain_trace = strain_trace + strain[i];
        Vec<D> dev;
        bool all_zero = true;
        for (int i = 0; i < D; ++i) {
            dev[i] = strain[i] - strain_trace / d;
            all_zero = all_zero && (dev[i] == 0.0f);
        }
        if (strain_trace > 0.0f || all_zero) return DpProjection{Vec<D>::splat(1.0f), length(strain), true};
        float dev_norm = length(dev);
        float gamma = dev_norm + (d * p.lambda + 2.0f * p.mu) / (2.0f * p.mu) * strain_trace * alpha;
        if (gamma <= 0.0f) return DpProjection{Vec<D>::zero(), 0.0f, false};
        Vec<D> h = strain - dev * (gamma / dev_norm);
        Vec<D> e;
        for (int i = 0; i < D; ++i) e[i] = std::exp(h[i]);
        return DpProjection{e, gamma, true};
    }
    static void dp_project(const Plasticity& p, PlasticState& state, Mat<D>& F) { // drucker_prager.wgsl:66-110, 135-158
        if (p.lambda == 0.0f) return;
        Svd<D> s = svd(F);
        float alpha = dp_alpha(p, state.plastic_hardening);
        DpProjection proj = project_deformation_gradient(p, s.S, state.log_vol_gain, alpha, s.Sd);
        if (proj.valid) {
            float prev_det = s.S[0], new_det = proj.singular_values[0];
            for (int i = 1; i < D; ++i) {
                prev_det = prev_det * s.S[i];
                new_det = new_det * proj.singular_values[i];
            }
            PlasticState ns;
            ns.plastic_deformation_gradient_det = state.plastic_deformation_gradient_det * prev_det / new_det;
            ns.log_vol_gain = state.log_vol_gain + std::log(prev_det) - std::log(new_det);
            ns.plastic_hardening = state.plastic_hardening + proj.plastic_hardening;
            Svd<D> r = s;
            r.S = proj.singular_values;
            F = recompose(r);
            state = ns;
        }
    }

    // ---- particle update (particle_update.wgsl:45-141) ------------------------------------------
    void particles_update() {
#pragma omp parallel for schedule(static)
        for (int64_t pid = 0; pid < (int64_t)num_particles(); ++pid) {
            const Dynamics<D> dynamics = particles_dyn[pid];
            const Vec<D> particle_pos = particles_pos[pid];
            Vec<D> new_vel = dynamics.velocity;
            if (dynamics.cdf.signed_distance < -0.05f * cell_width) {
                new_vel = dynamics.cdf.rigid_vel + project_velocity(new_vel - dynamics.cdf.rigid_vel, dynamics.cdf.normal);
            }
            if (length(new_vel) > cell_width / dt) {
                new_vel = new_vel / length(new_vel) * cell_width / dt;
            }
            Vec<D> new_pos = particle_pos + new_vel * dt;
            const float PENALTY_COEFF = 1.0e3f;
            if (dynamics.cdf.signed_distance < -0.05f * cell_width) {
                float corrected_dist = std::max(dynamics.cdf.signed_distance, -0.3f * cell_width);
                Vec<D> impulse = dynamics.cdf.normal * (dt * -corrected_dist * PENALTY_COEFF);
                new_vel = new_vel + impulse;
            }
            Mat<D> new_F = dynamics.def_grad + (dynamics.affine * dt) * dynamics.def_grad;
            float phase = phases[pid].phase;
            float max_stretch = phases[pid].max_stretch;
            if (phase > 0.0f && max_stretch > 0.0f) {
                Svd<D> s = svd(new_F);
                bool broken = false;
                for (int i = 0; i < D; ++i) broken = broken || (s.S[i] > max_stretch);
                if (broken) {
                    phases[pid].phase = 0.0f;
                    phase = 0.0f;
                }
            }
            if (phase == 0.0f) {
                dp_project(plasticity[pid], plastic_state[pid], new_F);
            }
            Mat<D> stress = (model_kind[pid] == B200MPM_MODEL_NEO_HOOKEAN)
                                ? kirchoff_stress_neo_hookean(constitutive_model[pid], new_F)
                                : kirchoff_stress_corotated(constitutive_model[pid], new_F);
            float invd = inv_d(cell_width);
            Mat<D> affine = dynamics.affine * dynamics.mass - stress * (dynamics.init_volume * invd * dt);
            particles_pos[pid] = new_pos;
            particles_dyn[pid].velocity = new_vel;
            particles_dyn[pid].def_grad = new_F;
            particles_dyn[pid].affine = affine;
        }
    }

    // ---- rigid bodies (rigid_impulses.wgsl:94-150; wgrapier body.wgsl contracts) ----------------
    void update_world_mass_properties() { // rigid_impulses.wgsl:139-150
        for (size_t id = 0; id < num_bodies(); ++id) {
            MassProperties<D> w;
            w.com = poses[id].mulPt(local_mprops[id].com);
            w.inv_mass = local_mprops[id].inv_mass;
            if constexpr (D == 2) {
                w.inv_inertia = local_mprops[id].inv_inertia;
            } else {
                w.inv_inertia = poses[id].R * local_mprops[id].inv_inertia * transpose(poses[id].R);
            }
            body_impulses[id].com = w.com;
            mprops[id] = w;
        }
    }
    void integrate_bodies() { // rigid_impulses.wgsl:94-137
        for (size_t id = 0; id < num_bodies(); ++id) {
            Vec<D> imp_lin;
            for (int k = 0; k < D; ++k) imp_lin[k] = int2flt(body_impulses[id].linear[k]);
            Vec<3> imp_ang{{int2flt(body_impulses[id].angular[0]), int2flt(body_impulses[id].angular[1]),
                            int2flt(body_impulses[id].angular[2])}};
            body_impulses[id].com = Vec<D>::zero();
            for (int k = 0; k < D; ++k) body_impulses[id].linear[k] = 0;
            for (int k = 0; k < 3; ++k) body_impulses[id].angular[k] = 0;

            Velocity<D> new_vel = body_vels[id];
            new_vel.linear = new_vel.linear + mul_comp(mprops[id].inv_mass, imp_lin); // applyImpulse
            float angvel_norm, imp_ang_norm;
            if constexpr (D == 2) {
                new_vel.angular += mprops[id].inv_inertia.at(0, 0) * imp_ang[0];
                angvel_norm = std::fabs(new_vel.angular);
                imp_ang_norm = std::fabs(imp_ang[0]);
            } else {
                new_vel.angular = new_vel.angular + mprops[id].inv_inertia * imp_ang;
                angvel_norm = length(new_vel.angular);
                imp_ang_norm = length(imp_ang);
            }
            float linvel_norm = length(new_vel.linear);
            float lin_limit = 0.1f * cell_width / dt;
            float ang_limit = 1.0f;
            if (length(imp_lin) != 0.0f || imp_ang_norm != 0.0f) {
                if (linvel_norm > lin_limit) new_vel.linear = new_vel.linear * (lin_limit / linvel_norm);
                if (angvel_norm > ang_limit) new_vel.angular = new_vel.angular * (ang_limit / angvel_norm);
            }
            // integrateVelocity: rotate about the world COM by exp(ang*dt), translate by lin*dt.
            Vec<D> com = poses[id].mulPt(local_mprops[id].com);
            float* rr = &pose_rot_raw[id * 4];
            Mat<D> dR;
            if constexpr (D == 2) {
                float a = new_vel.angular * dt;
                float c = std::cos(a), s = std::sin(a);
                float re = c * rr[0] - s * rr[1], im = s * rr[0] + c * rr[1];
                float n = std::sqrt(re * re + im * im);
                rr[0] = re / n;
                rr[1] = im / n;
                float dr[4] = {c, s, 0, 0};
                dR = rotation_from(dr, std::integral_constant<int, 2>{});
            } else {
                Vec<3> axis_angle = new_vel.angular * dt;
                float angle = length(axis_angle);
                float dq[4];
                if (angle > 0.0f) {
                    float s = std::sin(angle * 0.5f) / angle;
                    dq[0] = axis_angle[0] * s;
                    dq[1] = axis_angle[1] * s;
                    dq[2] = axis_angle[2] * s;
                    dq[3] = std::cos(angle * 0.5f);
                } else {
                    dq[0] = dq[1] = dq[2] = 0.0f;
                    dq[3] = 1.0f;
                }
                // q_new = dq * q
                float qi = rr[0], qj = rr[1], qk = rr[2], qw = rr[3];
                float ni = dq[3] * qi + dq[0] * qw + dq[1] * qk - dq[2] * qj;
                float nj = dq[3] * qj - dq[0] * qk + dq[1] * qw + dq[2] * qi;
                float nk = dq[3] * qk + dq[0] * qj - dq[1] * qi + dq[2] * qw;
                float nw = dq[3] * qw - dq[0] * qi - dq[1] * qj - dq[2] * qk;
                float n = std::sqrt(ni * ni + nj * nj + nk * nk + nw * nw);
                rr[0] = ni / n;
                rr[1] = nj / n;
                rr[2] = nk / n;
                rr[3] = nw / n;
                dR = rotation_from(dq, std::integral_constant<int, 3>{});
            }
            Pose<D> np;
            np.R = rotation_from(rr, std::integral_constant<int, D>{});
            np.t = dR * (poses[id].t - com) + new_vel.linear * dt + com;
            // gravity on bodies with non-zero inverse mass
            for (int k = 0; k < D; ++k)
                new_vel.linear[k] += gravity[k] * ((mprops[id].inv_mass[k] != 0.0f) ? 1.0f : 0.0f) * dt;
            body_vels[id] = new_vel;
            poses[id] = np;
        }
    }

    // ---- MpmPipeline::queue_step (pipeline.rs:195-281) -------------------------------------------
    void substep() {
        update_world_mass_properties(); // "update rigid particles"
        transform_rigid_points();
        queue_sort(); // "grid sort"
        grid_update_cdf(); // "grid_update_cdf"
        p2g_cdf(); // "p2g_cdf"
        g2p_cdf(); // "g2p_cdf"
        p2g(); // "p2g"
        grid_update(); // "grid_update"
        g2p(); // "g2p"
        particles_update(); // "particles_update"
        integrate_bodies(); // "integrate_bodies"
    }
};

} // namespace oracle
