// Per-frame arithmetic of the evaluation metrics (metrics.cu), written host+device so that tests/host_harness.cpp can run
// it on the CPU: FK of the 22 body joints from the folded joint regressor, Euclidean / Procrustes-aligned joint
// distances (empose/eval/metrics.py:19-66, 115-132).
#pragma once

#include <math.h>

#include "frame_math.h"

// These functions are big: keep them out of line on the device.  (A first version, fully inlined into metrics_kernel with
// all 22 joint rotations kept in a local array -- 255 registers, 3.8 KB of stack -- returned wrong joints below depth 1 of
// the kinematic tree on the B200; it was never run on the host, so whether that was the compiler or the code is open.
// This version is checked on the host AND on the device against the oracle: tests/test_oracle_metrics.py,
// tests/test_gpu_metrics.py.)
#if defined(__CUDACC__)
#define EMPOSE_HD_NOINLINE __host__ __device__ __noinline__
#else
#define EMPOSE_HD_NOINLINE inline
#endif

namespace empose {

struct MetricsParams {
    const float* j0; const float* jdirs; const int* parents;
    const float* pose; const float* shape; const float* pose_hat; const float* shape_hat;      // [R][66], [R][10]
    const float* joints; const float* joints_hat;                                              // or joints given directly [R][66]
    int R;
    int angle_local;                                                                           // 1: angles between local joint rotations
    float* eucl; float* eucl_pa; float* angle;                                                 // [R][22], [R][22], [R][21] (angle may be null)
};

// FK of the 22 body joints: posed joints and the global orientations with the root rotation removed
EMPOSE_HD_NOINLINE void fk_frame(const MetricsParams& p, const float* pose, const float* beta, float (&joints)[kJoints][3],
                         float (&orient)[kJoints][9]) {
    float grot[kJoints][9], jrest[kJoints][3];
    for (int i = 0; i < kPoseDim; ++i) {
        float acc = p.j0[i];
        for (int k = 0; k < kBetas; ++k) acc += p.jdirs[k * kPoseDim + i] * beta[k];
        jrest[i / 3][i % 3] = acc;
    }
    for (int j = 0; j < kJoints; ++j) {
        const int par = p.parents[j];
        float rj[9];
        rodrigues_fwd(pose + j * 3, rj);
        if (par < 0) {
            for (int e = 0; e < 9; ++e) { grot[j][e] = rj[e]; orient[j][e] = (e % 4 == 0) ? 1.0f : 0.0f; }
            for (int c = 0; c < 3; ++c) joints[j][c] = jrest[j][c];
            continue;
        }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                grot[j][r * 3 + c] = grot[par][r * 3] * rj[c] + grot[par][r * 3 + 1] * rj[3 + c] + grot[par][r * 3 + 2] * rj[6 + c];
                orient[j][r * 3 + c] = orient[par][r * 3] * rj[c] + orient[par][r * 3 + 1] * rj[3 + c] + orient[par][r * 3 + 2] * rj[6 + c];
            }
        for (int r = 0; r < 3; ++r)
            joints[j][r] = grot[par][r * 3] * (jrest[j][0] - jrest[par][0]) + grot[par][r * 3 + 1] * (jrest[j][1] - jrest[par][1]) +
                           grot[par][r * 3 + 2] * (jrest[j][2] - jrest[par][2]) + joints[par][r];
    }
}

// eigen-decomposition of a symmetric 3x3 matrix by cyclic Jacobi rotations (double): b -> eigenvalues on the diagonal,
// v -> eigenvectors as columns
EMPOSE_HD_NOINLINE void jacobi3(double (&b)[3][3], double (&v)[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(b[0][1]) + fabs(b[0][2]) + fabs(b[1][2]);
        if (off < 1e-30) break;
        for (int pq = 0; pq < 3; ++pq) {
            const int pi = pq == 2 ? 1 : 0, qi = pq == 0 ? 1 : 2;
            if (fabs(b[pi][qi]) < 1e-300) continue;
            const double theta = (b[qi][qi] - b[pi][pi]) / (2.0 * b[pi][qi]);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; ++k) {           // B <- B J
                const double bkp = b[k][pi], bkq = b[k][qi];
                b[k][pi] = c * bkp - s * bkq; b[k][qi] = s * bkp + c * bkq;
            }
            for (int k = 0; k < 3; ++k) {           // B <- J^T B
                const double bpk = b[pi][k], bqk = b[qi][k];
                b[pi][k] = c * bpk - s * bqk; b[qi][k] = s * bpk + c * bqk;
            }
            for (int k = 0; k < 3; ++k) {
                const double vkp = v[k][pi], vkq = v[k][qi];
                v[k][pi] = c * vkp - s * vkq; v[k][qi] = s * vkp + c * vkq;
            }
        }
    }
}

// distances of one frame: plain and after Procrustes alignment of `y` onto `x` (metrics.py:19-66, optimal scale)
EMPOSE_HD_NOINLINE void joint_distances(const float (&x)[kJoints][3], const float (&y)[kJoints][3], float* eucl, float* eucl_pa) {
    double mux[3] = {0, 0, 0}, muy[3] = {0, 0, 0};
    for (int j = 0; j < kJoints; ++j)
        for (int c = 0; c < 3; ++c) { mux[c] += x[j][c]; muy[c] += y[j][c]; }
    for (int c = 0; c < 3; ++c) { mux[c] /= kJoints; muy[c] /= kJoints; }
    double ssx = 0.0, ssy = 0.0, a[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int j = 0; j < kJoints; ++j) {
        double dx[3], dy[3];
        for (int c = 0; c < 3; ++c) { dx[c] = x[j][c] - mux[c]; dy[c] = y[j][c] - muy[c]; ssx += dx[c] * dx[c]; ssy += dy[c] * dy[c]; }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) a[r][c] += dx[r] * dy[c];
        const float d0 = x[j][0] - y[j][0], d1 = x[j][1] - y[j][1], d2 = x[j][2] - y[j][2];
        eucl[j] = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    }
    const double nx = sqrt(ssx), ny = sqrt(ssy);
    const double inv = (nx > 0.0 && ny > 0.0) ? 1.0 / (nx * ny) : 0.0;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) a[r][c] *= inv;          // A = X0^T Y0 of the unit-norm point sets
    // SVD A = U S V^T through the eigen-decomposition of A^T A
    double b[3][3], v[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) b[r][c] = a[0][r] * a[0][c] + a[1][r] * a[1][c] + a[2][r] * a[2][c];
    jacobi3(b, v);
    int order[3] = {0, 1, 2};                                 // singular values in descending order, as numpy returns them
    for (int i = 0; i < 2; ++i)
        for (int k = i + 1; k < 3; ++k)
            if (b[order[k]][order[k]] > b[order[i]][order[i]]) { const int t = order[i]; order[i] = order[k]; order[k] = t; }
    double s[3], u[3][3], vv[3][3];
    for (int i = 0; i < 3; ++i) {
        const int o = order[i];
        s[i] = sqrt(fmax(b[o][o], 0.0));
        for (int r = 0; r < 3; ++r) vv[r][i] = v[r][o];
    }
    for (int i = 0; i < 3; ++i) {
        if (s[i] > 1e-12 * s[0] && s[i] > 0.0) {
            for (int r = 0; r < 3; ++r) u[r][i] = (a[r][0] * vv[0][i] + a[r][1] * vv[1][i] + a[r][2] * vv[2][i]) / s[i];
        } else {        // rank-deficient: complete the basis (any unit vector orthogonal to the others gives the same product)
            const int i0 = (i + 1) % 3, i1 = (i + 2) % 3;
            u[0][i] = u[1][i0] * u[2][i1] - u[2][i0] * u[1][i1];
            u[1][i] = u[2][i0] * u[0][i1] - u[0][i0] * u[2][i1];
            u[2][i] = u[0][i0] * u[1][i1] - u[1][i0] * u[0][i1];
        }
    }
    auto det3 = [](const double (&m)[3][3]) {
        return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
               m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    };
    if (det3(vv) * det3(u) < 0.0) {                           // T = V U^T must be a rotation (metrics.py:52-56)
        for (int r = 0; r < 3; ++r) vv[r][2] = -vv[r][2];
        s[2] = -s[2];
    }
    double t[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) t[r][c] = vv[r][0] * u[c][0] + vv[r][1] * u[c][1] + vv[r][2] * u[c][2];
    const double scale = ny > 0.0 ? nx * (s[0] + s[1] + s[2]) / ny : 0.0;      // normX * traceTA applied to Y0 / normY
    for (int j = 0; j < kJoints; ++j) {
        const double dy[3] = {y[j][0] - muy[0], y[j][1] - muy[1], y[j][2] - muy[2]};
        double acc = 0.0;
        for (int c = 0; c < 3; ++c) {
            const double z = scale * (dy[0] * t[0][c] + dy[1] * t[1][c] + dy[2] * t[2][c]) + mux[c];
            const double d = x[j][c] - z;
            acc += d * d;
        }
        eucl_pa[j] = (float)sqrt(acc);
    }
}


}  // namespace empose
