// Per-frame SMPL-H sub-model -> sensor frames -> residual, forward and hand-derived reverse pass.
//
// This is the arithmetic the reference obtains from (i) the third-party BodyModel call at
// empose/bodymodels/smpl.py:121, (ii) VirtualMarkerHelper.get_virtual_pos_and_rot
// (empose/data/virtual_sensors.py:85-96, empose/helpers/utils.py:126-146), (iii) the offset
// application at empose/nn/models.py:478-479, (iv) reconstruction_loss (empose/nn/loss.py:23-41)
// and (v) torch.autograd's backward through all of it (empose/nn/models.py:576-579), restricted to
// the minimal sub-model prepared by submodel.py (22 joints, ~84 vertices, 12 sensors).
//
// The code is written as PHASES over a per-frame scratch state: each phase is a loop
// `for (i = lane; i < n; i += lanes)` with no intra-phase dependencies, so a group of `lanes`
// GPU threads runs a phase cooperatively and synchronises between phases, while the CPU test
// harness (tests/host_harness.cpp) runs the very same functions with lanes = 1.  Templated on the
// scalar type so the derivation can be checked in double on the host.
//
// The pose-blend contraction (189 features x vertex coordinates) is NOT done here: it is a dense
// GEMM over frames and runs on the tensor cores (pose_blend jobs in plan.cu); this file consumes
// its result (`vp_off`) and produces its reverse-mode input (`dvp`).
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define EMPOSE_HD __host__ __device__ __forceinline__
#else
#define EMPOSE_HD inline
#endif

namespace empose {

constexpr int kJoints = 22;           // root + 21 body joints (reference configuration.py:104)
constexpr int kPoseDim = 66;
constexpr int kBetas = 10;            // reference configuration.py:107
constexpr int kSensors = 12;          // reference configuration.py:32-34
constexpr int kPoseFeat = 189;        // 21 joints x 9 rotation entries
// Feature row of the blend GEMM: [vec(R_1..R_21 - I) (189) | 0 0 0 | beta (10) | 0 ...], so that ONE contraction gives
// S beta + P pf (the blended rest vertices minus the template, reference smpl.py:121 steps 1 and 4) and, in 66 extra
// output columns, Jdirs beta; v_template and J0 are the bias of that GEMM, i.e. they are added ONCE, in fp32, after
// the accumulation (metre-sized constants inside a tensor-core accumulator would cost the millimetre-sized terms
// their low bits).  Padded to a multiple of 32 floats = one 128B swizzle row.
constexpr int kFeatBeta = 192;        // first shape column of the feature row
constexpr int kFeatK = 202;           // real K of the blend GEMM
constexpr int kPoseFeatPad = 224;     // padded K
// The transposed contraction returns [dE/dpf (189) | 0 0 0 | dE/dbeta (10)] per frame in rows of kPoseFeatPad floats.
constexpr int kJrestLd = 68;          // row pitch of the rest-joint / dE/dJ buffers (66 values, 16-byte aligned rows)
constexpr int kDjLdHalf = 72;         // row pitch of the dE/dJ buffer when it holds fp16 elements
// fp16 form of the blend GEMMs: a value x is split as hi = fp16(x), lo = fp16((x - hi) * 2^11) -- 22 mantissa bits, like the
// tf32 split, at twice the tensor rate and half the bytes; the weights carry the matching 2^-11 (model.cu).  dE/dvp and
// dE/dJ (hundreds at most: 1 / edge length) are stored times kDvpScale so that even a very fine mesh stays inside fp16.
constexpr float kSplitLoScale = 2048.0f;
constexpr float kDvpScale = 0.125f;
constexpr int kMaxVp = 448;           // padded 3*Vs, supports sub-meshes of up to 144 vertices (12 sensors x 12 slots)
constexpr int kMaxDegree = 12;
constexpr int kSplitDegree = 7;        // sensor valence up to which the sensor phase is split over (sensor, face) items
constexpr int kMaxVj = 44;            // chunks of the joint->vertex lists (their partial sums alias dgr .. dj below)

// Packed sub-model constants (pointers into device or host memory).  Layouts: see submodel.py.
struct SubModel {
    int n_verts, vp_dim, n_faces, max_degree, n_skin;
    const float* v_template;   // [vp_dim]
    const float* shapedirs;    // [10][vp_dim]
    const float* j0;           // [66]
    const float* jdirs;        // [10][66]
    const float* skin_weight;  // [n_verts][n_skin]
    const int* skin_joint;     // [n_verts][n_skin]
    const int* jt_ptr;         // [23]   joint -> range in jt_vert / jt_weight
    const int* jt_vert;
    const float* jt_weight;
    int n_vj;                  // "virtual joints": the per-joint lists cut into chunks of bounded length
    const int* vj_ptr;         // [n_vj+1] chunk -> range in jt_vert / jt_weight
    const int* jvj_ptr;        // [23]     joint -> range of chunks
    const int* vinc_ptr;       // [n_verts+1] vertex -> its (item, code) incidences in the sensor phase
    const int* vinc_item;      // sensor * max_degree + face slot
    const int* vinc_code;      // 0/1/2: corner of that face, 3: the sensor vertex, 4: the helper vertex
    const int* parents;        // [22]
    const int* faces;          // [n_faces][3] local vertex ids
    const int* sensor_vert;    // [12]
    const int* helper_vert;    // [12]
    const int* sensor_faces;   // [12][max_degree] rows of faces, -1 padded
    const int* sensor_degree;  // [12]
};

// What the residual compares against and how it is weighted.
struct ResidualSpec {
    int use_pos, use_ori;      // reference flags use_marker_pos / use_marker_ori
    int sensor_active[kSensors];  // 1 if the sensor is in marker_idxs (models.py:386)
    float weight;              // scale of the sensor residual's gradient (1 for the LGD feature; training uses
                               // r_weight / (N+1) on the final iterate, models.py:672-674)
};

template <typename T, int VP = kMaxVp>
struct alignas(16) FrameState {
    // (the two vertex-sized vectors and beta come first so that the CUDA kernel can use 16-byte accesses on them)
    T vp[VP];                 // v_template + S beta + pose blend
    T dx[VP];                 // dE/dx, then reused for dE/dvp
    T beta[kBetas + 2];       // (two unused slots keep what follows 8-byte aligned for T = float)
    T theta[kPoseDim];
    T rot[kJoints][9];        // R_j = exp(theta_j)
    T jrest[kJoints][3];      // J(beta)
    T grot[kJoints][9];       // world rotation G_j^R  (== A_j^R)
    union {
        struct {
            T dar[kJoints][9];    // dE/dA^R
            T dat[kJoints][3];    // dE/dA^t
        };
        T fn[kSensors][kSplitDegree][3];   // split sensor phases: un-normalised face normals, then (slot 0) dE/dn
    };
    T dbeta_part[3][kBetas];
    union {
        T fg[kSensors * kSplitDegree][6];   // split sensor phases: dE/d(edge1), dE/d(edge2) of every (sensor, face) item
        T jup[kJoints][3];    // upstream dE/d(posed joint) of the FK loss (training, models.py:657-660); written by
                              // item_joint_residual, i.e. after item_sensor_gather has consumed `fg`
    };
    // Forward-only scratch and reverse-only scratch share storage: everything in `fwd` is dead once the sensor
    // outputs and joints have been written out, which is before the first member of `bwd` is written.
    union {
        struct {
            T gpos[kJoints][3];          // world position G_j^t  (posed joint)
            T atr[kJoints][3];           // A_j^t = G_j^t - G_j^R J_j
            T x[VP];                     // skinned vertices
            T sensor_pos[kSensors][3];   // p'_m (offsets applied)
            T sensor_ori[kSensors][9];   // R'_m row-major
        };
        union {
            struct {
                T dgr[kJoints][9];        // dE/dG^R
                T dgt[kJoints][3];        // dE/dG^t
                T drot[kJoints][9];       // dE/dR_j (chain part; the pose-blend part is added in phase_finish_theta or by the caller)
                T dj[kJoints][3];         // dE/dJ_j
            };
            T dav[kMaxVj][12];            // per-chunk partial sums of dE/dA (dead before dgr .. dj are written)
        };
    };
};

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
EMPOSE_HD void sincos_t(float a, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    sincosf(a, s, c);
#else
    *s = sinf(a); *c = cosf(a);
#endif
}
EMPOSE_HD void sincos_t(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
EMPOSE_HD float sqrt_t(float a) { return sqrtf(a); }
EMPOSE_HD double sqrt_t(double a) { return sqrt(a); }

template <typename T> EMPOSE_HD void cross3(const T* a, const T* b, T* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
template <typename T> EMPOSE_HD T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// y = x / |x|; returns |x|
template <typename T> EMPOSE_HD T normalize3(const T* x, T* y) {
    T len = sqrt_t(dot3(x, x));
    T inv = T(1) / len;
    y[0] = x[0] * inv; y[1] = x[1] * inv; y[2] = x[2] * inv;
    return len;
}
// reverse of y = x/|x| given unit y, len = |x| and dy:  dx = (dy - y (y.dy)) / len
template <typename T> EMPOSE_HD void normalize3_bwd(const T* y, T len, const T* dy, T* dx) {
    T proj = dot3(y, dy);
    T inv = T(1) / len;
    dx[0] = (dy[0] - y[0] * proj) * inv;
    dx[1] = (dy[1] - y[1] * proj) * inv;
    dx[2] = (dy[2] - y[2] * proj) * inv;
}

// Rodrigues with the third-party convention angle = ||r + 1e-8|| (axis = r / angle), row-major R.
template <typename T> EMPOSE_HD void rodrigues_fwd(const T* r, T* R) {
    const T e = T(1e-8);
    T r0 = r[0] + e, r1 = r[1] + e, r2 = r[2] + e;
    T a = sqrt_t(r0 * r0 + r1 * r1 + r2 * r2);
    T inv = T(1) / a;
    T kx = r[0] * inv, ky = r[1] * inv, kz = r[2] * inv;
    T s, c;
    sincos_t(a, &s, &c);
    T v = T(1) - c;
    // K = [[0,-kz,ky],[kz,0,-kx],[-ky,kx,0]];  K^2 = k k^T - |k|^2 I
    T kk = kx * kx + ky * ky + kz * kz;
    R[0] = T(1) + v * (kx * kx - kk);  R[1] = -s * kz + v * kx * ky;       R[2] = s * ky + v * kx * kz;
    R[3] = s * kz + v * kx * ky;       R[4] = T(1) + v * (ky * ky - kk);  R[5] = -s * kx + v * ky * kz;
    R[6] = -s * ky + v * kx * kz;      R[7] = s * kx + v * ky * kz;       R[8] = T(1) + v * (kz * kz - kk);
}

// dr += J^T dR for the map above.
template <typename T> EMPOSE_HD void rodrigues_bwd(const T* r, const T* dR, T* dr) {
    const T e = T(1e-8);
    T rp[3] = {r[0] + e, r[1] + e, r[2] + e};
    T a = sqrt_t(dot3(rp, rp));
    T inv = T(1) / a;
    T k[3] = {r[0] * inv, r[1] * inv, r[2] * inv};
    T s, c;
    sincos_t(a, &s, &c);
    T v = T(1) - c;
    T K[9] = {T(0), -k[2], k[1], k[2], T(0), -k[0], -k[1], k[0], T(0)};
    // K2 = K K
    T K2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
    T dot_k = T(0), dot_k2 = T(0);
    for (int i = 0; i < 9; ++i) { dot_k += dR[i] * K[i]; dot_k2 += dR[i] * K2[i]; }
    T da = c * dot_k + s * dot_k2;
    // dK = s dR + v (dR K^T + K^T dR)
    T dK[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T t1 = dR[i * 3] * K[j * 3] + dR[i * 3 + 1] * K[j * 3 + 1] + dR[i * 3 + 2] * K[j * 3 + 2];   // (dR K^T)_ij
            T t2 = K[i] * dR[j] + K[3 + i] * dR[3 + j] + K[6 + i] * dR[6 + j];                            // (K^T dR)_ij
            dK[i * 3 + j] = s * dR[i * 3 + j] + v * (t1 + t2);
        }
    T dk[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
    T kr = dk[0] * r[0] + dk[1] * r[1] + dk[2] * r[2];
    T inv3 = inv * inv * inv;
    for (int l = 0; l < 3; ++l) dr[l] += dk[l] * inv - kr * rp[l] * inv3 + da * rp[l] * inv;
}

// ----------------------------------------------------------------------------------------------
// forward phases
// ----------------------------------------------------------------------------------------------

// Every phase below is ONE loop `for (i = lane; i < n_items; i += lanes)` over independent items, so a caller
// may also run a single item i with (lane = i, lanes = huge); the item counts are given by the *_items helpers.

// F1a: joint rotations (22 items).
template <typename T, int VP>
EMPOSE_HD void item_rodrigues(FrameState<T, VP>& st, int j) {
    rodrigues_fwd(&st.theta[j * 3], st.rot[j]);
}
template <typename T, int VP>
EMPOSE_HD void phase_rodrigues(FrameState<T, VP>& st, int lane, int lanes) {
    for (int j = lane; j < kJoints; j += lanes) item_rodrigues(st, j);
}
// F1b: rest joints J(beta) (66 items).
template <typename T, int VP>
EMPOSE_HD void item_rest_joints(const SubModel& m, FrameState<T, VP>& st, int i) {
    T acc = T(m.j0[i]);
    for (int k = 0; k < kBetas; ++k) acc += T(m.jdirs[k * kPoseDim + i]) * st.beta[k];
    st.jrest[i / 3][i % 3] = acc;
}
template <typename T, int VP>
EMPOSE_HD void phase_rest_joints(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int i = lane; i < kPoseDim; i += lanes) item_rest_joints(m, st, i);
}
// F1c: blended rest vertices (3 * n_verts items).  vp_off may be null (treated as zero).
template <typename T, int VP, typename TIn>
EMPOSE_HD void item_blend_verts(const SubModel& m, FrameState<T, VP>& st, const TIn* vp_off, int i) {
    T acc = T(m.v_template[i]);
    for (int k = 0; k < kBetas; ++k) acc += T(m.shapedirs[k * m.vp_dim + i]) * st.beta[k];
    if (vp_off) acc += T(vp_off[i]);
    st.vp[i] = acc;
}
template <typename T, int VP, typename TIn>
EMPOSE_HD void phase_blend_verts(const SubModel& m, FrameState<T, VP>& st, const TIn* vp_off, int lane, int lanes) {
    for (int i = lane; i < m.n_verts * 3; i += lanes) item_blend_verts(m, st, vp_off, i);
}

// F2: kinematic chain.  Row r of every world rotation depends only on row r of its ancestors, so
// three lanes walk the whole tree independently (no synchronisation inside the chain).
template <typename T, int VP>
EMPOSE_HD void item_chain(const SubModel& m, FrameState<T, VP>& st, int r) {
    for (int c = 0; c < 3; ++c) st.grot[0][r * 3 + c] = st.rot[0][r * 3 + c];
    st.gpos[0][r] = st.jrest[0][r];
    for (int j = 1; j < kJoints; ++j) {
        const int p = m.parents[j];
        const T g0 = st.grot[p][r * 3], g1 = st.grot[p][r * 3 + 1], g2 = st.grot[p][r * 3 + 2];
        const T* R = st.rot[j];
        st.grot[j][r * 3 + 0] = g0 * R[0] + g1 * R[3] + g2 * R[6];
        st.grot[j][r * 3 + 1] = g0 * R[1] + g1 * R[4] + g2 * R[7];
        st.grot[j][r * 3 + 2] = g0 * R[2] + g1 * R[5] + g2 * R[8];
        st.gpos[j][r] = g0 * (st.jrest[j][0] - st.jrest[p][0]) + g1 * (st.jrest[j][1] - st.jrest[p][1]) +
                        g2 * (st.jrest[j][2] - st.jrest[p][2]) + st.gpos[p][r];
    }
    for (int j = 0; j < kJoints; ++j)
        st.atr[j][r] = st.gpos[j][r] - (st.grot[j][r * 3] * st.jrest[j][0] + st.grot[j][r * 3 + 1] * st.jrest[j][1] +
                                        st.grot[j][r * 3 + 2] * st.jrest[j][2]);
}
template <typename T, int VP>
EMPOSE_HD void phase_chain(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int r = lane; r < 3; r += lanes) item_chain(m, st, r);
}

// The SMPL body tree (reference configuration.py:118) as a compile-time function, so the chain can be fully
// unrolled with every world transform held in registers (no shared-memory round trip between joints).
EMPOSE_HD constexpr int smpl_parent(int j) {
    return j == 0 ? -1 : j <= 3 ? 0 : j <= 11 ? j - 3 : j <= 14 ? 9 : j <= 17 ? j - 3 : j - 2;
}

// F2': phase_chain specialised for the standard SMPL tree (same arithmetic, same order of operations).  Results
// are stored as soon as they exist so that only the transforms of pending parents stay live in registers.
template <typename T, int VP>
EMPOSE_HD void item_chain_static(FrameState<T, VP>& st, int r) {
    T g[kJoints][3], t[kJoints];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < kJoints; ++j) {
        if (j == 0) {
            g[0][0] = st.rot[0][r * 3]; g[0][1] = st.rot[0][r * 3 + 1]; g[0][2] = st.rot[0][r * 3 + 2];
            t[0] = st.jrest[0][r];
        } else {
            const int p = smpl_parent(j);
            const T* R = st.rot[j];
            g[j][0] = g[p][0] * R[0] + g[p][1] * R[3] + g[p][2] * R[6];
            g[j][1] = g[p][0] * R[1] + g[p][1] * R[4] + g[p][2] * R[7];
            g[j][2] = g[p][0] * R[2] + g[p][1] * R[5] + g[p][2] * R[8];
            t[j] = g[p][0] * (st.jrest[j][0] - st.jrest[p][0]) + g[p][1] * (st.jrest[j][1] - st.jrest[p][1]) +
                   g[p][2] * (st.jrest[j][2] - st.jrest[p][2]) + t[p];
        }
        st.grot[j][r * 3] = g[j][0]; st.grot[j][r * 3 + 1] = g[j][1]; st.grot[j][r * 3 + 2] = g[j][2];
        st.gpos[j][r] = t[j];
        st.atr[j][r] = t[j] - (g[j][0] * st.jrest[j][0] + g[j][1] * st.jrest[j][1] + g[j][2] * st.jrest[j][2]);
    }
}
template <typename T, int VP>
EMPOSE_HD void phase_chain_static(FrameState<T, VP>& st, int lane, int lanes) {
    for (int r = lane; r < 3; r += lanes) item_chain_static(st, r);
}

// F3: linear blend skinning of the sub-mesh; also clears dx for the reverse pass.
template <typename T, int VP>
EMPOSE_HD void item_skin(const SubModel& m, FrameState<T, VP>& st, int v) {
    const T p0 = st.vp[v * 3], p1 = st.vp[v * 3 + 1], p2 = st.vp[v * 3 + 2];
    T x0 = T(0), x1 = T(0), x2 = T(0);
    for (int s = 0; s < m.n_skin; ++s) {
        const T w = T(m.skin_weight[v * m.n_skin + s]);
        const int j = m.skin_joint[v * m.n_skin + s];
        const T* A = st.grot[j];
        x0 += w * (A[0] * p0 + A[1] * p1 + A[2] * p2 + st.atr[j][0]);
        x1 += w * (A[3] * p0 + A[4] * p1 + A[5] * p2 + st.atr[j][1]);
        x2 += w * (A[6] * p0 + A[7] * p1 + A[8] * p2 + st.atr[j][2]);
    }
    st.x[v * 3] = x0; st.x[v * 3 + 1] = x1; st.x[v * 3 + 2] = x2;
    st.dx[v * 3] = T(0); st.dx[v * 3 + 1] = T(0); st.dx[v * 3 + 2] = T(0);
}
template <typename T, int VP>
EMPOSE_HD void phase_skin(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int v = lane; v < m.n_verts; v += lanes) item_skin(m, st, v);
}

#if defined(__CUDA_ARCH__)
template <typename T> __device__ __forceinline__ void scatter_add(T* addr, T val) { atomicAdd(addr, val); }
#else
template <typename T> inline void scatter_add(T* addr, T val) { *addr += val; }
#endif

// F4 (+ its reverse): one lane per sensor builds the frame [on_surface | third | normal]
// (virtual_sensors.py:16-38), applies the offsets (models.py:478-479) and, if `want_grad`, forms the
// residual direction (loss.py:27-28) and pushes it back to the vertices it touched (st.dx).
//   meas_pos: [12][3] measured positions, meas_ori: [12][9] measured orientations (row-major),
//   off_r: [12][9], off_t: [12][3].
template <typename T, int VP, typename TIn>
EMPOSE_HD void item_sensors(const SubModel& m, FrameState<T, VP>& st, const TIn* off_r, const TIn* off_t,
                             const TIn* meas_pos, const TIn* meas_ori, const ResidualSpec& spec, bool want_grad,
                             int s) {
    const int vs = m.sensor_vert[s], vh = m.helper_vert[s];
    const int deg = m.sensor_degree[s];
    const T* xs = &st.x[vs * 3];
    // area-weighted normal: mean of un-normalised incident face normals (utils.py:134-140)
    T n[3] = {T(0), T(0), T(0)};
    for (int d = 0; d < deg; ++d) {
        const int* f = &m.faces[m.sensor_faces[s * m.max_degree + d] * 3];
        const T* a = &st.x[f[0] * 3];
        const T* b = &st.x[f[1] * 3];
        const T* c = &st.x[f[2] * 3];
        T e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
        T e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        T fn[3];
        cross3(e1, e2, fn);
        n[0] += fn[0]; n[1] += fn[1]; n[2] += fn[2];
    }
    const T inv_deg = T(1) / T(deg);
    n[0] *= inv_deg; n[1] *= inv_deg; n[2] *= inv_deg;
    T nh[3], u[3], s0[3], t[3], th[3], sv[3], sh[3];
    const T n_len = normalize3(n, nh);
    u[0] = st.x[vh * 3] - xs[0]; u[1] = st.x[vh * 3 + 1] - xs[1]; u[2] = st.x[vh * 3 + 2] - xs[2];
    const T u_len = normalize3(u, s0);
    cross3(nh, s0, t);
    const T t_len = normalize3(t, th);
    cross3(th, nh, sv);
    const T s_len = normalize3(sv, sh);
    // R = [sh | th | nh] as columns, row-major storage
    T R[9] = {sh[0], th[0], nh[0], sh[1], th[1], nh[1], sh[2], th[2], nh[2]};
    const TIn* Ro = off_r + s * 9;
    const TIn* to = off_t + s * 3;
    T Rc[9], pc[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            Rc[i * 3 + j] = R[i * 3] * T(Ro[j]) + R[i * 3 + 1] * T(Ro[3 + j]) + R[i * 3 + 2] * T(Ro[6 + j]);
        pc[i] = xs[i] + R[i * 3] * T(to[0]) + R[i * 3 + 1] * T(to[1]) + R[i * 3 + 2] * T(to[2]);
    }
    for (int i = 0; i < 9; ++i) st.sensor_ori[s][i] = Rc[i];
    for (int i = 0; i < 3; ++i) st.sensor_pos[s][i] = pc[i];
    if (!want_grad || !spec.sensor_active[s]) return;

    // ---- reverse ----
    T dpc[3] = {T(0), T(0), T(0)}, dRc[9];
    for (int i = 0; i < 9; ++i) dRc[i] = T(0);
    if (spec.use_pos) {
        T d[3] = {pc[0] - T(meas_pos[s * 3]), pc[1] - T(meas_pos[s * 3 + 1]), pc[2] - T(meas_pos[s * 3 + 2])};
        T len = sqrt_t(dot3(d, d));
        T inv = len > T(0) ? T(spec.weight) / len : T(0);   // reference: NaN at exactly zero residual (sqrt backward); we emit 0
        dpc[0] = d[0] * inv; dpc[1] = d[1] * inv; dpc[2] = d[2] * inv;
    }
    if (spec.use_ori) {
        T d[9], sq = T(0);
        for (int i = 0; i < 9; ++i) { d[i] = Rc[i] - T(meas_ori[s * 9 + i]); sq += d[i] * d[i]; }
        T len = sqrt_t(sq);
        T inv = len > T(0) ? T(spec.weight) / len : T(0);
        for (int i = 0; i < 9; ++i) dRc[i] = d[i] * inv;
    }
    // offsets: Rc = R Ro, pc = xs + R to
    T dR[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            dR[i * 3 + j] = dRc[i * 3] * T(Ro[j * 3]) + dRc[i * 3 + 1] * T(Ro[j * 3 + 1]) + dRc[i * 3 + 2] * T(Ro[j * 3 + 2]) +
                            dpc[i] * T(to[j]);
    T dsh[3] = {dR[0], dR[3], dR[6]}, dth[3] = {dR[1], dR[4], dR[7]}, dnh[3] = {dR[2], dR[5], dR[8]};
    // sh = sv/|sv|, sv = th x nh
    T dsv[3], tmp[3];
    normalize3_bwd(sh, s_len, dsh, dsv);
    cross3(nh, dsv, tmp); dth[0] += tmp[0]; dth[1] += tmp[1]; dth[2] += tmp[2];      // d th += nh x dsv
    cross3(dsv, th, tmp); dnh[0] += tmp[0]; dnh[1] += tmp[1]; dnh[2] += tmp[2];      // d nh += dsv x th
    // th = t/|t|, t = nh x s0
    T dt[3], ds0[3];
    normalize3_bwd(th, t_len, dth, dt);
    cross3(s0, dt, tmp); dnh[0] += tmp[0]; dnh[1] += tmp[1]; dnh[2] += tmp[2];       // d nh += s0 x dt
    cross3(dt, nh, ds0);                                                             // d s0  = dt x nh
    // s0 = u/|u|
    T du[3];
    normalize3_bwd(s0, u_len, ds0, du);
    // nh = n/|n|, n = mean of face normals
    T dn[3];
    normalize3_bwd(nh, n_len, dnh, dn);
    dn[0] *= inv_deg; dn[1] *= inv_deg; dn[2] *= inv_deg;
    for (int i = 0; i < 3; ++i) {
        scatter_add(&st.dx[vh * 3 + i], du[i]);
        scatter_add(&st.dx[vs * 3 + i], dpc[i] - du[i]);
    }
    for (int d = 0; d < deg; ++d) {
        const int* f = &m.faces[m.sensor_faces[s * m.max_degree + d] * 3];
        const T* a = &st.x[f[0] * 3];
        const T* b = &st.x[f[1] * 3];
        const T* c = &st.x[f[2] * 3];
        T e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
        T e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        T de1[3], de2[3];
        cross3(e2, dn, de1);      // fn = e1 x e2: de1 = e2 x dfn, de2 = dfn x e1
        cross3(dn, e1, de2);
        for (int i = 0; i < 3; ++i) {
            scatter_add(&st.dx[f[1] * 3 + i], de1[i]);
            scatter_add(&st.dx[f[2] * 3 + i], de2[i]);
            scatter_add(&st.dx[f[0] * 3 + i], -(de1[i] + de2[i]));
        }
    }
}
template <typename T, int VP, typename TIn>
EMPOSE_HD void phase_sensors(const SubModel& m, FrameState<T, VP>& st, const TIn* off_r, const TIn* off_t,
                             const TIn* meas_pos, const TIn* meas_ori, const ResidualSpec& spec, bool want_grad,
                             int lane, int lanes) {
    for (int s = lane; s < kSensors; s += lanes) item_sensors(m, st, off_r, off_t, meas_pos, meas_ori, spec, want_grad, s);
}

// F4 split in three so that the face work runs over (sensor, face) items instead of inside 12 long serial lanes
// (used when max_degree <= kSplitDegree; same arithmetic as phase_sensors).
// F4a: un-normalised face normals (12 * max_degree items).
template <typename T, int VP>
EMPOSE_HD void item_sensor_faces(const SubModel& m, FrameState<T, VP>& st, int it) {
    const int s = it / m.max_degree, d = it % m.max_degree;
    if (d >= m.sensor_degree[s]) return;
    const int* f = &m.faces[m.sensor_faces[s * m.max_degree + d] * 3];
    const T* a = &st.x[f[0] * 3];
    const T* b = &st.x[f[1] * 3];
    const T* c = &st.x[f[2] * 3];
    T e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    T e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    cross3(e1, e2, st.fn[s][d]);
}
template <typename T, int VP>
EMPOSE_HD void phase_sensor_faces(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int it = lane; it < kSensors * m.max_degree; it += lanes) item_sensor_faces(m, st, it);
}

// F4b: sensor frame, offsets, residual and its reverse up to dE/dn (12 items).  Leaves in st.fn[s][0..2]: dE/dn
// (already divided by the degree), the helper-vertex term and the sensor-vertex term (zero for unused sensors).
template <typename T, int VP, typename TIn>
EMPOSE_HD void item_sensor_frames(const SubModel& m, FrameState<T, VP>& st, const TIn* off_r, const TIn* off_t,
                                   const TIn* meas_pos, const TIn* meas_ori, const ResidualSpec& spec, bool want_grad,
                                   int s) {
    const int vs = m.sensor_vert[s], vh = m.helper_vert[s];
    const int deg = m.sensor_degree[s];
    const T* xs = &st.x[vs * 3];
    T n[3] = {T(0), T(0), T(0)};
    for (int d = 0; d < deg; ++d) { n[0] += st.fn[s][d][0]; n[1] += st.fn[s][d][1]; n[2] += st.fn[s][d][2]; }
    const T inv_deg = T(1) / T(deg);
    n[0] *= inv_deg; n[1] *= inv_deg; n[2] *= inv_deg;
    T nh[3], u[3], s0[3], t[3], th[3], sv[3], sh[3];
    const T n_len = normalize3(n, nh);
    u[0] = st.x[vh * 3] - xs[0]; u[1] = st.x[vh * 3 + 1] - xs[1]; u[2] = st.x[vh * 3 + 2] - xs[2];
    const T u_len = normalize3(u, s0);
    cross3(nh, s0, t);
    const T t_len = normalize3(t, th);
    cross3(th, nh, sv);
    const T s_len = normalize3(sv, sh);
    T R[9] = {sh[0], th[0], nh[0], sh[1], th[1], nh[1], sh[2], th[2], nh[2]};
    const TIn* Ro = off_r + s * 9;
    const TIn* to = off_t + s * 3;
    T Rc[9], pc[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            Rc[i * 3 + j] = R[i * 3] * T(Ro[j]) + R[i * 3 + 1] * T(Ro[3 + j]) + R[i * 3 + 2] * T(Ro[6 + j]);
        pc[i] = xs[i] + R[i * 3] * T(to[0]) + R[i * 3 + 1] * T(to[1]) + R[i * 3 + 2] * T(to[2]);
    }
    for (int i = 0; i < 9; ++i) st.sensor_ori[s][i] = Rc[i];
    for (int i = 0; i < 3; ++i) st.sensor_pos[s][i] = pc[i];
    if (!want_grad) return;
    if (!spec.sensor_active[s]) {
        for (int q = 0; q < 3; ++q) { st.fn[s][q][0] = T(0); st.fn[s][q][1] = T(0); st.fn[s][q][2] = T(0); }
        return;
    }
    T dpc[3] = {T(0), T(0), T(0)}, dRc[9];
    for (int i = 0; i < 9; ++i) dRc[i] = T(0);
    if (spec.use_pos) {
        T d[3] = {pc[0] - T(meas_pos[s * 3]), pc[1] - T(meas_pos[s * 3 + 1]), pc[2] - T(meas_pos[s * 3 + 2])};
        T len = sqrt_t(dot3(d, d));
        T inv = len > T(0) ? T(spec.weight) / len : T(0);
        dpc[0] = d[0] * inv; dpc[1] = d[1] * inv; dpc[2] = d[2] * inv;
    }
    if (spec.use_ori) {
        T d[9], sq = T(0);
        for (int i = 0; i < 9; ++i) { d[i] = Rc[i] - T(meas_ori[s * 9 + i]); sq += d[i] * d[i]; }
        T len = sqrt_t(sq);
        T inv = len > T(0) ? T(spec.weight) / len : T(0);
        for (int i = 0; i < 9; ++i) dRc[i] = d[i] * inv;
    }
    T dR[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            dR[i * 3 + j] = dRc[i * 3] * T(Ro[j * 3]) + dRc[i * 3 + 1] * T(Ro[j * 3 + 1]) + dRc[i * 3 + 2] * T(Ro[j * 3 + 2]) +
                            dpc[i] * T(to[j]);
    T dsh[3] = {dR[0], dR[3], dR[6]}, dth[3] = {dR[1], dR[4], dR[7]}, dnh[3] = {dR[2], dR[5], dR[8]};
    T dsv[3], tmp[3];
    normalize3_bwd(sh, s_len, dsh, dsv);
    cross3(nh, dsv, tmp); dth[0] += tmp[0]; dth[1] += tmp[1]; dth[2] += tmp[2];
    cross3(dsv, th, tmp); dnh[0] += tmp[0]; dnh[1] += tmp[1]; dnh[2] += tmp[2];
    T dt[3], ds0[3];
    normalize3_bwd(th, t_len, dth, dt);
    cross3(s0, dt, tmp); dnh[0] += tmp[0]; dnh[1] += tmp[1]; dnh[2] += tmp[2];
    cross3(dt, nh, ds0);
    T du[3];
    normalize3_bwd(s0, u_len, ds0, du);
    T dn[3];
    normalize3_bwd(nh, n_len, dnh, dn);
    for (int i = 0; i < 3; ++i) {
        st.fn[s][0][i] = dn[i] * inv_deg;
        st.fn[s][1][i] = du[i];                // goes to the helper vertex
        st.fn[s][2][i] = dpc[i] - du[i];       // goes to the sensor vertex
    }
}
template <typename T, int VP, typename TIn>
EMPOSE_HD void phase_sensor_frames(const SubModel& m, FrameState<T, VP>& st, const TIn* off_r, const TIn* off_t,
                                   const TIn* meas_pos, const TIn* meas_ori, const ResidualSpec& spec, bool want_grad,
                                   int lane, int lanes) {
    for (int s = lane; s < kSensors; s += lanes) item_sensor_frames(m, st, off_r, off_t, meas_pos, meas_ori, spec, want_grad, s);
}

// F4c: reverse of the face normals, once per (sensor, face) item (12 * max_degree items): with fn = e1 x e2,
// dE/de1 = e2 x dE/dfn and dE/de2 = dE/dfn x e1, where dE/dfn is the sensor's dE/dn.
template <typename T, int VP>
EMPOSE_HD void item_sensor_face_grads(const SubModel& m, FrameState<T, VP>& st, int it) {
    const int s = it / m.max_degree, d = it % m.max_degree;
    if (d >= m.sensor_degree[s]) return;
    const T* dn = st.fn[s][0];
    const int* f = &m.faces[m.sensor_faces[it] * 3];
    const T* a = &st.x[f[0] * 3];
    const T* b = &st.x[f[1] * 3];
    const T* c = &st.x[f[2] * 3];
    T e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    T e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    cross3(e2, dn, &st.fg[it][0]);
    cross3(dn, e1, &st.fg[it][3]);
}
template <typename T, int VP>
EMPOSE_HD void phase_sensor_face_grads(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int it = lane; it < kSensors * m.max_degree; it += lanes) item_sensor_face_grads(m, st, it);
}

// F4d: dE/dx by GATHER (n_verts items): every vertex sums, in a fixed order, the contributions of the faces it
// is a corner of and of the sensors it serves as sensor / helper vertex.  No atomics: the result is
// bit-reproducible, which window-sharded inference relies on.
template <typename T, int VP>
EMPOSE_HD void item_sensor_gather(const SubModel& m, FrameState<T, VP>& st, int v) {
    T g[3] = {T(0), T(0), T(0)};
    for (int q = m.vinc_ptr[v]; q < m.vinc_ptr[v + 1]; ++q) {
        const int item = m.vinc_item[q], code = m.vinc_code[q];
        if (code >= 3) {
            const T* t = st.fn[item / m.max_degree][code == 3 ? 2 : 1];
            g[0] += t[0]; g[1] += t[1]; g[2] += t[2];
        } else {
            const T* e = st.fg[item];
            if (code == 1) { g[0] += e[0]; g[1] += e[1]; g[2] += e[2]; }
            else if (code == 2) { g[0] += e[3]; g[1] += e[4]; g[2] += e[5]; }
            else { g[0] -= e[0] + e[3]; g[1] -= e[1] + e[4]; g[2] -= e[2] + e[5]; }
        }
    }
    st.dx[v * 3] = g[0]; st.dx[v * 3 + 1] = g[1]; st.dx[v * 3 + 2] = g[2];
}
template <typename T, int VP>
EMPOSE_HD void phase_sensor_gather(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int v = lane; v < m.n_verts; v += lanes) item_sensor_gather(m, st, v);
}

// ----------------------------------------------------------------------------------------------
// reverse phases
// ----------------------------------------------------------------------------------------------

// B0 (training only): upstream gradient of the FK loss sum_j ||J_j - Jgt_j|| (loss.py:27-28 applied to joints,
// models.py:657-660) times `weight`, kept in st.jup (22 items).  Must run while st.gpos is still alive.
template <typename T, int VP, typename TIn>
EMPOSE_HD void item_joint_residual(FrameState<T, VP>& st, const TIn* joints_gt, T weight, int j) {
    T d[3] = {st.gpos[j][0] - T(joints_gt[j * 3]), st.gpos[j][1] - T(joints_gt[j * 3 + 1]), st.gpos[j][2] - T(joints_gt[j * 3 + 2])};
    T len = sqrt_t(dot3(d, d));
    T inv = len > T(0) ? weight / len : T(0);
    st.jup[j][0] = d[0] * inv; st.jup[j][1] = d[1] * inv; st.jup[j][2] = d[2] * inv;
}
template <typename T, int VP, typename TIn>
EMPOSE_HD void phase_joint_residual(FrameState<T, VP>& st, const TIn* joints_gt, T weight, int lane, int lanes) {
    for (int j = lane; j < kJoints; j += lanes) item_joint_residual(st, joints_gt, weight, j);
}

// B1a: dE/dA summed over one chunk ("virtual joint") of a joint's vertex list (n_vj items).  Each chunk walks
// its (bounded) list once and accumulates all 12 entries in registers, so lanes stay balanced.
template <typename T, int VP>
EMPOSE_HD void item_skin_bwd_chunks(const SubModel& m, FrameState<T, VP>& st, int c) {
    T acc[12];
    for (int e = 0; e < 12; ++e) acc[e] = T(0);
    for (int q = m.vj_ptr[c]; q < m.vj_ptr[c + 1]; ++q) {
        const int v = m.jt_vert[q];
        const T w = T(m.jt_weight[q]);
        const T d0 = w * st.dx[v * 3], d1 = w * st.dx[v * 3 + 1], d2 = w * st.dx[v * 3 + 2];
        const T p0 = st.vp[v * 3], p1 = st.vp[v * 3 + 1], p2 = st.vp[v * 3 + 2];
        acc[0] += d0 * p0; acc[1] += d0 * p1; acc[2] += d0 * p2;
        acc[3] += d1 * p0; acc[4] += d1 * p1; acc[5] += d1 * p2;
        acc[6] += d2 * p0; acc[7] += d2 * p1; acc[8] += d2 * p2;
        acc[9] += d0; acc[10] += d1; acc[11] += d2;
    }
    for (int e = 0; e < 12; ++e) st.dav[c][e] = acc[e];
}
template <typename T, int VP>
EMPOSE_HD void phase_skin_bwd_chunks(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int c = lane; c < m.n_vj; c += lanes) item_skin_bwd_chunks(m, st, c);
}
// B1b: dE/dA_j = sum of its chunks (22 * 12 items).
template <typename T, int VP>
EMPOSE_HD void item_skin_bwd_reduce(const SubModel& m, FrameState<T, VP>& st, int it) {
    const int j = it / 12, e = it % 12;
    T acc = T(0);
    for (int c = m.jvj_ptr[j]; c < m.jvj_ptr[j + 1]; ++c) acc += st.dav[c][e];
    if (e < 9) st.dar[j][e] = acc;
    else st.dat[j][e - 9] = acc;
}
template <typename T, int VP>
EMPOSE_HD void phase_skin_bwd_reduce(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int it = lane; it < kJoints * 12; it += lanes) item_skin_bwd_reduce(m, st, it);
}

// B2: dE/dvp_v = sum_j w A_j^R^T dE/dx_v, in place over st.dx.  Must run AFTER phase_skin_bwd_chunks.
template <typename T, int VP>
EMPOSE_HD void item_skin_bwd_verts(const SubModel& m, FrameState<T, VP>& st, int v) {
    const T d0 = st.dx[v * 3], d1 = st.dx[v * 3 + 1], d2 = st.dx[v * 3 + 2];
    T g0 = T(0), g1 = T(0), g2 = T(0);
    for (int s = 0; s < m.n_skin; ++s) {
        const T w = T(m.skin_weight[v * m.n_skin + s]);
        const T* A = st.grot[m.skin_joint[v * m.n_skin + s]];
        g0 += w * (A[0] * d0 + A[3] * d1 + A[6] * d2);
        g1 += w * (A[1] * d0 + A[4] * d1 + A[7] * d2);
        g2 += w * (A[2] * d0 + A[5] * d1 + A[8] * d2);
    }
    st.dx[v * 3] = g0; st.dx[v * 3 + 1] = g1; st.dx[v * 3 + 2] = g2;
}
template <typename T, int VP>
EMPOSE_HD void phase_skin_bwd_verts(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int v = lane; v < m.n_verts; v += lanes) item_skin_bwd_verts(m, st, v);
}

// B3: partial sums of dE/dbeta through the shape blend shapes (st.dx now holds dE/dvp): 30 items,
// item (k, p) sums the vertices v = p (mod 3).
template <typename T, int VP>
EMPOSE_HD void item_shape_bwd_partial(const SubModel& m, FrameState<T, VP>& st, int it) {
    const int k = it % kBetas, p = it / kBetas;
    const float* S = m.shapedirs + k * m.vp_dim;
    T acc = T(0);
    for (int v = p; v < m.n_verts; v += 3)
        acc += T(S[v * 3]) * st.dx[v * 3] + T(S[v * 3 + 1]) * st.dx[v * 3 + 1] + T(S[v * 3 + 2]) * st.dx[v * 3 + 2];
    st.dbeta_part[p][k] = acc;
}
template <typename T, int VP>
EMPOSE_HD void phase_shape_bwd_partial(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes) {
    for (int it = lane; it < 3 * kBetas; it += lanes) item_shape_bwd_partial(m, st, it);
}

// B4: reverse sweep of the kinematic chain, row-parallel like the forward one: lane r owns row r of
// every dE/dG^R and entry r of every dE/dG^t.  After the sweep both are final for every joint.
template <typename T, int VP>
EMPOSE_HD void item_chain_bwd(const SubModel& m, FrameState<T, VP>& st, int r, bool joint_up = false) {
    for (int j = 0; j < kJoints; ++j) {
        // A_j^R = G_j^R,  A_j^t = G_j^t - G_j^R J_j;  the posed joint itself is G_j^t (FK loss upstream)
        const T a = st.dat[j][r];
        st.dgt[j][r] = joint_up ? a + st.jup[j][r] : a;
        for (int c = 0; c < 3; ++c) st.dgr[j][r * 3 + c] = st.dar[j][r * 3 + c] - a * st.jrest[j][c];
    }
    for (int j = kJoints - 1; j >= 1; --j) {
        const int p = m.parents[j];
        const T* R = st.rot[j];
        const T d0 = st.dgr[j][r * 3], d1 = st.dgr[j][r * 3 + 1], d2 = st.dgr[j][r * 3 + 2];
        const T dt = st.dgt[j][r];
        // G_j^R = G_p^R R_j           -> dG_p^R += dG_j^R R_j^T
        // G_j^t = G_p^R (J_j - J_p) + G_p^t -> dG_p^R += dG_j^t (J_j - J_p)^T,  dG_p^t += dG_j^t
        st.dgr[p][r * 3 + 0] += d0 * R[0] + d1 * R[1] + d2 * R[2] + dt * (st.jrest[j][0] - st.jrest[p][0]);
        st.dgr[p][r * 3 + 1] += d0 * R[3] + d1 * R[4] + d2 * R[5] + dt * (st.jrest[j][1] - st.jrest[p][1]);
        st.dgr[p][r * 3 + 2] += d0 * R[6] + d1 * R[7] + d2 * R[8] + dt * (st.jrest[j][2] - st.jrest[p][2]);
        st.dgt[p][r] += dt;
    }
}
template <typename T, int VP>
EMPOSE_HD void phase_chain_bwd(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes, bool joint_up = false) {
    for (int r = lane; r < 3; r += lanes) item_chain_bwd(m, st, r, joint_up);
}

// B4': phase_chain_bwd specialised for the standard SMPL tree.  Children contributions are accumulated in
// registers (`acc`, zero until first touched) and each joint is finalised and stored when the sweep reaches it,
// so only the partial sums of pending parents are live.
template <typename T, int VP>
EMPOSE_HD void item_chain_bwd_static(FrameState<T, VP>& st, int r, bool joint_up = false) {
    T acc[kJoints][3], acct[kJoints];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < kJoints; ++j) { acc[j][0] = T(0); acc[j][1] = T(0); acc[j][2] = T(0); acct[j] = T(0); }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = kJoints - 1; j >= 0; --j) {
        const T a = st.dat[j][r];
        const T dt = joint_up ? a + st.jup[j][r] + acct[j] : a + acct[j];
        const T d0 = st.dar[j][r * 3] - a * st.jrest[j][0] + acc[j][0];
        const T d1 = st.dar[j][r * 3 + 1] - a * st.jrest[j][1] + acc[j][1];
        const T d2 = st.dar[j][r * 3 + 2] - a * st.jrest[j][2] + acc[j][2];
        st.dgr[j][r * 3] = d0; st.dgr[j][r * 3 + 1] = d1; st.dgr[j][r * 3 + 2] = d2;
        st.dgt[j][r] = dt;
        if (j > 0) {
            const int p = smpl_parent(j);
            const T* R = st.rot[j];
            acc[p][0] += d0 * R[0] + d1 * R[1] + d2 * R[2] + dt * (st.jrest[j][0] - st.jrest[p][0]);
            acc[p][1] += d0 * R[3] + d1 * R[4] + d2 * R[5] + dt * (st.jrest[j][1] - st.jrest[p][1]);
            acc[p][2] += d0 * R[6] + d1 * R[7] + d2 * R[8] + dt * (st.jrest[j][2] - st.jrest[p][2]);
            acct[p] += dt;
        }
    }
}
template <typename T, int VP>
EMPOSE_HD void phase_chain_bwd_static(FrameState<T, VP>& st, int lane, int lanes, bool joint_up = false) {
    for (int r = lane; r < 3; r += lanes) item_chain_bwd_static(st, r, joint_up);
}

// B5: local gradients from the final dE/dG.
//   dE/dR_j = G_p^R^T dE/dG_j^R   (j > 0),   dE/dR_0 = dE/dG_0^R
//   dE/dJ_j = (G_p^R - G_j^R)^T dE/dG_j^t   (j > 0),   dE/dJ_0 = (I - G_0^R)^T dE/dG_0^t
// (the second line collects the three places J_j appears: A_j^t, its own bone and its children's bones,
//  using dE/dG_j^t = dE/dA_j^t + sum over children of dE/dG_child^t).
// With an upstream gradient u_j on the posed joint itself (FK loss), dE/dG_j^t also contains u_j, which does not
// belong to the A_j^t / children terms: dE/dJ_j += G_j^R^T u_j.
template <typename T, int VP>
EMPOSE_HD void item_chain_bwd_local(const SubModel& m, FrameState<T, VP>& st, int i, bool joint_up = false) {
    const int j = i / 12, e = i % 12;
    const int p = m.parents[j];
    if (e < 9) {
        const int a = e / 3, b = e % 3;
        T acc;
        if (j == 0) acc = st.dgr[0][e];
        else acc = st.grot[p][a] * st.dgr[j][b] + st.grot[p][3 + a] * st.dgr[j][3 + b] + st.grot[p][6 + a] * st.dgr[j][6 + b];
        st.drot[j][e] = acc;
    } else {
        const int c = e - 9;
        T acc = T(0);
        for (int r = 0; r < 3; ++r) {
            const T gp = (j == 0) ? (r == c ? T(1) : T(0)) : st.grot[p][r * 3 + c];
            acc += (gp - st.grot[j][r * 3 + c]) * st.dgt[j][r];
            if (joint_up) acc += st.grot[j][r * 3 + c] * st.jup[j][r];
        }
        st.dj[j][c] = acc;
    }
}
template <typename T, int VP>
EMPOSE_HD void phase_chain_bwd_local(const SubModel& m, FrameState<T, VP>& st, int lane, int lanes, bool joint_up = false) {
    for (int i = lane; i < kJoints * 12; i += lanes) item_chain_bwd_local(m, st, i, joint_up);
}

// B6: finish.  g_theta and the complete g_beta, both scaled by `coef`
// (= [f < len_b] * frame_mask * F / len_b, the per-frame form of models.py:560-579).  `dpf`
// (dE/d pose-feature, the result of the transposed pose-blend GEMM) may be null: the map is linear in
// dR, so a caller can add coef * rodrigues_bwd(theta_j, dpf_j) later (the split GPU kernels do that).
template <typename T, int VP, typename TOut, typename TPf>
EMPOSE_HD void item_finish_theta(FrameState<T, VP>& st, T coef, const TPf* dpf, TOut* g_theta, int j) {
    T g[3] = {T(0), T(0), T(0)};
    if (dpf && j > 0)
        for (int e = 0; e < 9; ++e) st.drot[j][e] += T(dpf[(j - 1) * 9 + e]);
    rodrigues_bwd(&st.theta[j * 3], st.drot[j], g);
    g_theta[j * 3] = TOut(coef * g[0]); g_theta[j * 3 + 1] = TOut(coef * g[1]); g_theta[j * 3 + 2] = TOut(coef * g[2]);
}
template <typename T, int VP, typename TOut, typename TPf>
EMPOSE_HD void phase_finish_theta(FrameState<T, VP>& st, T coef, const TPf* dpf, TOut* g_theta, int lane, int lanes) {
    for (int j = lane; j < kJoints; j += lanes) item_finish_theta(st, coef, dpf, g_theta, j);
}
template <typename T, int VP, typename TOut>
EMPOSE_HD void item_finish_beta(const SubModel& m, FrameState<T, VP>& st, T coef, TOut* g_beta, int k) {
    T acc = st.dbeta_part[0][k] + st.dbeta_part[1][k] + st.dbeta_part[2][k];
    for (int i = 0; i < kPoseDim; ++i) acc += T(m.jdirs[k * kPoseDim + i]) * st.dj[i / 3][i % 3];
    g_beta[k] = TOut(coef * acc);
}
template <typename T, int VP, typename TOut>
EMPOSE_HD void phase_finish_beta(const SubModel& m, FrameState<T, VP>& st, T coef, TOut* g_beta, int lane, int lanes) {
    for (int k = lane; k < kBetas; k += lanes) item_finish_beta(m, st, coef, g_beta, k);
}

}  // namespace empose
