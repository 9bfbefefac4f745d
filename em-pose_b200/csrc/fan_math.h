// The per-frame SMPL-H sub-model pass in "fan form": one work item per (frame, sensor), held in registers.
//
// Same arithmetic as frame_math.h -- SMPL-H linear blend skinning restricted to the 1-rings of the 12 sensor vertices
// (third-party BodyModel call at empose/bodymodels/smpl.py:121), the sensor frames of
// VirtualMarkerHelper.get_virtual_pos_and_rot (empose/data/virtual_sensors.py:16-38, 85-96; normals as in
// empose/helpers/utils.py:126-146), the offsets of empose/nn/models.py:478-479, reconstruction_loss
// (empose/nn/loss.py:23-41) and the hand-derived reverse pass that replaces autograd (models.py:576-579) -- but
// organised around what the hardware is good at:
//
//   * The sub-mesh is stored ring-major: sensor s owns the block [s * slots, (s+1) * slots) of sub-model vertices, slot 0
//     being the sensor vertex and slots 1..deg its neighbours in the winding order of the incident faces, so that face d
//     of the sensor is (slot 0, slot 1+d, slot 1+(d+1) % deg) -- a closed fan.  A face normal (v1-v0)x(v2-v0) is
//     invariant under cyclic rotation of its corners, so with e_d = x_{1+d} - x_0 the area-weighted vertex normal of
//     utils.py:134-140 is  (1/deg) sum_d e_d x e_{d+1}  and its reverse is  dE/de_d = (e_{d+1} - e_{d-1}) x dE/dN
//     (the sensor vertex itself drops out of a closed fan's normal).  No index tables, no scatter.
//   * One lane skins its 7..12 ring vertices, builds the sensor frame, forms the residual and runs the reverse pass down
//     to dE/dvp of its ring and its contributions to dE/dA_j -- all in registers.  The joints a ring is skinned to are
//     the union over its vertices (<= kMaxFanJoints, dense weights with zeros), so each joint transform is fetched once
//     per lane instead of once per (vertex, joint) pair.
//   * The serial parts (Rodrigues, the kinematic chain and its reverse sweep, the local gradients) work on a small
//     per-frame JointState; the lanes of a frame hand their dE/dA partial sums over through a list that is reduced in a
//     fixed order (bit-reproducible, no atomics).
//
// Shape blend and rest joints are NOT computed here: `vp` is the complete blended rest vertex (v_template + S beta +
// P pf) and `jrest` is J0 + Jdirs beta, both produced by the blend GEMM (kFeat* in frame_math.h); the reverse pass
// returns dE/dvp and dE/dJ, which the transposed GEMM contracts with P, S and Jdirs.
//
// Host + device, templated on the scalar type: tests/host_harness.cpp runs it in double against autograd.
#pragma once

#include "frame_math.h"

namespace empose {

constexpr int kMaxFanJoints = 8;                          // distinct skinning joints over one sensor's ring
constexpr int kMaxPartials = kSensors * kMaxFanJoints;    // (sensor, joint) partial sums of dE/dA per frame

// Fan tables of the sub-model (submodel.py: fan_tables); pointers into device or host memory.
struct FanModel {
    int ok;                 // 1: every sensor ring is a closed, consistently wound fan with <= kMaxFanJoints joints
    int slots;              // sub-model vertices per sensor block (8 or 12); a block is slots * 3 floats of vp
    int max_deg;            // largest sensor valence
    int n_part;             // number of (sensor, joint) pairs = sum of n_joints
    const int* deg;         // [12] valence (= faces = neighbours) of each sensor vertex
    const int* helper;      // [12] fan index (0..deg-1) of the helper vertex (virtual_sensors.py:47-59)
    const int* n_joints;    // [12]
    const int* part_ptr;    // [13] first partial of each sensor
    const int* joint;       // [12][kMaxFanJoints]
    const float* weight;    // [12][kMaxFanJoints][slots] skinning weight of (joint u, ring slot r), zero padded
    const int* jp_ptr;      // [23] joint -> range in jp_idx
    const int* jp_idx;      // partial indices that contribute to the joint, ascending
};

// Per-frame joint-level state (shared memory on the GPU).  `A[j]` = [A_j^R (= G_j^R) row-major | A_j^t]; `dA[j]` first
// holds dE/dA_j in the same layout and is turned in place into [dE/dG_j^R | dE/dG_j^t] by the reverse sweep.
template <typename T>
struct alignas(16) JointState {
    T A[kJoints][12];
    T dA[kJoints][12];
    T rot[kJoints][9];        // R_j = exp(theta_j)
    T jrest[kJoints][3];      // J(beta)
    T theta[kPoseDim];
    T gpos[kJoints][3];       // posed joints G_j^t; the FK-loss upstream gradient replaces it in place (training)
};
// Variable part that follows a JointState: first the partial sums [n_part][12], later (they are dead by then)
// dE/dR_j [22][9] followed by dE/dJ_j [22][3].
EMPOSE_HD constexpr int fan_var_floats(int n_part) { return (n_part * 12 > kJoints * 12 ? n_part * 12 : kJoints * 12); }

// ----------------------------------------------------------------------------------------------
// joint-level forward
// ----------------------------------------------------------------------------------------------
template <typename T>
EMPOSE_HD void jt_rodrigues(JointState<T>& st, int j) { rodrigues_fwd(&st.theta[j * 3], st.rot[j]); }

// kinematic chain, row r of every transform (three independent lanes per frame), standard SMPL tree in registers
template <typename T>
EMPOSE_HD void jt_chain_static(JointState<T>& st, int r) {
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
        st.A[j][r * 3] = g[j][0]; st.A[j][r * 3 + 1] = g[j][1]; st.A[j][r * 3 + 2] = g[j][2];
        st.gpos[j][r] = t[j];
        st.A[j][9 + r] = t[j] - (g[j][0] * st.jrest[j][0] + g[j][1] * st.jrest[j][1] + g[j][2] * st.jrest[j][2]);
    }
}
// same for an arbitrary (topologically ordered) tree
template <typename T>
EMPOSE_HD void jt_chain(const int* parents, JointState<T>& st, int r) {
    for (int c = 0; c < 3; ++c) st.A[0][r * 3 + c] = st.rot[0][r * 3 + c];
    st.gpos[0][r] = st.jrest[0][r];
    for (int j = 1; j < kJoints; ++j) {
        const int p = parents[j];
        const T g0 = st.A[p][r * 3], g1 = st.A[p][r * 3 + 1], g2 = st.A[p][r * 3 + 2];
        const T* R = st.rot[j];
        st.A[j][r * 3 + 0] = g0 * R[0] + g1 * R[3] + g2 * R[6];
        st.A[j][r * 3 + 1] = g0 * R[1] + g1 * R[4] + g2 * R[7];
        st.A[j][r * 3 + 2] = g0 * R[2] + g1 * R[5] + g2 * R[8];
        st.gpos[j][r] = g0 * (st.jrest[j][0] - st.jrest[p][0]) + g1 * (st.jrest[j][1] - st.jrest[p][1]) +
                        g2 * (st.jrest[j][2] - st.jrest[p][2]) + st.gpos[p][r];
    }
    for (int j = 0; j < kJoints; ++j)
        st.A[j][9 + r] = st.gpos[j][r] - (st.A[j][r * 3] * st.jrest[j][0] + st.A[j][r * 3 + 1] * st.jrest[j][1] +
                                          st.A[j][r * 3 + 2] * st.jrest[j][2]);
}

// ----------------------------------------------------------------------------------------------
// the (frame, sensor) item
// ----------------------------------------------------------------------------------------------
// 12 floats [a[0..9) = A^R row-major | a[9..12) = A^t] of one joint; 16-byte aligned, so three 128-bit loads on the GPU
template <typename T>
EMPOSE_HD void fan_load12(const T* src, T (&a)[12]) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        const float4 q0 = s4[0], q1 = s4[1], q2 = s4[2];
        a[0] = q0.x; a[1] = q0.y; a[2] = q0.z; a[3] = q0.w; a[4] = q1.x; a[5] = q1.y; a[6] = q1.z; a[7] = q1.w;
        a[8] = q2.x; a[9] = q2.y; a[10] = q2.z; a[11] = q2.w;
        return;
    }
#endif
    for (int i = 0; i < 12; ++i) a[i] = src[i];
}
template <typename T>
EMPOSE_HD void fan_store12(T* dst, const T (&a)[12]) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        d4[0] = make_float4(a[0], a[1], a[2], a[3]);
        d4[1] = make_float4(a[4], a[5], a[6], a[7]);
        d4[2] = make_float4(a[8], a[9], a[10], a[11]);
        return;
    }
#endif
    for (int i = 0; i < 12; ++i) dst[i] = a[i];
}
// four consecutive values (16-byte aligned): one 128-bit access on the GPU
template <typename T>
EMPOSE_HD void fan_load4(const T* src, T (&a)[4]) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        const float4 q = *reinterpret_cast<const float4*>(src);
        a[0] = q.x; a[1] = q.y; a[2] = q.z; a[3] = q.w;
        return;
    }
#endif
    for (int i = 0; i < 4; ++i) a[i] = src[i];
}
template <typename T>
EMPOSE_HD void fan_store4(T* dst, const T (&a)[4]) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(a[0], a[1], a[2], a[3]);
        return;
    }
#endif
    for (int i = 0; i < 4; ++i) dst[i] = a[i];
}
// one weight row [slots] of the dense (joint, slot) table -> registers
template <typename T, int SLOTS>
EMPOSE_HD void fan_load_weights(const float* w, T (&wr)[SLOTS]) {
#if defined(__CUDA_ARCH__)
    const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
    for (int q = 0; q < SLOTS / 4; ++q) {
        const float4 t = __ldg(w4 + q);
        wr[4 * q] = T(t.x); wr[4 * q + 1] = T(t.y); wr[4 * q + 2] = T(t.z); wr[4 * q + 3] = T(t.w);
    }
#else
    for (int r = 0; r < SLOTS; ++r) wr[r] = T(w[r]);
#endif
}

// Sensor s of one frame.
//   A       [22][12] joint transforms of the frame (JointState::A)
//   vp      this sensor's block of the blended rest vertices, (MAXD+1) * 3 values used
//   off     [12] = R_off row-major (9) | t_off (3)         (models.py:478-479)
//   meas    [12] = measured position (3) | orientation row-major (9); only read when want_grad
//   out_pos [3], out_ori [9]  p'_m, R'_m
//   dvp     (want_grad) dE/dvp of the ring, (MAXD+1) * 3 values
//   part    (want_grad) the frame's partial-sum list; entries part_ptr[s] .. part_ptr[s] + n_joints[s] are written
// SLOTS = vertices per sensor block (row pitch of the weight table), MAXD = largest valence handled (< SLOTS).
template <typename T, int SLOTS, int MAXD>
EMPOSE_HD void fan_sensor_item(const FanModel& fm, int s, const T* A, const T (&vp)[(MAXD + 1) * 3], const T (&off)[12],
                               const T (&meas)[12], const ResidualSpec& spec, bool want_grad, T (&out_pos)[3], T (&out_ori)[9],
                               T (&dvp)[(MAXD + 1) * 3], T* part) {
    constexpr int RING = MAXD + 1;
    static_assert(RING <= SLOTS, "ring must fit the sensor block");
    const int deg = fm.deg[s], nu = fm.n_joints[s], hx = fm.helper[s];
    const int* joints = fm.joint + s * kMaxFanJoints;
    const float* wtab = fm.weight + (size_t)s * kMaxFanJoints * SLOTS;

    // ---- linear blend skinning of the ring: x_r = sum_u w[u][r] (A_u^R vp_r + A_u^t) ----
    T x[RING * 3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < RING * 3; ++i) x[i] = T(0);
    for (int u = 0; u < nu; ++u) {
        T a[12], w[SLOTS];
        fan_load12(A + joints[u] * 12, a);
        fan_load_weights<T, SLOTS>(wtab + u * SLOTS, w);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < RING; ++r) {
            const T p0 = vp[r * 3], p1 = vp[r * 3 + 1], p2 = vp[r * 3 + 2];
            x[r * 3 + 0] += w[r] * (a[0] * p0 + a[1] * p1 + a[2] * p2 + a[9]);
            x[r * 3 + 1] += w[r] * (a[3] * p0 + a[4] * p1 + a[5] * p2 + a[10]);
            x[r * 3 + 2] += w[r] * (a[6] * p0 + a[7] * p1 + a[8] * p2 + a[11]);
        }
    }
    // ---- fan edges and the area-weighted normal ----
    const T xs[3] = {x[0], x[1], x[2]};
    T e[MAXD * 3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int d = 0; d < MAXD; ++d) {
        e[d * 3] = x[(d + 1) * 3] - xs[0]; e[d * 3 + 1] = x[(d + 1) * 3 + 1] - xs[1]; e[d * 3 + 2] = x[(d + 1) * 3 + 2] - xs[2];
    }
    T n[3] = {T(0), T(0), T(0)}, u[3] = {T(0), T(0), T(0)}, e_last[3] = {T(0), T(0), T(0)};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int d = 0; d < MAXD; ++d) {
        constexpr int kNextSentinel = 0;
        const int dn = d + 1 < MAXD ? d + 1 : kNextSentinel;
        const bool wrap = d + 1 >= deg;                       // the fan closes on edge 0
        const T b0 = wrap ? e[0] : e[dn * 3], b1 = wrap ? e[1] : e[dn * 3 + 1], b2 = wrap ? e[2] : e[dn * 3 + 2];
        const T a0 = e[d * 3], a1 = e[d * 3 + 1], a2 = e[d * 3 + 2];
        if (d < deg) {
            n[0] += a1 * b2 - a2 * b1;
            n[1] += a2 * b0 - a0 * b2;
            n[2] += a0 * b1 - a1 * b0;
        }
        if (d == hx) { u[0] = a0; u[1] = a1; u[2] = a2; }
        if (d == deg - 1) { e_last[0] = a0; e_last[1] = a1; e_last[2] = a2; }
    }
    const T inv_deg = T(1) / T(deg);
    n[0] *= inv_deg; n[1] *= inv_deg; n[2] *= inv_deg;
    // ---- sensor frame [on_surface | third | normal] (virtual_sensors.py:23-36), offsets (models.py:478-479) ----
    T nh[3], s0[3], t[3], th[3], sv[3], sh[3];
    const T n_len = normalize3(n, nh);
    const T u_len = normalize3(u, s0);
    cross3(nh, s0, t);
    const T t_len = normalize3(t, th);
    cross3(th, nh, sv);
    const T s_len = normalize3(sv, sh);
    const T R[9] = {sh[0], th[0], nh[0], sh[1], th[1], nh[1], sh[2], th[2], nh[2]};
    T Rc[9], pc[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rc[i * 3 + j] = R[i * 3] * off[j] + R[i * 3 + 1] * off[3 + j] + R[i * 3 + 2] * off[6 + j];
        pc[i] = xs[i] + R[i * 3] * off[9] + R[i * 3 + 1] * off[10] + R[i * 3 + 2] * off[11];
    }
    for (int i = 0; i < 9; ++i) out_ori[i] = Rc[i];
    for (int i = 0; i < 3; ++i) out_pos[i] = pc[i];
    if (!want_grad) return;

    const int p0i = fm.part_ptr[s];
    if (!spec.sensor_active[s]) {                           // not part of the residual (6-sensor models): zero gradient
        for (int i = 0; i < RING * 3; ++i) dvp[i] = T(0);
        T z[12];
        for (int i = 0; i < 12; ++i) z[i] = T(0);
        for (int q = 0; q < nu; ++q) fan_store12(part + (size_t)(p0i + q) * 12, z);
        return;
    }
    // ---- residual direction (loss.py:27-28) and its way back to the frame ----
    T dpc[3] = {T(0), T(0), T(0)}, dRc[9];
    for (int i = 0; i < 9; ++i) dRc[i] = T(0);
    if (spec.use_pos) {
        const T d[3] = {pc[0] - meas[0], pc[1] - meas[1], pc[2] - meas[2]};
        const T len = sqrt_t(dot3(d, d));
        const T inv = len > T(0) ? T(spec.weight) / len : T(0);   // reference: NaN at exactly zero residual (sqrt backward); we emit 0
        dpc[0] = d[0] * inv; dpc[1] = d[1] * inv; dpc[2] = d[2] * inv;
    }
    if (spec.use_ori) {
        T d[9], sq = T(0);
        for (int i = 0; i < 9; ++i) { d[i] = Rc[i] - meas[3 + i]; sq += d[i] * d[i]; }
        const T len = sqrt_t(sq);
        const T inv = len > T(0) ? T(spec.weight) / len : T(0);
        for (int i = 0; i < 9; ++i) dRc[i] = d[i] * inv;
    }
    T dR[9];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            dR[i * 3 + j] = dRc[i * 3] * off[j * 3] + dRc[i * 3 + 1] * off[j * 3 + 1] + dRc[i * 3 + 2] * off[j * 3 + 2] + dpc[i] * off[9 + j];
    T dsh[3] = {dR[0], dR[3], dR[6]}, dth[3] = {dR[1], dR[4], dR[7]}, dnh[3] = {dR[2], dR[5], dR[8]};
    T dsv[3], tmp[3];
    normalize3_bwd(sh, s_len, dsh, dsv);
    cross3(nh, dsv, tmp); dth[0] += tmp[0]; dth[1] += tmp[1]; dth[2] += tmp[2];      // sv = th x nh
    cross3(dsv, th, tmp); dnh[0] += tmp[0]; dnh[1] += tmp[1]; dnh[2] += tmp[2];
    T dt[3], ds0[3];
    normalize3_bwd(th, t_len, dth, dt);
    cross3(s0, dt, tmp); dnh[0] += tmp[0]; dnh[1] += tmp[1]; dnh[2] += tmp[2];       // t = nh x s0
    cross3(dt, nh, ds0);
    T du[3], dN[3];
    normalize3_bwd(s0, u_len, ds0, du);
    normalize3_bwd(nh, n_len, dnh, dN);
    dN[0] *= inv_deg; dN[1] *= inv_deg; dN[2] *= inv_deg;
    // ---- dE/dx of the ring: edges get (e_{d+1} - e_{d-1}) x dE/dN, the helper edge +du, the sensor vertex dpc - du ----
    T dx[RING * 3];
    dx[0] = dpc[0] - du[0]; dx[1] = dpc[1] - du[1]; dx[2] = dpc[2] - du[2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int d = 0; d < MAXD; ++d) {
        const int dn = d + 1 < MAXD ? d + 1 : 0;
        const bool wrap = d + 1 >= deg;
        T c0 = (wrap ? e[0] : e[dn * 3]), c1 = (wrap ? e[1] : e[dn * 3 + 1]), c2 = (wrap ? e[2] : e[dn * 3 + 2]);
        if (d == 0) { c0 -= e_last[0]; c1 -= e_last[1]; c2 -= e_last[2]; }
        else { c0 -= e[(d - 1) * 3]; c1 -= e[(d - 1) * 3 + 1]; c2 -= e[(d - 1) * 3 + 2]; }
        T g0 = c1 * dN[2] - c2 * dN[1], g1 = c2 * dN[0] - c0 * dN[2], g2 = c0 * dN[1] - c1 * dN[0];
        if (d >= deg) { g0 = T(0); g1 = T(0); g2 = T(0); }
        if (d == hx) { g0 += du[0]; g1 += du[1]; g2 += du[2]; }
        dx[(d + 1) * 3] = g0; dx[(d + 1) * 3 + 1] = g1; dx[(d + 1) * 3 + 2] = g2;
    }
    // ---- reverse skinning: dE/dvp_r = sum_u w A_u^R^T dx_r;  dE/dA_u = sum_r w dx_r [vp_r^T | 1] ----
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < RING * 3; ++i) dvp[i] = T(0);
    for (int q = 0; q < nu; ++q) {
        T a[12], w[SLOTS], acc[12];
        fan_load12(A + joints[q] * 12, a);
        fan_load_weights<T, SLOTS>(wtab + q * SLOTS, w);
        for (int i = 0; i < 12; ++i) acc[i] = T(0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < RING; ++r) {
            const T d0 = w[r] * dx[r * 3], d1 = w[r] * dx[r * 3 + 1], d2 = w[r] * dx[r * 3 + 2];
            const T p0 = vp[r * 3], p1 = vp[r * 3 + 1], p2 = vp[r * 3 + 2];
            dvp[r * 3 + 0] += a[0] * d0 + a[3] * d1 + a[6] * d2;
            dvp[r * 3 + 1] += a[1] * d0 + a[4] * d1 + a[7] * d2;
            dvp[r * 3 + 2] += a[2] * d0 + a[5] * d1 + a[8] * d2;
            acc[0] += d0 * p0; acc[1] += d0 * p1; acc[2] += d0 * p2;
            acc[3] += d1 * p0; acc[4] += d1 * p1; acc[5] += d1 * p2;
            acc[6] += d2 * p0; acc[7] += d2 * p1; acc[8] += d2 * p2;
            acc[9] += d0; acc[10] += d1; acc[11] += d2;
        }
        fan_store12(part + (size_t)(p0i + q) * 12, acc);
    }
}

// ----------------------------------------------------------------------------------------------
// joint-level reverse
// ----------------------------------------------------------------------------------------------
// dE/dA_j[e] = sum of the partials that name joint j, in list order (22 * 12 items)
template <typename T>
EMPOSE_HD void jt_reduce(const FanModel& fm, JointState<T>& st, const T* part, int it) {
    const int j = it / 12, e = it - j * 12;
    T acc = T(0);
    for (int q = fm.jp_ptr[j]; q < fm.jp_ptr[j + 1]; ++q) acc += part[fm.jp_idx[q] * 12 + e];
    st.dA[j][e] = acc;
}
// The same reduction for ALL frames of a CTA in the form the kernel runs: thread `tid` of `nt` owns a (joint, quarter)
// item of EVERY frame, so the index lists are read once and every access is 16 bytes wide (22 * 3 items).
// state(f) -> JointState<T>&, var_of(f) -> T* (the frame's partial sums).
template <typename T, typename StateFn, typename VarFn>
EMPOSE_HD void jt_reduce_frames(const FanModel& fm, StateFn state, VarFn var_of, int nf, int tid, int nt) {
    for (int it = tid; it < kJoints * 3; it += nt) {
        const int j = it / 3, q4 = (it - j * 3) * 4;
        const int q0 = fm.jp_ptr[j], n = fm.jp_ptr[j + 1] - q0;
        int idx[4] = {0, 0, 0, 0};                     // the first four partials of the joint stay in registers
        for (int k = 0; k < 4; ++k) if (k < n) idx[k] = fm.jp_idx[q0 + k] * 12 + q4;
#if defined(__CUDA_ARCH__)
#pragma unroll 4          // independent frames: four frames' loads in flight instead of one (the phase is latency bound)
#endif
        for (int f = 0; f < nf; ++f) {
            const T* part = var_of(f);
            T acc[4] = {T(0), T(0), T(0), T(0)}, v[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 4; ++k)
                if (k < n) { fan_load4(part + idx[k], v); acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3]; }
            for (int k = 4; k < n; ++k) {
                fan_load4(part + fm.jp_idx[q0 + k] * 12 + q4, v);
                acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
            }
            fan_store4(&state(f).dA[j][q4], acc);
        }
    }
}
// upstream gradient of the FK loss sum_j ||J_j - Jgt_j|| (models.py:657-660) times `weight`, in place over gpos (22 items)
template <typename T, typename TIn>
EMPOSE_HD void jt_joint_residual(JointState<T>& st, const TIn* joints_gt, T weight, int j) {
    const T d[3] = {st.gpos[j][0] - T(joints_gt[j * 3]), st.gpos[j][1] - T(joints_gt[j * 3 + 1]), st.gpos[j][2] - T(joints_gt[j * 3 + 2])};
    const T len = sqrt_t(dot3(d, d));
    const T inv = len > T(0) ? weight / len : T(0);
    st.gpos[j][0] = d[0] * inv; st.gpos[j][1] = d[1] * inv; st.gpos[j][2] = d[2] * inv;
}
// reverse sweep of the chain, row r (frame_math.h item_chain_bwd_static / item_chain_bwd), in place over dA
template <typename T>
EMPOSE_HD void jt_chain_bwd_static(JointState<T>& st, int r, bool joint_up) {
    T acc[kJoints][3], acct[kJoints];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < kJoints; ++j) { acc[j][0] = T(0); acc[j][1] = T(0); acc[j][2] = T(0); acct[j] = T(0); }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = kJoints - 1; j >= 0; --j) {
        const T a = st.dA[j][9 + r];
        const T dt = joint_up ? a + st.gpos[j][r] + acct[j] : a + acct[j];
        const T d0 = st.dA[j][r * 3] - a * st.jrest[j][0] + acc[j][0];
        const T d1 = st.dA[j][r * 3 + 1] - a * st.jrest[j][1] + acc[j][1];
        const T d2 = st.dA[j][r * 3 + 2] - a * st.jrest[j][2] + acc[j][2];
        st.dA[j][r * 3] = d0; st.dA[j][r * 3 + 1] = d1; st.dA[j][r * 3 + 2] = d2;
        st.dA[j][9 + r] = dt;
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
template <typename T>
EMPOSE_HD void jt_chain_bwd(const int* parents, JointState<T>& st, int r, bool joint_up) {
    for (int j = 0; j < kJoints; ++j) {
        const T a = st.dA[j][9 + r];
        for (int c = 0; c < 3; ++c) st.dA[j][r * 3 + c] -= a * st.jrest[j][c];
        if (joint_up) st.dA[j][9 + r] = a + st.gpos[j][r];
    }
    for (int j = kJoints - 1; j >= 1; --j) {
        const int p = parents[j];
        const T* R = st.rot[j];
        const T d0 = st.dA[j][r * 3], d1 = st.dA[j][r * 3 + 1], d2 = st.dA[j][r * 3 + 2];
        const T dt = st.dA[j][9 + r];
        st.dA[p][r * 3 + 0] += d0 * R[0] + d1 * R[1] + d2 * R[2] + dt * (st.jrest[j][0] - st.jrest[p][0]);
        st.dA[p][r * 3 + 1] += d0 * R[3] + d1 * R[4] + d2 * R[5] + dt * (st.jrest[j][1] - st.jrest[p][1]);
        st.dA[p][r * 3 + 2] += d0 * R[6] + d1 * R[7] + d2 * R[8] + dt * (st.jrest[j][2] - st.jrest[p][2]);
        st.dA[p][9 + r] += dt;
    }
}
// local gradients from the final dE/dG (frame_math.h item_chain_bwd_local): item i = j * 12 + e writes
// var[j * 9 + e] = dE/dR_j (e < 9) or var[198 + j * 3 + e - 9] = dE/dJ_j
template <typename T>
EMPOSE_HD void jt_local(const int* parents, const JointState<T>& st, T* var, int i, bool joint_up) {
    const int j = i / 12, e = i - j * 12;
    const int p = parents[j];
    if (e < 9) {
        const int a = e / 3, b = e - a * 3;
        T acc;
        if (j == 0) acc = st.dA[0][e];
        else acc = st.A[p][a] * st.dA[j][b] + st.A[p][3 + a] * st.dA[j][3 + b] + st.A[p][6 + a] * st.dA[j][6 + b];
        var[j * 9 + e] = acc;
    } else {
        const int c = e - 9;
        T acc = T(0);
        for (int r = 0; r < 3; ++r) {
            const T gp = (j == 0) ? (r == c ? T(1) : T(0)) : st.A[p][r * 3 + c];
            acc += (gp - st.A[j][r * 3 + c]) * st.dA[j][9 + r];
            if (joint_up) acc += st.A[j][r * 3 + c] * st.gpos[j][r];
        }
        var[kJoints * 9 + j * 3 + c] = acc;
    }
}
// jt_local for ALL frames of a CTA: thread `tid` of `nt` owns (joint j, column a) of EVERY frame (22 * 3 items): row a of
// dE/dR_j = G_p^T dE/dG_j^R and entry a of dE/dJ_j.  One column of G_p and of G_j, all of dE/dG_j per item and frame.
template <typename T, typename StateFn, typename VarFn>
EMPOSE_HD void jt_local_frames(const int* parents, StateFn state, VarFn var_of, int nf, int tid, int nt, bool joint_up) {
    for (int it = tid; it < kJoints * 3; it += nt) {
        const int j = it / 3, a = it - j * 3;
        const int p = j > 0 ? parents[j] : 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
        for (int f = 0; f < nf; ++f) {
            const JointState<T>& st = state(f);
            T* var = var_of(f);
            T d[12];
            fan_load12(&st.dA[j][0], d);
            const T gj0 = st.A[j][a], gj1 = st.A[j][3 + a], gj2 = st.A[j][6 + a];
            T gp0 = a == 0 ? T(1) : T(0), gp1 = a == 1 ? T(1) : T(0), gp2 = a == 2 ? T(1) : T(0);   // root: G_p = I
            if (j > 0) { gp0 = st.A[p][a]; gp1 = st.A[p][3 + a]; gp2 = st.A[p][6 + a]; }
            var[j * 9 + a * 3 + 0] = gp0 * d[0] + gp1 * d[3] + gp2 * d[6];
            var[j * 9 + a * 3 + 1] = gp0 * d[1] + gp1 * d[4] + gp2 * d[7];
            var[j * 9 + a * 3 + 2] = gp0 * d[2] + gp1 * d[5] + gp2 * d[8];
            T dj = (gp0 - gj0) * d[9] + (gp1 - gj1) * d[10] + (gp2 - gj2) * d[11];
            if (joint_up) dj += gj0 * st.gpos[j][0] + gj1 * st.gpos[j][1] + gj2 * st.gpos[j][2];
            var[kJoints * 9 + j * 3 + a] = dj;
        }
    }
}
// chain part of dE/dtheta_j, times coef (22 items); the pose-blend part is added by the caller (the map is linear in dR)
template <typename T, typename TOut>
EMPOSE_HD void jt_finish_theta(const JointState<T>& st, const T* var, T coef, TOut* g_theta, int j) {
    T g[3] = {T(0), T(0), T(0)};
    rodrigues_bwd(&st.theta[j * 3], var + j * 9, g);
    g_theta[j * 3] = TOut(coef * g[0]); g_theta[j * 3 + 1] = TOut(coef * g[1]); g_theta[j * 3 + 2] = TOut(coef * g[2]);
}

}  // namespace empose
