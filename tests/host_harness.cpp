// CPU harness around em-pose_b200/csrc/frame_math.h (TEST INFRASTRUCTURE, never shipped).
// The per-frame forward / reverse arithmetic of the CUDA kernels is written host+device; this file
// compiles it with g++ and runs it with one "lane" so that tests/test_frame_math.py can compare the
// hand-derived reverse pass with the oracle's autograd without a GPU.
#include <vector>
#include <cstring>
#include "frame_math.h"

using namespace empose;

struct HostSub {
    int n_verts, vp_dim, n_faces, max_degree, n_skin;
    const float *v_template, *shapedirs, *posedirs, *j0, *jdirs, *skin_weight, *jt_weight;
    const int *skin_joint, *jt_ptr, *jt_vert, *parents, *faces, *sensor_vert, *helper_vert, *sensor_faces, *sensor_degree;
    const int *vj_ptr, *jvj_ptr, *vinc_ptr, *vinc_item, *vinc_code;
    int n_vj;
    int use_static_tree;
};

template <typename T>
static void run(const HostSub& h, int n_frames, const float* theta, const float* beta, const float* off_r,
                const float* off_t, const float* meas_pos, const float* meas_ori, const int* active, int use_pos,
                int use_ori, const float* coef, int want_grad, float sensor_weight, const float* joints_gt, float joint_weight,
                double* sensor_pos, double* sensor_ori, double* joints, double* g_theta, double* g_beta, double* verts) {
    SubModel m;
    m.n_verts = h.n_verts; m.vp_dim = h.vp_dim; m.n_faces = h.n_faces; m.max_degree = h.max_degree; m.n_skin = h.n_skin;
    m.v_template = h.v_template; m.shapedirs = h.shapedirs; m.j0 = h.j0; m.jdirs = h.jdirs;
    m.skin_weight = h.skin_weight; m.skin_joint = h.skin_joint; m.jt_ptr = h.jt_ptr; m.jt_vert = h.jt_vert;
    m.jt_weight = h.jt_weight; m.parents = h.parents; m.faces = h.faces; m.sensor_vert = h.sensor_vert;
    m.helper_vert = h.helper_vert; m.sensor_faces = h.sensor_faces; m.sensor_degree = h.sensor_degree;
    m.n_vj = h.n_vj; m.vj_ptr = h.vj_ptr; m.jvj_ptr = h.jvj_ptr;
    m.vinc_ptr = h.vinc_ptr; m.vinc_item = h.vinc_item; m.vinc_code = h.vinc_code;
    ResidualSpec spec;
    spec.use_pos = use_pos; spec.use_ori = use_ori; spec.weight = sensor_weight;
    for (int s = 0; s < kSensors; ++s) spec.sensor_active[s] = active[s];
    bool use_static = h.use_static_tree != 0;
    for (int j = 0; j < kJoints; ++j) use_static = use_static && (h.parents[j] == smpl_parent(j));
    std::vector<FrameState<T>> st_store(1);
    FrameState<T>& st = st_store[0];
    std::vector<T> vp_off(h.vp_dim), dpf(kPoseFeatPad);
    for (int f = 0; f < n_frames; ++f) {
        for (int i = 0; i < kPoseDim; ++i) st.theta[i] = T(theta[f * kPoseDim + i]);
        for (int i = 0; i < kBetas; ++i) st.beta[i] = T(beta[f * kBetas + i]);
        // pose blend (the GPU path does this as a GEMM): pf = vec(R_1..R_21 - I)
        T pf[kPoseFeat];
        for (int j = 1; j < kJoints; ++j) {
            T R[9];
            rodrigues_fwd(&st.theta[j * 3], R);
            for (int e = 0; e < 9; ++e) pf[(j - 1) * 9 + e] = R[e] - ((e % 4 == 0) ? T(1) : T(0));
        }
        for (int i = 0; i < h.vp_dim; ++i) {
            T acc = T(0);
            for (int k = 0; k < kPoseFeat; ++k) acc += T(h.posedirs[k * h.vp_dim + i]) * pf[k];
            vp_off[i] = acc;
        }
        phase_rodrigues(st, 0, 1);
        phase_rest_joints(m, st, 0, 1);
        phase_blend_verts(m, st, vp_off.data(), 0, 1);
        if (use_static) phase_chain_static(st, 0, 1); else phase_chain(m, st, 0, 1);
        phase_skin(m, st, 0, 1);
        if (h.use_static_tree && m.max_degree <= kSplitDegree) {       // the split form the GPU kernel uses
            phase_sensor_faces(m, st, 0, 1);
            phase_sensor_frames(m, st, off_r + f * 108, off_t + f * 36, meas_pos + f * 36, meas_ori + f * 108, spec,
                                want_grad != 0, 0, 1);
            if (want_grad) { phase_sensor_face_grads(m, st, 0, 1); phase_sensor_gather(m, st, 0, 1); }
        } else {
            phase_sensors(m, st, off_r + f * 108, off_t + f * 36, meas_pos + f * 36, meas_ori + f * 108, spec,
                          want_grad != 0, 0, 1);
        }
        for (int i = 0; i < 36; ++i) sensor_pos[f * 36 + i] = double(st.sensor_pos[i / 3][i % 3]);
        for (int i = 0; i < 108; ++i) sensor_ori[f * 108 + i] = double(st.sensor_ori[i / 9][i % 9]);
        for (int i = 0; i < 66; ++i) joints[f * 66 + i] = double(st.gpos[i / 3][i % 3]);
        if (verts) for (int i = 0; i < h.n_verts * 3; ++i) verts[f * h.n_verts * 3 + i] = double(st.x[i]);
        if (!want_grad) continue;
        if (joints_gt) phase_joint_residual(st, joints_gt + f * kPoseDim, T(joint_weight), 0, 1);
        phase_skin_bwd_chunks(m, st, 0, 1);
        phase_skin_bwd_reduce(m, st, 0, 1);
        phase_skin_bwd_verts(m, st, 0, 1);
        phase_shape_bwd_partial(m, st, 0, 1);
        for (int k = 0; k < kPoseFeat; ++k) {
            T acc = T(0);
            for (int i = 0; i < h.n_verts * 3; ++i) acc += T(h.posedirs[k * h.vp_dim + i]) * st.dx[i];
            dpf[k] = acc;
        }
        if (use_static) phase_chain_bwd_static(st, 0, 1, joints_gt != nullptr); else phase_chain_bwd(m, st, 0, 1, joints_gt != nullptr);
        phase_chain_bwd_local(m, st, 0, 1, joints_gt != nullptr);
        std::vector<T> gt(kPoseDim), gb(kBetas);
        phase_finish_theta(st, T(coef[f]), dpf.data(), gt.data(), 0, 1);
        phase_finish_beta(m, st, T(coef[f]), gb.data(), 0, 1);
        for (int i = 0; i < kPoseDim; ++i) g_theta[f * kPoseDim + i] = double(gt[i]);
        for (int i = 0; i < kBetas; ++i) g_beta[f * kBetas + i] = double(gb[i]);
    }
}

extern "C" int host_frame_eval(const HostSub* h, int n_frames, const float* theta, const float* beta,
                               const float* off_r, const float* off_t, const float* meas_pos, const float* meas_ori,
                               const int* active, int use_pos, int use_ori, const float* coef, int want_grad,
                               int use_double, float sensor_weight, const float* joints_gt, float joint_weight,
                               double* sensor_pos, double* sensor_ori, double* joints, double* g_theta,
                               double* g_beta, double* verts) {
    if (h->vp_dim > kMaxVp || h->max_degree > kMaxDegree || h->n_vj > kMaxVj) return -1;
    if (use_double)
        run<double>(*h, n_frames, theta, beta, off_r, off_t, meas_pos, meas_ori, active, use_pos, use_ori, coef,
                    want_grad, sensor_weight, joints_gt, joint_weight, sensor_pos, sensor_ori, joints, g_theta, g_beta, verts);
    else
        run<float>(*h, n_frames, theta, beta, off_r, off_t, meas_pos, meas_ori, active, use_pos, use_ori, coef,
                   want_grad, sensor_weight, joints_gt, joint_weight, sensor_pos, sensor_ori, joints, g_theta, g_beta, verts);
    return 0;
}

// ---- fan form (csrc/fan_math.h): the per-(frame, sensor) arithmetic of the production kernel -------------------------
#include "fan_math.h"

struct HostFan {
    int ok, slots, max_deg, n_part;
    const int *deg, *helper, *n_joints, *part_ptr, *joint;
    const float* weight;
    const int *jp_ptr, *jp_idx;
};

template <typename T, int SLOTS, int MAXD>
static void run_fan(const HostSub& h, const HostFan& hf, int n_frames, const float* theta, const float* beta, const float* off_r,
                    const float* off_t, const float* meas_pos, const float* meas_ori, const int* active, int use_pos,
                    int use_ori, const float* coef, int want_grad, float sensor_weight, const float* joints_gt, float joint_weight,
                    double* sensor_pos, double* sensor_ori, double* joints, double* g_theta, double* g_beta) {
    constexpr int RING = MAXD + 1;
    FanModel fm;
    fm.ok = hf.ok & 1; fm.slots = hf.slots; fm.max_deg = hf.max_deg; fm.n_part = hf.n_part;
    fm.deg = hf.deg; fm.helper = hf.helper; fm.n_joints = hf.n_joints; fm.part_ptr = hf.part_ptr; fm.joint = hf.joint;
    fm.weight = hf.weight; fm.jp_ptr = hf.jp_ptr; fm.jp_idx = hf.jp_idx;
    ResidualSpec spec;
    spec.use_pos = use_pos; spec.use_ori = use_ori; spec.weight = sensor_weight;
    for (int s = 0; s < kSensors; ++s) spec.sensor_active[s] = active[s];
    bool use_static = h.use_static_tree != 0;
    for (int j = 0; j < kJoints; ++j) use_static = use_static && (h.parents[j] == smpl_parent(j));
    const bool joint_up = joints_gt != nullptr;
    std::vector<JointState<T>> store(1);
    JointState<T>& st = store[0];
    std::vector<T> var(fan_var_floats(hf.n_part)), vp(h.vp_dim), dvp_all(h.vp_dim), feat(kPoseFeatPad);
    for (int f = 0; f < n_frames; ++f) {
        for (int i = 0; i < kPoseDim; ++i) st.theta[i] = T(theta[f * kPoseDim + i]);
        // the blend GEMM: feature row [pf | beta] against [P ; S] and [0 ; Jdirs], bias [v_template | J0]
        for (int k = 0; k < kPoseFeatPad; ++k) feat[k] = T(0);
        for (int j = 0; j < kJoints; ++j) jt_rodrigues(st, j);
        for (int j = 1; j < kJoints; ++j)
            for (int e = 0; e < 9; ++e) feat[(j - 1) * 9 + e] = st.rot[j][e] - ((e % 4 == 0) ? T(1) : T(0));
        for (int k = 0; k < kBetas; ++k) feat[kFeatBeta + k] = T(beta[f * kBetas + k]);
        for (int i = 0; i < h.vp_dim; ++i) {
            T acc = T(0);
            for (int k = 0; k < kPoseFeat; ++k) acc += T(h.posedirs[k * h.vp_dim + i]) * feat[k];
            for (int k = 0; k < kBetas; ++k) acc += T(h.shapedirs[k * h.vp_dim + i]) * feat[kFeatBeta + k];
            vp[i] = acc + T(h.v_template[i]);                 // the template is the bias of the GEMM
        }
        for (int i = 0; i < kPoseDim; ++i) {
            T acc = T(0);
            for (int k = 0; k < kBetas; ++k) acc += T(h.jdirs[k * kPoseDim + i]) * feat[kFeatBeta + k];
            st.jrest[i / 3][i % 3] = acc + T(h.j0[i]);
        }
        for (int r = 0; r < 3; ++r) { if (use_static) jt_chain_static(st, r); else jt_chain(h.parents, st, r); }
        for (int i = 0; i < h.vp_dim; ++i) dvp_all[i] = T(0);
        for (int s = 0; s < kSensors; ++s) {
            T vps[RING * 3], off[12], meas[12], opos[3], oori[9], dvp[RING * 3];
            for (int i = 0; i < RING * 3; ++i) vps[i] = vp[s * SLOTS * 3 + i];
            for (int i = 0; i < 9; ++i) off[i] = T(off_r[f * 108 + s * 9 + i]);
            for (int i = 0; i < 3; ++i) off[9 + i] = T(off_t[f * 36 + s * 3 + i]);
            for (int i = 0; i < 3; ++i) meas[i] = T(meas_pos[f * 36 + s * 3 + i]);
            for (int i = 0; i < 9; ++i) meas[3 + i] = T(meas_ori[f * 108 + s * 9 + i]);
            fan_sensor_item<T, SLOTS, MAXD>(fm, s, &st.A[0][0], vps, off, meas, spec, want_grad != 0, opos, oori, dvp, var.data());
            for (int i = 0; i < 3; ++i) sensor_pos[f * 36 + s * 3 + i] = double(opos[i]);
            for (int i = 0; i < 9; ++i) sensor_ori[f * 108 + s * 9 + i] = double(oori[i]);
            if (want_grad) for (int i = 0; i < RING * 3; ++i) dvp_all[s * SLOTS * 3 + i] = dvp[i];
        }
        for (int i = 0; i < 66; ++i) joints[f * 66 + i] = double(st.gpos[i / 3][i % 3]);
        if (!want_grad) continue;
        auto state_of = [&](int) -> JointState<T>& { return st; };
        auto var_of = [&](int) -> T* { return var.data(); };
        if (hf.ok & 2) jt_reduce_frames<T>(fm, state_of, var_of, 1, 0, 1);          // the form the kernel runs
        else for (int it = 0; it < kJoints * 12; ++it) jt_reduce(fm, st, var.data(), it);
        if (joint_up) for (int j = 0; j < kJoints; ++j) jt_joint_residual(st, joints_gt + f * kPoseDim, T(joint_weight), j);
        for (int r = 0; r < 3; ++r) { if (use_static) jt_chain_bwd_static(st, r, joint_up); else jt_chain_bwd(h.parents, st, r, joint_up); }
        if (hf.ok & 2) jt_local_frames<T>(h.parents, state_of, var_of, 1, 0, 1, joint_up);
        else for (int i = 0; i < kJoints * 12; ++i) jt_local(h.parents, st, var.data(), i, joint_up);
        // the transposed GEMM: [dvp | dJ] against [P^T | 0], [S^T | Jdirs^T]
        T gt[kPoseDim];
        for (int j = 0; j < kJoints; ++j) jt_finish_theta(st, var.data(), T(coef[f]), gt, j);
        for (int j = 1; j < kJoints; ++j) {
            T dR[9], g[3] = {T(0), T(0), T(0)};
            for (int e = 0; e < 9; ++e) {
                T acc = T(0);
                for (int i = 0; i < h.vp_dim; ++i) acc += T(h.posedirs[((j - 1) * 9 + e) * h.vp_dim + i]) * dvp_all[i];
                dR[e] = acc * T(coef[f]);
            }
            rodrigues_bwd(&st.theta[j * 3], dR, g);
            for (int c = 0; c < 3; ++c) gt[j * 3 + c] += g[c];
        }
        for (int i = 0; i < kPoseDim; ++i) g_theta[f * kPoseDim + i] = double(gt[i]);
        for (int k = 0; k < kBetas; ++k) {
            T acc = T(0);
            for (int i = 0; i < h.vp_dim; ++i) acc += T(h.shapedirs[k * h.vp_dim + i]) * dvp_all[i];
            for (int i = 0; i < kPoseDim; ++i) acc += T(h.jdirs[k * kPoseDim + i]) * var[kJoints * 9 + i];
            g_beta[f * kBetas + k] = double(T(coef[f]) * acc);
        }
    }
}

extern "C" int host_fan_eval(const HostSub* h, const HostFan* hf, int n_frames, const float* theta, const float* beta,
                             const float* off_r, const float* off_t, const float* meas_pos, const float* meas_ori,
                             const int* active, int use_pos, int use_ori, const float* coef, int want_grad,
                             int use_double, int force_maxd, float sensor_weight, const float* joints_gt, float joint_weight,
                             double* sensor_pos, double* sensor_ori, double* joints, double* g_theta, double* g_beta) {
    if (!(hf->ok & 1) || h->vp_dim > kMaxVp) return -1;
    int maxd = force_maxd > 0 ? force_maxd : hf->max_deg;
#define EMPOSE_RUN_FAN(T, S, D)                                                                                                  \
    run_fan<T, S, D>(*h, *hf, n_frames, theta, beta, off_r, off_t, meas_pos, meas_ori, active, use_pos, use_ori, coef, want_grad, \
                     sensor_weight, joints_gt, joint_weight, sensor_pos, sensor_ori, joints, g_theta, g_beta)
    if (hf->slots == 8 && maxd <= 6) { if (use_double) EMPOSE_RUN_FAN(double, 8, 6); else EMPOSE_RUN_FAN(float, 8, 6); }
    else if (hf->slots == 8 && maxd <= 7) { if (use_double) EMPOSE_RUN_FAN(double, 8, 7); else EMPOSE_RUN_FAN(float, 8, 7); }
    else if (hf->slots == 12 && maxd <= 11) { if (use_double) EMPOSE_RUN_FAN(double, 12, 11); else EMPOSE_RUN_FAN(float, 12, 11); }
    else return -2;
#undef EMPOSE_RUN_FAN
    return 0;
}

// ---- evaluation metrics (csrc/metrics_math.h) on the host --------------------------------------------------------
#include "metrics_math.h"

extern "C" int host_metrics_eval(const float* j0, const float* jdirs, const int* parents, int n_frames, const float* pose,
                                 const float* shape, const float* pose_hat, const float* shape_hat, float* eucl, float* eucl_pa,
                                 float* joints_out) {
    MetricsParams p;
    memset(&p, 0, sizeof(p));
    p.j0 = j0; p.jdirs = jdirs; p.parents = parents;
    for (int f = 0; f < n_frames; ++f) {
        float jg[kJoints][3], jh[kJoints][3], og[kJoints][9], oh[kJoints][9];
        fk_frame(p, pose + f * kPoseDim, shape + f * kBetas, jg, og);
        fk_frame(p, pose_hat + f * kPoseDim, shape_hat + f * kBetas, jh, oh);
        joint_distances(jg, jh, eucl + f * kJoints, eucl_pa + f * kJoints);
        if (joints_out) memcpy(joints_out + f * kPoseDim, jg, sizeof(jg));
    }
    return 0;
}
