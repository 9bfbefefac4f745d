/*
 * empose_b200 -- C ABI of the B200-native EM-POSE learned-gradient-descent (LGD / IEF) hot path.
 *
 * The reference (facebookresearch/em-pose) has no FFI: its seam for this path is two Python
 * classes.  This header is the boundary a binding for those classes talks to; every entry point
 * names the reference interface it stands in for (paths are relative to the reference tree).
 *
 *   empose_ief_create / _destroy   <- create_model(config, smpl_model)            empose/nn/models.py:23-33
 *                                     + nn.Module.load_state_dict (state-dict keys of models.py:424-454,
 *                                       layers.py:13-77,114) + VirtualMarkerHelper topology
 *                                       (empose/data/virtual_sensors.py:47-75)
 *   empose_ief_forward             <- IterativeErrorFeedback.forward              empose/nn/models.py:485-632
 *   empose_ief_forward_host        <- same, host buffers (what scripts/evaluate_real.py:61 does around it:
 *                                     chunk.to_gpu(), net(chunk), .cpu())
 *   empose_ief_submit_host / _wait_host <- the same loop with two or more chunks in flight (copies hidden behind compute)
 *   empose_sensor_project          <- IterativeErrorFeedback.get_estimated_real_markers
 *                                                                                 empose/nn/models.py:471-483
 *   empose_sensors_create          <- SMPLFK + SampleMarkersWithOffsets (data synthesis before the hot path)
 *                                                                                 empose/data/transforms.py:163-226, 259-282
 *   empose_smpl_create / _forward  <- SMPLLayer.__init__ / forward / fk / _fk     empose/bodymodels/smpl.py:31-165
 *                                     (the third-party BodyModel call at smpl.py:121)
 *   empose_metrics_compute         <- MetricsEngine.compute (per-frame FK x2, Procrustes, angular distance)
 *                                                                                 empose/eval/metrics.py:183-241, 19-66
 *   empose_train_create / _layout  <- create_model + net.parameters() as ONE flat vector (scripts/train.py:125)
 *   empose_train_forward           <- IterativeErrorFeedback.forward with net.train()  empose/nn/models.py:485-632
 *                                     (BatchNorm1d on batch statistics, layers.py:26,57; the gradient side effect
 *                                     of models.py:576)
 *   empose_train_backward          <- IterativeErrorFeedback.backward                  empose/nn/models.py:634-688
 *   empose_rnn_create / _forward   <- create_model(m_type='rnn') + SimpleRNN.forward (the BiRNN baseline)
 *                                                                                 empose/nn/models.py:265-317
 *   empose_gemm_selftest           <- (no reference counterpart) checks the tcgen05 GEMM engine
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * EMPOSE_E_* code and never throws; empose_last_error() gives the message of the last failure on the
 * calling thread.  Unless a parameter says "host", data pointers are DEVICE pointers on the context's
 * device, borrowed for the duration of the call; work is enqueued on `stream` (a cudaStream_t passed
 * as void*) and is stream-ordered -- the call does not synchronise unless noted.  A context may be
 * used from one host thread at a time.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with EMPOSE_E_CUDA.
 *
 * Layouts are the reference's: row-major float32; poses are axis-angle [root(3) | 21 body joints];
 * sensors are in S_ORDER (empose/helpers/configuration.py:86-88); orientations are row-major 3x3.
 */
#ifndef EMPOSE_B200_H
#define EMPOSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMPOSE_ABI_VERSION 1

enum {
    EMPOSE_OK = 0,
    EMPOSE_E_ARG = -1,       /* bad argument / unsupported configuration (reference: ValueError / assert) */
    EMPOSE_E_MISSING = -2,   /* a required tensor is absent from the tensor table (reference: load_state_dict KeyError) */
    EMPOSE_E_SHAPE = -3,     /* tensor with the wrong shape or dtype */
    EMPOSE_E_CUDA = -4,      /* CUDA runtime / driver failure, or no device */
    EMPOSE_E_NOMEM = -5
};

enum { EMPOSE_F32 = 0, EMPOSE_I32 = 1, EMPOSE_I64 = 2 };

/* Arithmetic of the learned layers and the pose-blend contraction. */
enum {
    EMPOSE_PRECISION_TF32 = 0,   /* tcgen05.mma kind::tf32 everywhere, fp32 accumulate in TMEM (training default) */
    EMPOSE_PRECISION_FP32 = 1,   /* plain fp32 FFMA kernels: exact-arithmetic mode for parity studies */
    EMPOSE_PRECISION_FP16 = 2    /* inference: MLP / LSTM / heads operands in fp16 (kind::f16, the same 10-bit mantissa
                                    as tf32 at twice the rate and half the bytes), fp32 accumulate; the pose blend stays
                                    error-compensated tf32 */
};

/* One named host tensor.  Names are the reference's state-dict keys ("rnn.lstm.weight_ih_l0",
 * "pose_net_iter.hidden_layers.0.layers.1.running_var", ...) plus the "sub.*" arrays produced by
 * submodel.py (the SMPL-H sub-model and sensor topology).
 *
 * The sub-model is stored ring-major: sensor s owns sub-model vertices [s*slots, (s+1)*slots) (slots = 8 or 12): its
 * sensor vertex, then the neighbours in the winding order of its incident faces, then zero padding.  "sub.fan_dims" =
 * {fan_ok, slots, max valence, number of (sensor, joint) pairs}; fan_ok = 1 (every closed manifold mesh, e.g. SMPL-H)
 * selects the register-resident fan kernel (csrc/fan_kernel.cu), which reads "sub.fan_helper", "sub.fan_n_joints",
 * "sub.fan_part_ptr", "sub.fan_joint", "sub.fan_weight", "sub.fan_jp_ptr", "sub.fan_jp_idx"; otherwise the general,
 * index-table driven kernel (csrc/frame_kernels.cu) runs on "sub.faces", "sub.sensor_faces", "sub.skin_*", ... */
typedef struct {
    const char* name;
    const void* data;        /* HOST pointer, contiguous row-major */
    int32_t dtype;           /* EMPOSE_F32 / EMPOSE_I32 / EMPOSE_I64 */
    int32_t ndim;
    int64_t shape[4];
} empose_tensor;

/* The flags of empose/helpers/configuration.py:150-209 that the path reads. */
typedef struct {
    int32_t n_markers;          /* 6 or 12 (models.py:385) */
    int32_t num_iterations;     /* m_num_iterations */
    float step_size;            /* m_step_size */
    int32_t rnn_init;           /* m_rnn_init */
    int32_t average_shape;      /* m_average_shape */
    int32_t use_gradient;       /* m_use_gradient */
    int32_t use_marker_pos;
    int32_t use_marker_ori;
    int32_t hidden_size;        /* m_hidden_size */
    int32_t num_layers;         /* m_num_layers: LinearLayers blocks per MLP */
    int32_t rnn_hidden_size;    /* m_rnn_hidden_size */
    int32_t rnn_num_layers;     /* m_rnn_num_layers */
    int32_t skip_connections;   /* m_skip_connections */
    int32_t batch_norm;         /* !m_no_batch_norm */
    int32_t precision;          /* EMPOSE_PRECISION_* */
    int32_t device;             /* CUDA device ordinal */
} empose_ief_config;

typedef struct empose_ief empose_ief;   /* opaque: packed weights + sub-model + cached execution plans */

/* Optional per-iterate outputs (the five *_history lists of models.py:620-629).  Any pointer may be
 * NULL.  Each is [N+1][B][F][dof] with dof = 66, 10, 66, 36, 108. */
typedef struct {
    float* pose;
    float* shape;
    float* joints;
    float* markers;
    float* markers_ori;
} empose_ief_history;

int empose_abi_version(void);
const char* empose_last_error(void);

/* Development switches (process-wide, never needed in production; the tests use them for A/B comparisons of two
 * implementations of the same arithmetic).  Keys: "main_general" (1: run the general sub-model kernel even when the
 * sub-model is in fan form), "fan_variant" (frames per CTA / CTAs per SM variant of the fan kernel), "lstm_persistent"
 * (0: one launch per wavefront diagonal instead of the persistent wavefront kernel).  Returns EMPOSE_E_ARG for an
 * unknown key. */
int empose_set_option(const char* key, int32_t value);

/* Build a model context: folds BatchNorm (eval statistics) into the Linear layers in double
 * precision, packs and (TF32 mode) rounds the weights, uploads the sub-model.  Inference semantics
 * (nn.Module.eval()).  Host pointers in `tensors` are only read during the call. */
int empose_ief_create(const empose_ief_config* cfg, const empose_tensor* tensors, int32_t n_tensors,
                      empose_ief** out);
void empose_ief_destroy(empose_ief* ctx);

/* One pass of the hot path over a batch of B windows of F frames.
 *   marker_pos  [B][F][36]   marker_oris [B][F][108]   (always 12 sensors, as batch.get_inputs() supplies)
 *   offset_r    [B][12][9]   offset_t    [B][12][3]
 *   seq_lengths [B] int32 (1..F)      marker_masks [B][F][12] float (non-zero = present) or NULL
 *   lstm_state  [2][rnn_num_layers][B][H] (h then c), in/out, or NULL.  Read unless is_new_sequence;
 *               always written when non-NULL (RNNLayer.final_state, models.py:489-492).
 * Outputs (any may be NULL): pose_hat [B][F][66] (root first: callers slice [3:] / [:3] as models.py:606-607),
 *   shape_hat [B][F][10], joints_hat [B][F][66]. */
int empose_ief_forward(empose_ief* ctx, const float* marker_pos, const float* marker_oris, const float* offset_r,
                       const float* offset_t, const int32_t* seq_lengths, const float* marker_masks,
                       float* lstm_state, int32_t is_new_sequence, int32_t B, int32_t F, float* pose_hat,
                       float* shape_hat, float* joints_hat, const empose_ief_history* history, void* stream);

/* Same call with HOST buffers (pinned or pageable): uploads the inputs, runs the pass and downloads
 * the outputs on `stream`, then synchronises the stream.  lstm_state / history are host pointers too. */
int empose_ief_forward_host(empose_ief* ctx, const float* marker_pos, const float* marker_oris,
                            const float* offset_r, const float* offset_t, const int32_t* seq_lengths,
                            const float* marker_masks, float* lstm_state, int32_t is_new_sequence, int32_t B,
                            int32_t F, float* pose_hat, float* shape_hat, float* joints_hat,
                            const empose_ief_history* history, void* stream);

/* Streaming form of empose_ief_forward_host (the loop of scripts/evaluate_real.py:39-61 -- chunk.to_gpu(), net(chunk), .cpu() --
 * with the copies of one chunk hidden behind the pass of another): enqueue one request into the in-flight slot `slot`
 * (0..3) and return without waiting.  The upload runs on an internal stream (not ordered behind `stream`), the pass on
 * `stream`, the download on a third stream; each slot has its own device workspace.  The host buffers (pinned, or the
 * copies are not asynchronous) must stay untouched until empose_ief_wait_host(ctx, slot) has returned, and a slot must be
 * waited for before it is submitted into again.  Arguments as for empose_ief_forward_host. */
int empose_ief_submit_host(empose_ief* ctx, const float* marker_pos, const float* marker_oris,
                           const float* offset_r, const float* offset_t, const int32_t* seq_lengths,
                           const float* marker_masks, float* lstm_state, int32_t is_new_sequence, int32_t B,
                           int32_t F, float* pose_hat, float* shape_hat, float* joints_hat,
                           const empose_ief_history* history, int32_t slot, void* stream);
/* Blocks the calling host thread until the request submitted into `slot` has written its results to the host buffers. */
int empose_ief_wait_host(empose_ief* ctx, int32_t slot);

/* SMPL-H sub-model -> 12 sensor frames with offsets applied -> first 22 joints, for R frames.
 *   poses [R][66], shapes [R][10], offset_r [R][12][9], offset_t [R][12][3]
 *   -> sensor_pos [R][36], sensor_ori [R][108], joints [R][66] (any output may be NULL). */
int empose_sensor_project(empose_ief* ctx, const float* poses, const float* shapes, const float* offset_r,
                          const float* offset_t, int32_t R, float* sensor_pos, float* sensor_ori, float* joints,
                          void* stream);

/* A context that holds ONLY the SMPL sub-model ("sub.*" arrays), for empose_sensor_project without a learned model:
 * the device side of the reference's training-data synthesis, SMPLFK + SampleMarkersWithOffsets
 * (empose/data/transforms.py:163-226, 259-282), i.e. ground-truth SMPL evaluation -> 12 sensor frames -> offsets in
 * one pass that never materialises the 6890-vertex mesh.  Destroy with empose_ief_destroy; the forward entry points
 * refuse such a context.  `precision`: EMPOSE_PRECISION_FP32, or a tensor-core mode (pose blend 3xTF32). */
int empose_sensors_create(const empose_tensor* tensors, int32_t n_tensors, int32_t precision, int32_t device, empose_ief** out);

/* Evaluation metrics of MetricsEngine.compute (empose/eval/metrics.py:183-241) for R frames in one kernel: FK of ground
 * truth and prediction (22 body joints; `ctx` provides the sub-model: any empose_ief, e.g. from empose_sensors_create),
 * per-joint Euclidean distance eucl [R][22] (metres), the same after per-frame Procrustes alignment with optimal scale
 * eucl_pa [R][22] (metrics.py:19-66) and the geodesic angle angle_deg [R][21] (degrees; may be NULL) between the joint
 * orientations: global ones obtained with a zero root (angle_local = 0: MetricsEngine.angle_glob, metrics.py:229-238) or
 * the local joint rotations themselves (angle_local = 1, metrics.py:239-240).  pose / pose_hat [R][66] (root first),
 * shape / shape_hat [R][10]. */
int empose_metrics_compute(empose_ief* ctx, const float* pose, const float* shape, const float* pose_hat, const float* shape_hat,
                           int32_t R, int32_t angle_local, float* eucl, float* eucl_pa, float* angle_deg, void* stream);
/* Same distances from given joints [R][66] (MetricsEngine.compute_joint_dist, metrics.py:243-265). */
int empose_metrics_joints(const float* joints, const float* joints_hat, int32_t R, float* eucl, float* eucl_pa, void* stream);

/* Number of kernels the last empose_ief_forward* call on this context launched (for bench accounting). */
int64_t empose_ief_last_launch_count(const empose_ief* ctx);

/* ---- full-mesh SMPL-H layer -------------------------------------------------------------------------------- */
typedef struct empose_smpl empose_smpl;   /* opaque: full-mesh constants (all vertices, 52 joints) on one device */

/* `tensors`: the "smpl.*" arrays of submodel.extract_fullmodel (host pointers, read during the call only). */
int empose_smpl_create(const empose_tensor* tensors, int32_t n_tensors, int32_t precision, int32_t device,
                       empose_smpl** out);
void empose_smpl_destroy(empose_smpl* ctx);

/* SMPLLayer._fk (smpl.py:81-122) with a zero hand pose: poses_body [N][63], betas [N][10], poses_root [N][3] or NULL
 * (zeros), trans [N][3] or NULL (zeros) -> verts [N][V][3], joints [N][52][3] (either output may be NULL).
 * Device pointers; evaluated in slabs internally, so N is unbounded. */
int empose_smpl_forward(empose_smpl* ctx, const float* poses_root, const float* poses_body, const float* betas,
                        const float* trans, int32_t N, float* verts, float* joints, void* stream);

/* ---- training step ------------------------------------------------------------------------------------------
 * Parameters live in ONE flat float32 device vector, their gradients in a second one of the same layout and the
 * BatchNorm running statistics in a third; the caller owns all three (PyTorch tensors whose views are the module's
 * nn.Parameters), so an optimiser updates the weights in place and data-parallel training needs a single
 * all-reduce over `grads`.  The layout is fixed by the configuration alone. */
typedef struct empose_train empose_train;

typedef struct {
    float pose_weight;            /* m_pose_loss_weight         (models.py:58)  */
    float shape_weight;           /* m_shape_loss_weight        (models.py:57)  */
    float reprojection_weight;    /* m_reprojection_loss_weight (models.py:376) */
    float fk_weight;              /* m_fk_loss                  (models.py:51)  */
} empose_loss_weights;

/* Entry `index` of the layout: state-dict key, kind (0 = parameter -> params/grads, 1 = running statistic ->
 * bn_buffers), offset and size in floats.  Returns EMPOSE_E_MISSING one past the last entry. */
int empose_train_layout(const empose_ief_config* cfg, int32_t index, char* name_out, int32_t name_cap, int32_t* kind,
                        int64_t* offset, int64_t* numel);
int empose_train_sizes(const empose_ief_config* cfg, int64_t* n_params, int64_t* n_buffers);

/* `tensors`: as for empose_ief_create (state dict + "sub.*").  params / grads / bn_buffers are DEVICE pointers that
 * stay valid for the lifetime of the context (bn_buffers may be NULL when batch_norm = 0).  Dropout is not
 * applied (the released configurations train with p = 0, configuration.py:168,176). */
int empose_train_create(const empose_ief_config* cfg, const empose_tensor* tensors, int32_t n_tensors, float* params,
                        float* grads, float* bn_buffers, empose_train** out);
void empose_train_destroy(empose_train* ctx);

/* Train-mode forward pass from a zero LSTM state (scripts/train.py:146): same inputs / outputs as
 * empose_ief_forward; BatchNorm uses batch statistics and updates bn_buffers; everything the backward pass
 * needs is kept inside the context. */
int empose_train_forward(empose_train* ctx, const float* marker_pos, const float* marker_oris, const float* offset_r,
                         const float* offset_t, const int32_t* seq_lengths, const float* marker_masks, int32_t B, int32_t F,
                         float* pose_hat, float* shape_hat, float* joints_hat, const empose_ief_history* history,
                         void* stream);

/* Losses and parameter gradients of the preceding forward pass.  poses_gt [B][F][66] (root first), shapes_gt [B][10],
 * joints_gt [B][F][66] or NULL (required when fk_weight > 0).  ADDS to `grads` what the reference leaves in .grad
 * after forward + backward: d(total_loss) plus, when use_gradient is set, the N reconstruction-energy gradients of
 * the forward pass (models.py:576).  loss_vals: HOST float[5] = pose, shape, reconstruction, fk, total_loss
 * (models.py:676-680); filling it synchronises the stream (the reference does, too: five .cpu().item() calls).
 * Data-parallel training overlaps the gradient all-reduce with the tail of this pass: with loss_vals = NULL the call only
 * ENQUEUES work, and `dense_ready_event` (a cudaEvent_t, may be NULL) is recorded on `stream` as soon as every gradient
 * except the LSTM's is final -- before the LSTM's backward-through-time runs -- so the caller can reduce that bucket
 * (empose_train_layout: the LSTM tensors come first in the flat vector) on another stream meanwhile ("issued
 * per bucket ... LSTM last").  empose_train_loss_values then synchronises and returns the five numbers. */
int empose_train_backward(empose_train* ctx, const float* poses_gt, const float* shapes_gt, const float* joints_gt,
                          const empose_loss_weights* weights, float* loss_vals, void* dense_ready_event, void* stream);
int empose_train_loss_values(empose_train* ctx, float* loss_vals);

/* SyncBatchNorm for data-parallel training (with per-rank statistics -- the DDP default, and what this library does
 * unless told otherwise -- a run on N ranks is not the run of one device on the global batch; the BatchNorm sites are
 * empose/nn/layers.py:26,57).  `fn` must all-reduce (sum) `count` doubles at the DEVICE pointer `buf` over all ranks,
 * ordered on `stream` (e.g. an NCCL call enqueued on it), and return 0.  The library calls it once per BatchNorm evaluation
 * in the forward pass (sum and sum of squares per unit) and once per BatchNorm site in the backward pass (sums of dy and
 * dy * xhat), between its own kernels; means, variances and the running statistics are then those of the global batch,
 * exactly as torch.nn.SyncBatchNorm under DDP: parameter gradients come from the local sums and are averaged with the
 * flat gradient afterwards.  Every rank must run the same number of rows per step; `world_size` = number of ranks.
 * fn = NULL restores per-rank statistics. */
typedef int (*empose_allreduce_fn)(void* user, double* buf, int64_t count, void* stream);
int empose_train_set_sync_batchnorm(empose_train* ctx, empose_allreduce_fn fn, void* user, int32_t world_size);
int64_t empose_train_last_launch_count(const empose_train* ctx);

/* ---- (Bi)RNN baseline (BASELINE config 4) ------------------------------------------------------- */
typedef struct {
    int32_t n_markers;          /* 6 or 12 */
    int32_t hidden_size;        /* m_hidden_size (1024 in the released BiRNNs, configuration.py:165) */
    int32_t num_layers;         /* m_num_layers */
    int32_t bidirectional;      /* m_bidirectional */
    int32_t learn_init_state;   /* m_learn_init_state: must be 0 (not supported) */
    int32_t estimate_shape;     /* m_estimate_shape: the BatchNorm-free to_shape MLP (models.py:276-280) */
    int32_t shape_hidden_size;  /* m_shape_hidden_size */
    int32_t average_shape;      /* m_average_shape */
    int32_t do_fk;              /* m_fk_loss > 0: also return the 22 joints (models.py:134-144) */
    int32_t use_marker_pos;
    int32_t use_marker_ori;
    int32_t precision;          /* EMPOSE_PRECISION_* */
    int32_t device;
} empose_rnn_config;

typedef struct empose_rnn empose_rnn;

/* `tensors`: the reference's state dict ("rnn.lstm.weight_ih_l0[_reverse]", "to_pose.*", "to_shape.*") plus, when
 * do_fk is set, the "sub.*" arrays of the SMPL sub-model. */
int empose_rnn_create(const empose_rnn_config* cfg, const empose_tensor* tensors, int32_t n_tensors, empose_rnn** out);
void empose_rnn_destroy(empose_rnn* ctx);

/* SimpleRNN.forward over B sequences of F frames (device pointers, stream-ordered):
 *   marker_pos [B][F][36], marker_oris [B][F][108], seq_lengths [B] int32
 *   lstm_state [2][num_layers * directions][B][H] (h then c; index layer * directions + direction), in/out or NULL:
 *              read unless is_new_sequence, always written (RNNLayer.final_state, models.py:292-294)
 *   pose_hat [B][F][66] (root first); shape_hat [B][F][10] and joints_hat [B][F][66] are written when the
 *   configuration produces them (estimate_shape / do_fk) and the pointer is non-NULL. */
int empose_rnn_forward(empose_rnn* ctx, const float* marker_pos, const float* marker_oris, const int32_t* seq_lengths,
                       float* lstm_state, int32_t is_new_sequence, int32_t B, int32_t F, float* pose_hat, float* shape_hat,
                       float* joints_hat, void* stream);
int64_t empose_rnn_last_launch_count(const empose_rnn* ctx);

/* Optional timing of the two kernels that make up the step: while enabled, every launch of the tensor-core GEMM
 * executor and of the per-frame sub-model kernel is bracketed by CUDA events on its stream (tensor-core modes only).
 * empose_ief_profile_read / _read_main wait for them, return the summed device time and the number of launches
 * since the last read / enable, and reset their counters. */
int empose_ief_set_profiling(empose_ief* ctx, int32_t enable);
int empose_ief_profile_read(empose_ief* ctx, double* gemm_ms, int64_t* gemm_launches);
int empose_ief_profile_read_main(empose_ief* ctx, double* main_ms, int64_t* main_launches);

/* Engine self-test: C[M][N] = A[M][K] . W[N][K]^T + bias through the same job executor the model uses
 * (precision selects tcgen05 or FFMA).  Device pointers; lda / ldw / ldc in floats. */
int empose_gemm_selftest(int32_t precision, const float* A, int64_t lda, const float* W, int64_t ldw,
                         const float* bias, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, void* stream);

/* Same engine, event-timed: one warm-up launch, then `reps` launches; *ms_per_launch receives the mean device time. */
int empose_gemm_bench(int32_t precision, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                      float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t reps, float* ms_per_launch,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMPOSE_B200_H */
