/*
 * qa_b200.h -- C ABI of libqa_b200.so, the B200-native (sm_100a) hot path of
 * NJU-RLC/quadrupedal-agility.
 *
 * The reference has no FFI of its own (it is pure Python over PyTorch eager ops); the
 * drop-in boundary is the Python class API.  This header is the C
 * ABI that sits directly under those classes: every entry point replaces one reference
 * method (cited as file:line under /root/reference), takes a POD struct of raw DEVICE
 * pointers + sizes + scalar config, enqueues its kernel(s) on the caller's CUDA stream and
 * returns without synchronising.
 *
 * Conventions
 *   - PyTorch (or any caller) owns every buffer; the library never allocates or frees
 *     device memory and keeps no global state.
 *   - return value: 0 on success, a positive cudaError_t if a launch failed, a negative
 *     QA_E* code if argument validation failed.  The Python wrappers raise RuntimeError.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - float = IEEE fp32; masks are uint8 (torch.bool / torch.uint8 storage); all arrays
 *     are contiguous unless a pitch/stride argument says otherwise.
 *   - one host thread per process / GPU (matches the reference's single-threaded loop).
 */
#ifndef QA_B200_H_
#define QA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QA_ABI_VERSION 1

#define QA_EINVAL (-1)   /* null pointer / bad size */
#define QA_ERANGE (-2)   /* dimension outside the compiled limits */

#define QA_NUM_DOF 12
#define QA_NUM_PROP 57
#define QA_OBS_WIDTH 671       /* 57+4+29+570+11, go2_locomotion_config.py:12-17 */
#define QA_HIST_LEN 10
#define QA_NUM_OBS_DISC 49
#define QA_DIM_C 5
#define QA_NUM_COMMANDS 5
#define QA_NUM_REWARDS 14
#define QA_EPSUM_PITCH 16
#define QA_ACT_HIST_LEN 8
#define QA_MOCAP_W 49
#define QA_MAX_BODIES 32
#define QA_MAX_NOISE_LANES 64

int qa_version(void);
/* human readable build string ("sm_100a, nvcc 12.9, ...") */
const char* qa_build_info(void);
/* sizeof() of argument struct number `which` (order of declaration in this header, QaActionPushArgs
 * = 0 ... QaGaeArgs = 9, QaGatherArgs = 10, QaClipAdamArgs = 11, QaLinearArgs = 12, QaActBwdArgs = 13, QaPpoLossArgs = 14, QaLinearBwdArgs = 15, QaHistEncArgs = 16, QaRowLossArgs = 17, QaPpoScalarsArgs = 18, QaDepthArgs = 19, QaPpoLossTscArgs = 20, QaTscConst = 21, QaTscStepArgs = 22, QaDiscInputArgs = 23, QaDiscRewardArgs = 24, QaHeadFwdArgs = 25, QaHeadBwdArgs = 26, QaPolicySampleArgs = 27, QaDiscPrepareArgs = 28, QaDiscHeadsArgs = 29, QaDiscGpArgs = 30, QaDiscRegArgs = 31, QaNormMomentsArgs = 32, QaNormMergeArgs = 33, QaPeerAllreduceArgs = 34, QaAdamChainArgs = 35; -1 if unknown): a layout handshake for FFI mirrors of these structs */
int qa_struct_size(int which);
/* stream-ordered fill-with-zero / device-to-device copy of `bytes` bytes (cudaMemsetAsync / cudaMemcpyAsync): lets a captured
 * training step zero its flat gradient buffer (optimizer.zero_grad(), gail.py:361, :409) and move device scalars without a
 * framework kernel in between */
int qa_zero_async(void* dst, uint64_t bytes, void* stream);
int qa_copy_async(void* dst, const void* src, uint64_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K0  action history push + delayed-action select + clip
 *     replaces bbc/legged_gym/envs/base/legged_robot.py:84-98 (LeggedRobot.step, front half)
 * ------------------------------------------------------------------------------------------ */
typedef struct QaActionPushArgs {
    int32_t num_envs;
    int32_t delay;                  /* self.delay, 0..7 */
    float clip;                     /* clip_actions / action_scale */
    const float* actions_in;        /* (N,12) policy output */
    float* action_history_buf;      /* (N,8,12) in/out, shifted in place */
    float* actions_out;             /* (N,12) delayed + clipped action  (self.actions) */
} QaActionPushArgs;
int qa_action_push(const QaActionPushArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1  PD torques  -- replaces LeggedRobot._compute_torques, legged_robot.py:547-579
 *     (control_type 'P', randomize_motor=True).  Called `decimation` (4) times per env step.
 * ------------------------------------------------------------------------------------------ */
typedef struct QaTorqueArgs {
    int32_t num_envs;
    float action_scale;             /* 0.25 */
    float hip_scale_reduction;      /* 0.5, applied to DOF 0,3,6,9 */
    const float* actions;           /* (N,12) */
    const float* dof_state;         /* (N,12,2) interleaved pos,vel (IsaacGym layout) */
    const float* motor_strength;    /* (2,N,12) */
    const float* p_gains;           /* (12) */
    const float* d_gains;           /* (12) */
    const float* default_dof_pos;   /* (12) */
    const float* torque_limits;     /* (12) */
    float* torques;                 /* (N,12) clipped, goes to set_dof_actuation_force_tensor */
    float* torques_org;             /* (N,12) un-clipped (self.torques_org) */
} QaTorqueArgs;
int qa_pd_torques(const QaTorqueArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  terrain height scan -- replaces LeggedRobot._get_heights, legged_robot.py:1190-1228
 * ------------------------------------------------------------------------------------------ */
typedef struct QaTerrain {
    const int16_t* height_samples;  /* (rows, cols) */
    int32_t rows, cols;
    float border_size;              /* 30 */
    float horizontal_scale;         /* 0.1 */
    float vertical_scale;           /* 0.005 */
} QaTerrain;

typedef struct QaHeightScanArgs {
    int32_t num_envs;
    int32_t num_points;             /* 187 */
    const float* root_states;       /* (N,13) */
    const float* height_points;     /* (P,3) base-frame sample points (identical for all envs) */
    QaTerrain terrain;
    float* measured_heights;        /* (N,P) */
} QaHeightScanArgs;
int qa_height_scan(const QaHeightScanArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mocap clip table (reference-state initialisation at reset)
 *   MotionLoader.get_full_frame_batch -> traj_time_sample_batch -> get_full_frame_at_time_batch,
 *   bbc/rsl_rl/datasets/motion_loader.py:461-474, 333-341, 410-447; quaternion_slerp,
 *   bbc/rsl_rl/utils/utils.py:126-159
 * ------------------------------------------------------------------------------------------ */
typedef struct QaMocapTable {
    const float* frames;            /* (F,49) */
    const int32_t* clip_start;      /* (K) */
    const double* clip_nframes;     /* (K) */
    const double* clip_len_s;       /* (K) */
    const double* clip_frame_dur;   /* (K) */
    const int32_t* mode_offset;     /* (DIM_C+1) CSR of clips per behaviour mode */
    const int32_t* mode_clips;      /* (K) */
    const double* mode_cdf;         /* (K) inclusive CDF of the within-mode clip weights */
    int32_t num_clips;
    int32_t num_frames;
} QaMocapTable;

/* K4 standalone blend (also used inside K2's reset path) */
typedef struct QaMocapBlendArgs {
    int32_t num;                    /* rows to produce */
    QaMocapTable table;
    const int32_t* clip_idx;        /* (num) */
    const double* time_u;           /* (num) uniform draws in [0,1) */
    double time_between_frames;     /* env.dt */
    int32_t disc_obs_len;           /* 2 */
    float* frames_out;              /* (num,49) */
} QaMocapBlendArgs;
int qa_mocap_blend(const QaMocapBlendArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  fused post-physics step: counters, base-frame quantities, contact filter, periodic
 *     command/latent resampling, centre terrain height, push, termination, 14 reward terms,
 *     reset (resample + mocap frame blend + state write), observations (671 / 49 / history),
 *     last_* carries, reset statistics.
 *     replaces LeggedRobot.post_physics_step and everything it calls:
 *       post_physics_step :124-166, _post_physics_step_callback :449-472, _resample_* :474-540,
 *       _push_robots :682-687, check_termination :168-176, compute_reward :242-259 with the
 *       active _reward_* :1248-1335, reset_idx :178-240 (+ _reset_dofs_mocap :598-612,
 *       _reset_root_states_mocap :660-680), compute_observations :261-331,
 *       compute_flat_key_pos :1377-1396.
 *     One warp per env; see DESIGN.md for the data layout and the byte budget.
 * ------------------------------------------------------------------------------------------ */
typedef struct QaBbcConst {
    /* body roles */
    int32_t num_bodies;
    int32_t feet_indices[4];
    uint32_t termination_body_mask;     /* bit b set: body b terminates the episode on contact */
    uint32_t penalised_body_mask;       /* bit b set: body b counts for _reward_collision */
    /* per-DOF constants */
    float default_dof_pos[12];
    float dof_pos_lower[12];            /* soft limits */
    float dof_pos_upper[12];
    float dof_vel_limits[12];
    float torque_limits[12];
    uint32_t hip_dof_mask;              /* bit d set: DOF d is a hip joint (0,3,6,9) */
    /* reward scales (already multiplied by dt), dir() order, legged_robot.py:922-946 */
    float reward_scale[QA_NUM_REWARDS];
    float dt;                           /* 0.02 */
    float tracking_sigma;
    float soft_dof_vel_limit;
    float soft_torque_limit;
    float jump_goal;                    /* 10 */
    float jump_height_lo;               /* command_ranges["jump_height"][0] = 0.45 */
    int32_t only_positive_rewards;
    /* episode bookkeeping */
    float max_episode_length;           /* 1000 */
    int32_t resample_period;            /* 300 */
    float episode_length_s;             /* 20 */
    /* command ranges per mode */
    float lin_vel_x[QA_DIM_C][2];
    float lin_vel_y[QA_DIM_C][2];
    float ang_vel_yaw[QA_DIM_C][2];
    float jump_h_lo, jump_h_span;       /* lo and (float)(hi-lo) with hi-lo evaluated in double */
    float loco_h_lo, loco_h_span;
    float lin_vel_x_clip, lin_vel_y_clip, ang_vel_yaw_clip;
    float prior_cdf[QA_DIM_C];          /* softmax(prior/T) inclusive CDF, used in Philox mode */
    /* observation scales */
    float s_lin_vel, s_ang_vel, s_dof_pos, s_dof_vel, s_key_pos, s_foot_contact;
    float s_lin_vel_dist, s_ang_vel_dist;
    float clip_obs;
    int32_t add_noise;
    int32_t root_height_obs;
    int32_t measure_heights;
    /* centre height sample point (base frame), legged_robot.py:266 */
    float center_px, center_py;
    float max_push_vel_xy;
    double time_between_frames;         /* env.dt as double (mocap time sampling) */
    int32_t disc_obs_len;
    /* compact form of noise_scale_vec (legged_robot.py:721-740): the lanes of the 671-row with a non-zero
     * scale (32 in the shipped config), so that noise costs one pass over <= 64 lanes, not 671 */
    int32_t num_noise;
    int32_t noise_idx[QA_MAX_NOISE_LANES];
    float noise_scale[QA_MAX_NOISE_LANES];
} QaBbcConst;

#define QA_K2_BULK_STORE 1u   /* obs/priv rows leave through TMA bulk stores (needs pitch 671, N % 4 == 0) */
#define QA_K2_PDL 4u          /* launch the tiled kernel with programmatic stream serialization (back-to-back steps overlap
                               * launch latency with the predecessor's tail; results are identical) */
#define QA_K2_TILED 2u        /* 8-env CTA tiles, every per-env array staged by TMA bulk copies (needs N % 8 == 0,
                               * 16 B aligned bases, pitch 671); falls back to the warp-per-env kernel otherwise */

typedef struct QaBbcStepArgs {
    int32_t num_envs;
    int32_t do_push;                    /* common_step_counter % push_interval == 0 */
    int32_t obs_pitch;                  /* floats between obs rows (>= 671) */
    int32_t contact_ring_head;          /* slot of the contact rings written this step */
    int32_t contact_ring_len;           /* 100 */
    uint32_t flags;                     /* QA_K2_* */
    uint64_t rng_seed;                  /* Philox key   (perf mode) */
    uint64_t rng_step;                  /* Philox counter high word: global step index */

    /* simulator-owned state (IsaacGym tensors), updated in place on reset / push */
    float* root_states;                 /* (N,13) */
    float* dof_state;                   /* (N,12,2) */
    const float* rigid_body_state;      /* (N,B,13) */
    const float* contact_forces;        /* (N,B,3) */

    /* per-env constants */
    const float* motor_strength;        /* (2,N,12) */
    const float* mass_params;           /* (N,4) */
    const float* friction_coeffs;       /* (N,1) */
    const float* env_origins;           /* (N,3) */
    const float* noise_scale_vec;       /* (671) */
    QaTerrain terrain;
    QaMocapTable mocap;

    /* carried env buffers, in/out */
    int64_t* episode_length_buf;        /* (N) */
    uint8_t* last_contacts;             /* (N,4) */
    float* commands;                    /* (N,5) */
    float* latent_eps;                  /* (N,1) */
    float* latent_c;                    /* (N,5) */
    const float* actions;               /* (N,12) */
    float* last_actions;                /* (N,12) */
    const float* torques_org;           /* (N,12) */
    float* last_torques_org;            /* (N,12) */
    float* last_dof_vel;                /* (N,12) */
    float* last_root_vel;               /* (N,6) */
    float* action_history_buf;          /* (N,8,12) */
    float* obs_history_buf;             /* (N,10,57) */
    float* episode_sums;                /* (N,16) env-major, first 14 used */
    float* feet_air_time;               /* (N,4) */
    float* contact_buf;                 /* (N,L,4) ring, may be NULL */
    float* contact_force_buf;           /* (N,L,4) ring, may be NULL */

    /* step outputs */
    float* obs_buf;                     /* (N,pitch) */
    float* privileged_obs_buf;          /* (N,pitch); may alias obs_buf (byte-identical content) */
    float* obs_disc_buf;                /* (N,49) this step's; the previous one stays intact for terminal states */
    float* rew_buf;                     /* (N) */
    uint8_t* reset_buf;                 /* (N) */
    uint8_t* time_out_buf;              /* (N) */
    float* base_lin_vel;                /* (N,3) */
    float* base_ang_vel;                /* (N,3) */
    float* projected_gravity;           /* (N,3) */
    float* rpy;                         /* (N,3) roll,pitch,yaw */
    float* feet_forces;                 /* (N,4) */
    uint8_t* contact_filt;              /* (N,4) */
    float* root_h;                      /* (N) z - centre terrain height, pre-reset (rewards) */

    /* reset statistics, written by the last CTA to finish */
    float* episode_rew_means;           /* (14) mean over reset envs / episode_length_s; untouched if no reset */
    uint8_t* time_outs_latched;         /* (N) extras["time_outs"]: refreshed only on steps with >=1 reset (:239-240) */
    int32_t* num_resets;                /* (1) */
    void* workspace;                    /* >= 128 bytes, zero-initialised once by the caller */
    /* optional device-resident step counter (CUDA-graph replay): when non-NULL, step_state[0] is
     * common_step_counter BEFORE this step; the kernel derives rng_step = counter + 1, do_push =
     * (push_interval > 0 && (counter + 1) % push_interval == 0), contact_ring_head = counter % ring_len
     * from it (the scalar fields above are ignored) and its last CTA stores counter + 1 back */
    int64_t* step_state;
    int32_t push_interval;              /* ceil(push_interval_s / dt) = 400, or 0 when push_robots is off */

    /* parity-mode random draws (dense, one per env).  ALL NULL => in-kernel Philox4x32-10 */
    const float* noise_u;               /* (N,671) */
    const double* rs_eps_u;             /* (N) periodic-resample site */
    const int32_t* rs_c_idx;            /* (N) */
    const float* rs_cmd_u;              /* (N,5) */
    const double* rt_eps_u;             /* (N) reset site */
    const int32_t* rt_c_idx;            /* (N) */
    const float* rt_cmd_u;              /* (N,5) */
    const float* push_u;                /* (N,2) */
    const int32_t* mocap_clip_idx;      /* (N) */
    const double* mocap_time_u;         /* (N) */
    /* live skill prior (legged_robot.py:536-538): inclusive CDF of softmax(prior_parameters / T), (QA_DIM_C) floats in
     * DEVICE memory, re-derived by the host wrapper whenever the trainer moves env.prior_parameters (gail.py:462-464);
     * read by the Philox-mode mode draw -- also under CUDA-graph replay.  NULL => QaBbcConst.prior_cdf */
    const float* prior_cdf;
} QaBbcStepArgs;
int qa_post_physics_bbc(const QaBbcConst* c, const QaBbcStepArgs* a, void* stream);

/* sorted compaction of the reset mask (+ gather of the terminal discriminator states)
 * replaces `reset_buf.nonzero()` / `obs_disc_buf[env_ids]`, legged_robot.py:153-154 */
typedef struct QaCompactArgs {
    int32_t num_envs;
    const uint8_t* reset_buf;           /* (N) */
    const float* prev_obs_disc_buf;     /* (N,49) may be NULL */
    int64_t* reset_env_ids;             /* (N) first *count entries valid, ascending */
    int32_t* reset_env_ids_i32;         /* (N) same, int32 for gym.set_*_indexed; may be NULL */
    float* terminal_disc_states;        /* (N,49) first *count rows valid; may be NULL */
    int32_t* count;                     /* (1) */
} QaCompactArgs;
int qa_compact_resets(const QaCompactArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5  GAE -- replaces RolloutStorage.compute_returns, bbc/rsl_rl/storage/rollout_storage.py:97-111
 *     (identical in tsc/rsl_rl/storage/rollout_storage.py:102-116)
 * ------------------------------------------------------------------------------------------ */
typedef struct QaGaeArgs {
    int32_t num_steps;                  /* T = 24 (<= 32: one warp scans the horizon) */
    int32_t num_envs;
    float gamma, lam;
    const float* rewards;               /* (T,N) */
    const float* values;                /* (T,N) */
    const uint8_t* dones;               /* (T,N) */
    const float* last_values;           /* (N) */
    float* returns;                     /* (T,N) */
    float* advantages;                  /* (T,N) normalised: (A-mean)/(std_unbiased+1e-8) */
    double* workspace;                  /* >= 64 bytes, zero-initialised once by the caller */
} QaGaeArgs;
int qa_gae(const QaGaeArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6  minibatch gather -- replaces the 9 index ops per minibatch of RolloutStorage.mini_batch_generator,
 *     bbc/rsl_rl/storage/rollout_storage.py:147-155: dst[t][j,:] = src[t][indices[j],:] for every tensor t
 * ------------------------------------------------------------------------------------------ */
#define QA_GATHER_MAX_TENSORS 16
typedef struct QaGatherArgs {
    int64_t num_rows;                   /* minibatch size (24 576) */
    int32_t num_tensors;
    const int64_t* indices;             /* (num_rows) rows of the flattened (T*N) storage */
    const float* src[QA_GATHER_MAX_TENSORS];   /* (T*N, width[t]) */
    float* dst[QA_GATHER_MAX_TENSORS];         /* (num_rows, width[t]), row pitch dst_pitch[t] */
    int32_t width[QA_GATHER_MAX_TENSORS];
    int32_t dst_pitch[QA_GATHER_MAX_TENSORS];  /* destination row pitch in floats (0 = width[t]); a pitch that is a multiple
                                                  of 4 floats makes the minibatch a legal TMA operand without a copy */
    /* column windows (all 0 = whole rows): entry t copies src[t][row, src_col0 : src_col0+width] (rows src_pitch floats
     * apart, 0 = width) to dst[t][j, dst_col0 : dst_col0+width] -- e.g. the actor's input row [prop 57 | explicit 4 | . 29 . |
     * command 11] is assembled from two windows of the stored observation row (actor_critic.py:171-187) by the gather */
    int32_t src_pitch[QA_GATHER_MAX_TENSORS];
    int32_t src_col0[QA_GATHER_MAX_TENSORS];
    int32_t dst_col0[QA_GATHER_MAX_TENSORS];
} QaGatherArgs;
int qa_gather_minibatch(const QaGatherArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K8  fused gradient clipping + Adam on one flat fp32 parameter buffer -- replaces
 *     nn.utils.clip_grad_norm_(params, max_grad_norm) + optimizer.step(),
 *     bbc/rsl_rl/algorithms/gail.py:409-412 (actor-critic) and :361-365 (estimator).
 *     grads are first multiplied by grad_scale (1/world_size after an NCCL all-reduce(SUM)); the learning
 *     rate and the step counter live on the device so that the adaptive-KL schedule (:368-379) needs no
 *     host round trip.
 * ------------------------------------------------------------------------------------------ */
typedef struct QaClipAdamArgs {
    int64_t numel;
    float* params;                      /* (numel) in/out */
    const float* grads;                 /* (numel) */
    float* exp_avg;                     /* (numel) in/out */
    float* exp_avg_sq;                  /* (numel) in/out */
    const float* lr;                    /* (1) device scalar */
    int32_t* step;                      /* (1) device scalar, incremented by the call */
    float beta1, beta2, eps;
    float max_grad_norm;                /* <= 0: no clipping */
    float grad_scale;
    float* grad_norm_out;               /* (1) total norm before clipping, may be NULL */
    double* workspace;                  /* >= 16 bytes */
    float weight_decay;                 /* torch.optim.Adam(weight_decay=...): grad += weight_decay * param (after scaling /
                                           clipping); 0 for the PPO optimisers, 1e-3 for the discriminator's (gail.py:107-128) */
} QaClipAdamArgs;
int qa_clip_adam(const QaClipAdamArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7  Y = act(X W^T + b) on the tcgen05 tensor cores (TF32 operands, fp32 accumulate in TMEM) -- replaces every
 *     nn.Linear (+ the ELU / ReLU that follows it) of ActorCritic, Estimator and Discriminator:
 *     bbc/rsl_rl/modules/actor_critic.py:96-129, modules/estimator.py:24-33, algorithms/discriminator.py:36-46.
 *     x and w are read by TMA: bases 16 B aligned, pitches multiples of 4 floats; M, N, K are arbitrary.
 * ------------------------------------------------------------------------------------------ */
typedef struct QaLinearArgs {
    int32_t M, N, K;
    int32_t act;                        /* 0 none, 1 ELU(alpha=1), 2 ReLU */
    const float* x;                     /* (M,K) row-major */
    int64_t x_pitch;                    /* floats between rows of x */
    const float* w;                     /* (N,K) row-major: PyTorch Linear.weight */
    int64_t w_pitch;
    const float* bias;                  /* (N) or NULL */
    float* y;                           /* (M,N) */
    int64_t y_pitch;
    /* column windows inside wider rows (0 = the tensor starts at its base): x[:, x_col0 : x_col0+K] is the input and
     * y[:, y_col0 : y_col0+N] the output.  x_col0 must be a multiple of 4 floats (a TMA box starts on a 16-byte boundary);
     * y_col0 may be any column -- an unaligned output window (the actor-input lanes 61..89) leaves through plain stores */
    int32_t x_col0, y_col0;
} QaLinearArgs;
int qa_linear_fwd(const QaLinearArgs* a, void* stream);

/* backward contractions of y = x W^T + b on the same tcgen05 kernel with MN-major operands:
 *   dx  = gz W        (overwritten; skipped when dx == NULL)      -- autograd of actor_critic.py:113-129
 *   dw += gz^T x      (ACCUMULATED with fp32 atomics over a split of the M reduction; skipped when dw == NULL)
 * gz is the gradient w.r.t. the pre-activation (K9 output).  Same TMA constraints as qa_linear_fwd. */
typedef struct QaLinearBwdArgs {
    int32_t M, N, K;                    /* forward shapes: x (M,K), w (N,K), gz (M,N) */
    const float* gz; int64_t gz_pitch;
    const float* x;  int64_t x_pitch;
    const float* w;  int64_t w_pitch;
    float* dx;       int64_t dx_pitch;
    float* dw;       int64_t dw_pitch;
    /* optional fusion with the PREVIOUS layer's activation backward (x is that layer's output y_prev = act(z_prev)):
     * when act_prev != 0 the dx epilogue writes dx * act'(z_prev) -- i.e. the gradient w.r.t. z_prev -- and, if db_prev is
     * given, its column sums (the previous layer's bias gradient, overwritten) */
    int32_t act_prev;                   /* 0 none, 1 ELU, 2 ReLU */
    const float* y_prev; int64_t y_prev_pitch;   /* read through TMA: 16 B aligned base, pitch % 4 == 0 */
    float* db_prev;                     /* (K) or NULL */
    int32_t db_accumulate;              /* 0: db_prev is zeroed first; 1: accumulate (flat gradient buffer zeroed by the caller) */
    int32_t x_col0;                     /* dw: x[:, x_col0 : x_col0+K] is the layer input; multiple of 4 */
    int32_t w_col0;                     /* dx: uses w[:, w_col0 : w_col0+K] (gradient w.r.t. a window of the input row); multiple of 4 */
} QaLinearBwdArgs;
int qa_linear_bwd(const QaLinearBwdArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K9  gz = gy * act'(y), db = column sums of gz -- the element-wise half of Linear+ELU/ReLU backward
 *     (autograd of bbc/rsl_rl/modules/actor_critic.py:113-129).  y is the layer OUTPUT saved by the forward.
 * ------------------------------------------------------------------------------------------ */
typedef struct QaActBwdArgs {
    int64_t M;
    int32_t N;
    int32_t act;                        /* 0 none, 1 ELU, 2 ReLU */
    const float* gy; int64_t gy_pitch;  /* (M,N) upstream gradient */
    const float* y;  int64_t y_pitch;   /* (M,N) forward output (unused when act == 0) */
    float* gz;       int64_t gz_pitch;  /* (M,N) out, may alias gy, may be NULL (bias gradient only) */
    float* db;                          /* (N) out, may be NULL */
    int32_t zero_db;                    /* 1: db is zeroed first, 0: accumulate */
    /* optional second upstream gradient: gz = (gy + *addend_scale * addend) * act'(y) -- the privileged-latent encoder's output
     * receives the actor's input gradient AND the regulariser's (gail.py:352-354, coefficient = a device scalar) */
    const float* addend; int64_t addend_pitch;
    const float* addend_scale;          /* (1) device scalar, NULL = 1 */
} QaActBwdArgs;
int qa_act_bwd(const QaActBwdArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K20 / K21  narrow output layers (N <= 16 columns) on the CUDA cores, full fp32 -- actor_head Linear(128,12) and critic_head
 *     Linear(128,1) (bbc/rsl_rl/modules/actor_critic.py:118-119, 128-129), the estimator's last Linear(64,4)
 *     (modules/estimator.py:24-33), the discriminator's three heads as one (7,256) matrix (algorithms/discriminator.py:42-46).
 *     Kh in {32, 64, 128} (forward also 256); h: 16 B aligned base, pitch % 4 == 0.
 *     qa_head_fwd:  y = h W^T + b
 *     qa_head_bwd:  gz_prev = ((gz_scale * gz) W) * act'(h)  [act = activation that PRODUCED h: 0 none, 1 ELU, 2 ReLU],
 *                   dw += (gz_scale * gz)^T h, db += colsum(gz_scale * gz), db_prev += colsum(gz_prev)   (all ACCUMULATED)
 * ------------------------------------------------------------------------------------------ */
typedef struct QaHeadFwdArgs {
    int64_t M;
    int32_t N, Kh;
    const float* h; int64_t h_pitch;    /* (M,Kh) */
    const float* w; int64_t w_pitch;    /* (N,Kh) */
    const float* bias;                  /* (N) or NULL */
    float* y; int64_t y_pitch;          /* (M,N) */
} QaHeadFwdArgs;
int qa_head_fwd(const QaHeadFwdArgs* a, void* stream);

typedef struct QaHeadBwdArgs {
    int64_t M;
    int32_t N, Kh;
    int32_t act;
    float gz_scale;
    const float* gz; int64_t gz_pitch;  /* (M,N) gradient w.r.t. the head's output */
    const float* h;  int64_t h_pitch;   /* (M,Kh) the head's input = previous layer's output */
    const float* w;  int64_t w_pitch;   /* (N,Kh) */
    float* gz_prev;  int64_t gz_prev_pitch;   /* (M,Kh) out, may be NULL */
    float* dw;       int64_t dw_pitch;  /* (N,Kh) accumulated, may be NULL */
    float* db;                          /* (N) accumulated, may be NULL */
    float* db_prev;                     /* (Kh) accumulated, may be NULL */
} QaHeadBwdArgs;
int qa_head_bwd(const QaHeadBwdArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K22 policy sample -- Normal(mean, std).sample() + log_prob(actions).sum(-1) + the transition's storage writes of
 *     SSInfoGAIL.act, bbc/rsl_rl/algorithms/gail.py:186-196 (modules/actor_critic.py:189-197), one launch.
 *     noise != NULL: actions = mu + std * noise (parity mode, the reference draws from torch's generator);
 *     noise == NULL: standard normals from Philox4x32-10 keyed by (rng_seed; env, site, step) + Box-Muller, where step =
 *     *step_state + 1 when step_state != NULL (device-resident env step counter, CUDA-graph replay) else rng_step.
 * ------------------------------------------------------------------------------------------ */
typedef struct QaPolicySampleArgs {
    int64_t M;
    int32_t A;                          /* action dims (12) */
    const float* mu; int64_t mu_pitch;  /* (M,A) */
    const float* std;                   /* (A) */
    const float* noise;                 /* (M,A) or NULL */
    uint64_t rng_seed, rng_step;
    const int64_t* step_state;          /* (1) or NULL */
    float* actions;                     /* (M,A) out: what env.step() receives */
    float* logp;                        /* (M) out or NULL */
    float* actions_st;                  /* (M,A) storage.actions[t] or NULL */
    float* logp_st;                     /* (M) storage.actions_log_prob[t] or NULL */
    float* mu_st;                       /* (M,A) storage.mu[t] or NULL */
    float* sigma_st;                    /* (M,A) storage.sigma[t] or NULL */
} QaPolicySampleArgs;
int qa_policy_sample(const QaPolicySampleArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K10 PPO loss forward + backward -- replaces the element-wise graph of SSInfoGAIL.update_actor_critic,
 *     bbc/rsl_rl/algorithms/gail.py:367-408 (KL, clipped surrogate, clipped value loss, bound loss, entropy).
 *     loss = c_surr*mean(surr) + c_value*mean(vl) + c_bound*mean(b) - c_entropy*mean(entropy); gradients w.r.t. the
 *     action mean, the value and the std parameter are written, stats = {surrogate, value, bound, kl} means.
 * ------------------------------------------------------------------------------------------ */
#define QA_PPO_STATS 4
typedef struct QaPpoLossArgs {
    int64_t M;
    const float* mu; int64_t mu_pitch;  /* (M,12) action mean */
    const float* std;                   /* (12) */
    const float* value; int64_t value_pitch; /* (M,1) */
    const float* actions;               /* (M,12) */
    const float* old_logp;              /* (M) */
    const float* advantages;            /* (M) */
    const float* returns;               /* (M) */
    const float* target_values;         /* (M) */
    const float* old_mu;                /* (M,12) */
    const float* old_sigma;             /* (M,12) */
    float clip, c_surr, c_value, c_bound, c_entropy;
    int32_t use_clipped_value_loss;
    float* dmu;                         /* (M,12) */
    float* dvalue;                      /* (M) */
    float* dstd;                        /* (12) */
    float* stats;                       /* (QA_PPO_STATS) */
} QaPpoLossArgs;
int qa_ppo_loss(const QaPpoLossArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K11 StateHistoryEncoder forward, fused (tsteps = 10, ELU): Linear(57->30) per step, Conv1d(30->20,k4,s2),
 *     Conv1d(20->10,k2), Flatten, Linear(30->29) -- replaces bbc/rsl_rl/modules/actor_critic.py:51-59
 *     (StateHistoryEncoder.forward) as used by infer_hist_latent (:219-220).
 * ------------------------------------------------------------------------------------------ */
typedef struct QaHistEncArgs {
    int64_t M;
    const float* hist; int64_t hist_pitch;     /* (M, 570) = (M, 10, 57), e.g. obs + 90 with the obs pitch */
    const float* w0; int64_t w0_pitch;         /* encoder.0.weight (30,57) */
    const float* b0;                           /* (30) */
    const float* w1; const float* b1;          /* conv_layers.0 weight (20,30,4) contiguous, bias (20) */
    const float* w2; const float* b2;          /* conv_layers.2 weight (10,20,2) contiguous, bias (10) */
    const float* w3; int64_t w3_pitch;         /* linear_output.0.weight (29,30) */
    const float* b3;                           /* (29) */
    float* out; int64_t out_pitch;             /* (M,29) */
} QaHistEncArgs;
int qa_hist_encoder_fwd(const QaHistEncArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K12 row losses, forward + backward in one pass -- the two auxiliary losses of a PPO minibatch step,
 *     bbc/rsl_rl/algorithms/gail.py:352-365:
 *       mode 0  mean squared error   loss = mean_{i,k} (a[i,k] - b[i,k])^2        (estimator loss, :359)
 *       mode 1  mean row L2 distance loss = mean_i ||a[i,:] - b[i,:]||_2          (priv_reg_loss, :354)
 *     da = d loss / d a is written in the same kernel (b is treated as a constant); *loss is zeroed first.
 * ------------------------------------------------------------------------------------------ */
typedef struct QaRowLossArgs {
    int64_t M;
    int32_t W;                          /* row width, 1..32 */
    int32_t mode;
    const float* a; int64_t a_pitch;    /* (M,W) */
    const float* b; int64_t b_pitch;    /* (M,W) */
    float* da; int64_t da_pitch;        /* (M,W) out */
    float* loss;                        /* (1) out */
} QaRowLossArgs;
int qa_row_loss(const QaRowLossArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K13 per-minibatch scalar bookkeeping on the device (one thread): the adaptive-KL learning-rate rule
 *     (gail.py:368-379: lr /= 1.5 above 2*desired_kl, *= 1.5 below desired_kl/2, clamped to [1e-5, 1e-2]) and the
 *     running sums of the seven logged statistics (:277-282) -- replaces ~25 one-element torch kernels.
 *     stats_accum[0..6] += {surrogate, value, bound, entropy, priv_reg, estimator, kl}.
 * ------------------------------------------------------------------------------------------ */
typedef struct QaPpoScalarsArgs {
    const float* ppo_stats;             /* (4) K10 output: surrogate, value, bound, kl */
    const float* std;                   /* (num_actions) policy std parameter -> entropy */
    int32_t num_actions;
    const float* priv_reg_loss;         /* (1) */
    const float* estimator_loss;        /* (1) */
    const float* kl;                    /* (1) kl used by the schedule (all-reduced across ranks when sharded) */
    float desired_kl;                   /* <= 0: fixed schedule, lr untouched */
    float lr_min, lr_max;
    float* lr;                          /* (1) in/out */
    float* stats_accum;                 /* (7) in/out */
} QaPpoScalarsArgs;
int qa_ppo_scalars(const QaPpoScalarsArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K14 TSC student depth preprocessing, all envs in one launch -- replaces the per-env loop of
 *     LeggedRobot.update_depth_buffer / process_depth_image / crop_depth_image / normalize_depth_image,
 *     tsc/legged_gym/envs/base/legged_robot.py:154-202.
 *     Per env: crop the (in_h,in_w) camera image at (crop_top, crop_left) to (out_h,out_w), clip to [-far,-near],
 *     x = (-x - near)/(far - near) - 0.5, + depth_noise*2*(u2-0.5) + (depth_noise*u1)*2*(u_pixel-0.5); then
 *     depth_buffer[e] = new frame in all L slots if episode_length_buf[e] <= 1, else shift by one and append.
 *     Random sources: dense uniforms (parity mode: all three arrays given) or in-kernel Philox4x32-10 keyed by
 *     (rng_seed; env, site, rng_step).
 * ------------------------------------------------------------------------------------------ */
typedef struct QaDepthArgs {
    int32_t num_envs;
    int32_t in_h, in_w;                 /* 60, 106 (cfg.depth.original is (W,H) = (106,60)) */
    int32_t crop_top, crop_left;        /* 1, 10 */
    int32_t out_h, out_w;               /* 58, 87 */
    int32_t buffer_len;                 /* 2 */
    const float* const* image_ptrs;     /* (N) device array of per-env camera tensors (gym.get_camera_image_gpu_tensor), or NULL */
    const float* images;                /* batched alternative: env e at images + e*image_stride */
    int64_t image_stride;
    const int64_t* episode_length_buf;  /* (N) */
    float near_clip, far_clip, depth_noise;
    float clip_span;                    /* float32(far_clip - near_clip), the difference taken in double like the reference's Python */
    const float* noise_scale_u;         /* (N) u1 or NULL */
    const float* offset_u;              /* (N) u2 or NULL */
    const float* pixel_u;               /* (N,out_h*out_w) or NULL */
    uint64_t rng_seed, rng_step;
    float* depth_buffer;                /* (N,L,out_h,out_w) in/out */
} QaDepthArgs;
int qa_depth_update(const QaDepthArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K15 TSC PPO loss forward + backward -- the element-wise graph of PPO.update, tsc/rsl_rl/algorithms/ppo.py:176-262:
 *     a Categorical over the behaviour modes (probs = softmax(logits), torch's clamped-log semantics) and a Normal over
 *     the continuous actions, one clipped surrogate each, clipped value loss, KL of the Normal part, entropy
 *     = mean_j H(Normal_j) + H(Categorical).
 *     loss = surr_d + surr_c + c_value*value_loss - c_entropy*mean(entropy)   (+ priv_reg, added by the caller;
 *     the reference's bound loss has coefficient 0.0).  Gradients w.r.t. the logits, the action mean, the value
 *     and the std parameter are written; stats = means of {surrogate, value_loss, entropy, kl}.
 * ------------------------------------------------------------------------------------------ */
#define QA_TSC_NUM_MODES 3
#define QA_TSC_NUM_CONT 18
typedef struct QaPpoLossTscArgs {
    int64_t M;
    const float* logits; int64_t logits_pitch;   /* (M,3) mode head output (pre-softmax) */
    const float* mu; int64_t mu_pitch;           /* (M,18) */
    const float* std;                            /* (18) */
    const float* value; int64_t value_pitch;     /* (M,1) */
    const float* actions; int64_t actions_pitch; /* (M,19): [mode index as float | 18 continuous] */
    const float* old_logp_d;                     /* (M) */
    const float* old_logp_c;                     /* (M) */
    const float* advantages;                     /* (M) */
    const float* returns;                        /* (M) */
    const float* target_values;                  /* (M) */
    const float* old_mu;                         /* (M,18) */
    const float* old_sigma;                      /* (M,18) */
    float clip, c_value, c_entropy;
    int32_t use_clipped_value_loss;
    float* dlogits; int64_t dlogits_pitch;       /* (M,3) */
    float* dmu; int64_t dmu_pitch;               /* (M,18) */
    float* dvalue;                               /* (M) */
    float* dstd;                                 /* (18) */
    float* stats;                                /* (4) */
} QaPpoLossTscArgs;
int qa_ppo_loss_tsc(const QaPpoLossTscArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K16 / K17  TSC (agility) post-physics step -- replaces LeggedRobot.post_physics_step,
 *     tsc/legged_gym/envs/base/legged_robot.py:226-298, split where the reference steps physics inside reset_idx
 *     (:381-384), so reward/termination and the observations cannot be one kernel:
 *
 *     qa_post_physics_tsc_pre   :236-270 + the simulator-state half of reset_idx: counters, base-frame quantities, euler
 *         angles, foot contacts, _update_goals (:204-224), the 132-point height scan (:1708-1755), current obstacle type,
 *         check_termination (:322-346), the active reward terms (reach_goal, tracking_goal_vel, tracking_yaw, collision,
 *         action_hl_rate, latent_c_rate, feet_edge, termination; :1779-1930), episode sums, and for envs that reset:
 *         cur_goal_idx = 0, _reset_dofs (:798-838), _reset_root_states (:840-900), episode statistics (:398-405),
 *         episode_length_buf = reach_goal_timer = 0.  The LAST block finalises: episode reward means, time-out latch
 *         (:408-410), reset count, obst_dof_vel[:] = 0 flag.
 *     -- caller: gym.set_*_tensor_indexed(reset ids), gym.simulate, refresh_rigid_body_state_tensor --
 *     qa_post_physics_tsc_post  the buffer half of reset_idx (:386-396), cur/next goal gathers (:271-272),
 *         compute_observations (:432-515: obs 800, obs_bbc 671, obs_disc 49, history fill/shift, contact ring, clip),
 *         last_* copies (:278-281).
 *     Random sources of the reset (yaw / x / y): dense uniforms (parity) or in-kernel Philox4x32-10.
 * ------------------------------------------------------------------------------------------ */
#define QA_TSC_NUM_REWARDS 8            /* 7 terms in dir() order + termination */
#define QA_TSC_OBS 800
#define QA_TSC_SCAN 132
#define QA_TSC_AUX 8
typedef struct QaTscConst {
    int32_t num_bodies;
    int32_t feet_indices[4];
    uint32_t termination_body_mask, penalised_body_mask;
    float dt, max_episode_length, episode_length_s;
    float next_goal_threshold, leave_goal_threshold, reach_goal_delay_steps;   /* reach_goal_delay / dt */
    int32_t num_goals_total, num_goals_per_obstacle, last_goal_repeat, num_obstacle_types;
    int32_t update_interval, use_camera, root_height_obs, only_positive_rewards;
    float target_lin_vel;
    float reward_scale[QA_TSC_NUM_REWARDS];   /* already multiplied by dt; order: action_hl_rate, collision, feet_edge,
                                                 latent_c_rate, reach_goal, tracking_goal_vel, tracking_yaw, termination */
    float default_dof_pos[12];
    float base_init_state[13];
    float rand_yaw_range, rand_x_range, rand_y_range, frame_ang0, seesaw_dof_pos;
    float s_lin_vel, s_ang_vel, s_dof_pos, s_dof_vel, s_key_pos, s_foot_contact, s_lin_vel_dist, s_ang_vel_dist, clip_obs;
    int32_t num_height_points;                /* 132 */
    int32_t hl_hist_len, hl_action_dim;       /* action_hl_history_buf (N, hl_hist_len, hl_action_dim); 0 = None */
    int32_t contact_ring_len;                 /* 100 */
} QaTscConst;

typedef struct QaTscStepArgs {
    int32_t num_envs;
    int64_t global_counter;             /* value DURING this step (after step()'s increment) */
    /* simulator tensors (IsaacGym layouts) */
    float* root_states;                 /* (N,13) in/out */
    float* dof_state;                   /* (N,12,2) in/out */
    const float* rigid_body_state;      /* (N,B,13): pre kernel = before the reset step, post kernel = refreshed */
    const float* contact_forces;        /* (N,B,3) */
    float* obst_dof_state;              /* (num_obst_dofs,2) in/out */
    const int64_t* seesaw_dof_index;    /* (N) row of env e's seesaw in obst_dof_state */
    int64_t num_obst_dofs;
    /* static */
    QaTerrain terrain;                  /* obstacle height field */
    const uint8_t* x_edge_mask;         /* (rows, cols) bool */
    const float* height_points;         /* (N,P,3) */
    const float* env_goals;             /* (N,G,3) */
    const int64_t* obstacle_types;      /* (N, num_obstacle_types) */
    const float* mass_params;           /* (N,4) */
    const float* friction_coeffs;       /* (N,1) */
    const float* motor_strength;        /* (2,N,12) */
    /* carried buffers */
    int64_t* episode_length_buf;        /* (N) */
    const float* last_root_vel_in;      /* (N,6) read by the pre kernel (base_lin_acc) */
    uint8_t* last_contacts;             /* (N,4) */
    float* reach_goal_timer;            /* (N) */
    int64_t* cur_goal_idx;              /* (N) */
    float* cur_goals;                   /* (N,3) */
    float* next_goals;                  /* (N,3) */
    const float* actions;               /* (N,12) */
    const float* torques_org;           /* (N,12) */
    float* last_actions; float* last_dof_vel; float* last_torques_org; float* last_root_vel;   /* (N,12|12|12|6) */
    const float* commands;              /* (N,5) */
    const float* latent_eps;            /* (N,1) */
    const float* latent_c;              /* (N,5) */
    const float* action_hl_history_buf; /* (N,H,A) or NULL */
    float* episode_sums;                /* (N,QA_TSC_NUM_REWARDS) */
    float* feet_air_time;               /* (N,4) */
    float* obs_history_buf;             /* (N,10,57) */
    float* action_history_buf;          /* (N,8,12) */
    float* contact_buf;                 /* (N,L,4) ring, slot head = newest */
    int32_t contact_ring_head;
    float* measured_heights;            /* (N,P) */
    float* delta_yaw; float* delta_next_yaw;            /* (N) */
    /* per-step outputs */
    float* base_lin_vel; float* base_ang_vel; float* projected_gravity; float* base_lin_acc; float* rpy;   /* (N,3) */
    uint8_t* contact_filt;              /* (N,4) */
    float* target_yaw; float* next_target_yaw;          /* (N) */
    int64_t* cur_obstacle_types;        /* (N) */
    uint8_t* reached_goal; uint8_t* reach_goal_cutoff; uint8_t* feet_at_edge;   /* (N) (N) (N,4) */
    uint8_t* reset_buf; uint8_t* time_out_buf; uint8_t* time_outs_latched;      /* (N) */
    float* rew_buf;                     /* (N) */
    float* episode_rew_means;           /* (QA_TSC_NUM_REWARDS) written when >= 1 env reset */
    int32_t* num_resets;                /* (1) */
    void* workspace;                    /* >= 128 B, zero-initialised once */
    float* obs_buf;                     /* (N,800) */
    float* obs_bbc_buf;                 /* (N,671) */
    float* obs_disc_buf;                /* (N,49) */
    /* reset randomness */
    const float* yaw_u; const float* x_u; const float* y_u;   /* (N) each, or all NULL */
    uint64_t rng_seed, rng_step;
} QaTscStepArgs;
int qa_post_physics_tsc_pre(const QaTscConst* c, const QaTscStepArgs* a, void* stream);
int qa_post_physics_tsc_post(const QaTscConst* c, const QaTscStepArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K18 discriminator input assembly -- the runner's disc-history bookkeeping and the front half of
 *     Discriminator.predict_disc_reward in one pass: bbc/rsl_rl/runners/on_policy_runner.py:163-181 (terminal-state
 *     patch, 2-step history roll, restart of the history of reset envs), bbc/rsl_rl/algorithms/discriminator.py:74-88
 *     (task-obs weighting, per-step weight), bbc/rsl_rl/utils/utils.py:97-103 (Normalizer.normalize_torch).
 * ------------------------------------------------------------------------------------------ */
typedef struct QaDiscInputArgs {
    int32_t num_envs;
    const uint8_t* dones;               /* (N) reset mask of this step */
    const float* prev_disc;             /* (N,49) disc obs of the previous step (terminal state of envs that reset) */
    const float* next_disc;             /* (N,49) disc obs of this step */
    const float* hist_prev;             /* (N,2,49) */
    float* hist_new;                    /* (N,2,49) history the reward is computed on / the replay buffer stores */
    float* hist_next;                   /* (N,2,49) history carried to the next step (reset envs: [next, next]) */
    float* x_norm; int64_t x_pitch;     /* (N,98) normalised discriminator input, row pitch in floats */
    const float* norm_mean;             /* (98) float32(mean) */
    const float* norm_std;              /* (98) sqrt(float32(var + eps)) */
    float norm_clip;                    /* 10 */
    int32_t task_obs_weight_decay; float task_obs_weight;
    float obs_disc_weight_step;
    /* optional (all may be NULL).  task_obs_weight_dev: (1) device scalar that overrides task_obs_weight (the weight decays
     * every iteration, on_policy_runner.py:224-225: a captured rollout graph must not freeze it).  Snapshots for a reward tail
     * that runs concurrently with the NEXT env step (which overwrites rew_buf / reset_buf / time_outs): copied here, by the
     * launch that already reads `dones` */
    const float* task_obs_weight_dev;
    const float* rewards_in; float* rewards_snap;               /* (N) */
    uint8_t* dones_snap;                                        /* (N) */
    const uint8_t* time_outs_in; uint8_t* time_outs_snap;       /* (N) */
    /* optional: the env's latent_eps (N,1) / latent_c (N,dim_c) of this step into the replay-buffer rows that go with hist_new
     * (storage/replay_buffer.py insert; saves the separate copy launch of the rollout step) */
    const float* latent_eps_in; float* latent_eps_out;
    const float* latent_c_in; float* latent_c_out;
} QaDiscInputArgs;
int qa_disc_input(const QaDiscInputArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K19 style-reward tail -- the back half of Discriminator.predict_disc_reward (discriminator.py:64-69, 90-118, MSELoss
 *     mapping) plus the reward half of SSInfoGAIL.process_env_step (gail.py:199-206): heads -> reward_i, reward_us,
 *     reward_ss (float64 cross-entropy of the already soft-maxed classifier output), weighted total in float64,
 *     + gamma * V * time_out, stored as float32 (rollout_storage.py:67).
 * ------------------------------------------------------------------------------------------ */
typedef struct QaDiscRewardArgs {
    int32_t num_envs;
    const float* heads; int64_t heads_pitch;   /* (N, 2 + dim_c): [d | eps | classifier logits] */
    const float* obs; int64_t obs_pitch; int32_t obs_width;   /* observation rows; the last dim_c + 1 lanes are eps, c */
    const float* reward_t;              /* (N) task reward (env.rew_buf) */
    float dt, coef_i, coef_us, coef_ss, coef_t;
    const float* values; int64_t values_pitch;   /* (N,1) critic values of this step (row pitch in floats), or NULL */
    const uint8_t* time_outs;           /* (N) infos["time_outs"], or NULL */
    float gamma;
    const uint8_t* dones;               /* (N) or NULL */
    float* rewards_out;                 /* (N) e.g. storage.rewards[step] */
    uint8_t* dones_out;                 /* (N) e.g. storage.dones[step], or NULL */
    float* reward_terms;                /* (N,4) reward_i, reward_us, reward_ss, reward_t (already x dt), or NULL */
} QaDiscRewardArgs;
int qa_disc_reward(const QaDiscRewardArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K24-K30  discriminator update, SURVEY 8(f)-1 -- the row-wise half of one SSInfoGAIL.update_ss_info_gail minibatch step,
 *     bbc/rsl_rl/algorithms/gail.py:415-541 (see csrc/qa_disc_update.cu for the line-by-line map).  B = rows per batch; the three
 *     batches [policy | labelled expert | unlabelled expert] are stacked into one (3B, width) matrix.
 * ------------------------------------------------------------------------------------------ */
typedef struct QaDiscPrepareArgs {
    int64_t B;
    int32_t width, obs_dim;             /* 98 = disc_obs_len * 49, 49 */
    const float* replay_states;         /* (R,width) policy replay ring */
    const float* replay_eps;            /* (R) */
    const float* replay_c;              /* (R,5) */
    const float* expert_lb;             /* (n,width) */
    const int64_t* expert_label;        /* (n) */
    const float* expert_ulb;            /* (m,width) */
    const int64_t* idx_pi; const int64_t* idx_lb; const int64_t* idx_ulb;   /* (B) each */
    int32_t task_obs_weight_decay;
    const float* task_obs_weight;       /* (1) DEVICE scalar (decays every iteration) or NULL */
    float obs_disc_weight_step;
    const float* norm_mean; const float* norm_std; float norm_clip;          /* (width) fp32 */
    float* x; int64_t x_pitch;          /* (3B,width) out */
    float* tgt_eps;                     /* (B) out: latent_eps of the policy rows */
    int32_t* tgt_c;                     /* (B) out: argmax latent_c of the policy rows */
    int32_t* tgt_label;                 /* (B) out: labels of the labelled expert rows */
} QaDiscPrepareArgs;
int qa_disc_prepare(const QaDiscPrepareArgs* a, void* stream);

typedef struct QaDiscHeadsArgs {
    int64_t B;
    const float* h2; int64_t h2_pitch;  /* (3B,256) trunk output (post ReLU) */
    const float* w_d; const float* b_d; /* linear (1,256), (1) */
    const float* w_eps; const float* b_eps;       /* encoder_eps (1,256), (1) */
    const float* w_c; int64_t w_c_pitch; const float* b_c;   /* classifier (5,256), (5) */
    const float* tgt_eps; const int32_t* tgt_c; const int32_t* tgt_label;
    float ss_coef, disc_coef, us_coef;
    const float* info_max_coef;         /* (1) device scalar (ramps up during training) or NULL = 0 */
    float* gz2; int64_t gz2_pitch;      /* (3B,256) out: d loss / d (pre-activation of trunk layer 2) */
    float* v2; int64_t v2_pitch;        /* (B,256) out: relu'(z2) * w_d for the unlabelled rows, or NULL */
    float* dw_d; float* db_d; float* dw_eps; float* db_eps; float* dw_c; int64_t dw_c_pitch; float* db_c;   /* accumulated */
    float* db2;                         /* (256) accumulated: bias gradient of trunk layer 2 */
    float* stats;                       /* (11) accumulated, the reference's return order (:540-541) */
    float* prior_batch;                 /* (5) accumulated: mean over the unlabelled rows of the class probabilities */
} QaDiscHeadsArgs;
int qa_disc_heads_loss(const QaDiscHeadsArgs* a, void* stream);

typedef struct QaDiscGpArgs {
    int64_t B; int32_t width;
    float coef;                         /* disc_grad_penalty */
    float* g; int64_t g_pitch;          /* (B,width) in: d D / d x; out: d loss / d g */
    float* stats;
} QaDiscGpArgs;
int qa_disc_gp_loss(const QaDiscGpArgs* a, void* stream);

typedef struct QaDiscRegArgs {
    const float* params; float* grads;  /* flat discriminator parameter / gradient buffers */
    int64_t seg_off[3], seg_len[3];     /* trunk.0.weight, trunk.2.weight, linear.weight (padded extents) */
    float logit_reg_coef, weight_decay_coef;
    float* stats;
} QaDiscRegArgs;
int qa_disc_reg(const QaDiscRegArgs* a, void* stream);

typedef struct QaNormMomentsArgs {
    int64_t B; int32_t width, num_batches;
    const float* x; int64_t x_pitch;    /* (num_batches*B, width) */
    double* moments;                    /* (num_batches, 2, width): mean, E[x^2] */
} QaNormMomentsArgs;
int qa_norm_moments(const QaNormMomentsArgs* a, void* stream);

typedef struct QaNormMergeArgs {
    int64_t B; int32_t width, num_batches, world_size;   /* moments hold SUMS over world_size ranks */
    const double* moments;
    double* mean; double* var; double* count;             /* running state, (width), (width), (1) */
    float* mean32; float* std32; double epsilon;          /* what the normalising kernels read */
    float* prior; const float* prior_batch; float prior_soft_coef;   /* (5) or NULL */
    float* std; const float* min_std; int32_t num_std;    /* policy std floor or NULL */
} QaNormMergeArgs;
int qa_norm_merge(const QaNormMergeArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K8c  a CHAIN of Adam steps (torch.optim.Adam semantics, weight decay added to the gradient, no clipping) over one flat
 *      parameter buffer in ONE launch: op k updates params[lo_k, hi_k) with its own moments / learning rate / step counter;
 *      an element covered by several ops receives their updates in op order, each seeing the parameter the previous one
 *      left -- the discriminator's three optimisers, which all step the shared trunk (bbc/rsl_rl/algorithms/gail.py:107-128,
 *      :519-521), as one kernel instead of fifteen graph nodes.  Every step counter is incremented by one.
 * ------------------------------------------------------------------------------------------ */
#define QA_ADAM_CHAIN_MAX 8
typedef struct QaAdamChainOp {
    int64_t lo, hi;                     /* element range in the flat buffer, multiples of 4 */
    float* exp_avg; float* exp_avg_sq;  /* (hi - lo), 16-byte aligned */
    const float* lr;                    /* (1) device scalar */
    int32_t* step;                      /* (1) device counter: value BEFORE this step */
    float weight_decay;
} QaAdamChainOp;
typedef struct QaAdamChainArgs {
    float* params; const float* grads;  /* flat buffers, 16-byte aligned */
    int32_t num_ops;
    QaAdamChainOp ops[QA_ADAM_CHAIN_MAX];
    float beta1, beta2, eps, grad_scale;
    uint32_t* ticket;                   /* (1) zero-initialised once */
} QaAdamChainArgs;
int qa_adam_chain(const QaAdamChainArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K31  peer-memory all-reduce of the PPO gradient arena fused with K8's gradient-norm pass (ONE all-reduce of the
 *      PPO gradients per optimiser step).  Every rank maps every other rank's arena and control block (cudaIpc, same node,
 *      NVLink peers); see csrc/qa_peer.cu for the protocol.  In place: afterwards every rank's arena holds the bit-identical
 *      SUM over ranks.  sumsq_out[0] / [1] (fp64, may be NULL) receive sum(g^2) of arena[0 : seg_split) and of
 *      arena[seg_split : norm_end) -- the actor-critic and the estimator gradients of the arena layout
 *      [actor-critic | estimator | kl, pad] -- i.e. what K8's own norm kernel would compute on the reduced gradients.
 * ------------------------------------------------------------------------------------------ */
#define QA_PEER_MAX_RANKS 8
typedef struct QaPeerAllreduceArgs {
    int32_t world_size, rank;
    int64_t n;                              /* floats in the arena, multiple of 4 */
    int64_t seg_split, norm_end;            /* 0 <= seg_split <= norm_end <= n */
    float* arena[QA_PEER_MAX_RANKS];        /* arena[p] = rank p's arena as mapped in THIS process (arena[rank] = own) */
    uint32_t* ctrl[QA_PEER_MAX_RANKS];      /* control blocks (qa_peer_ctrl_bytes(n) each, zero-initialised once), same mapping */
    double* sumsq_out[2];
    /* so that K8 can run as its update pass alone (qa_adam_apply): the norms are stored multiplied by grad_scale^2 (K8 squares
     * grad * grad_scale), step_inc[k] (may be NULL) is incremented by one like K8's own norm pass does, and
     * arena[scale_index] (the KL scalar of the arena layout; -1 = none) is multiplied by grad_scale (rank mean) */
    float grad_scale;
    int32_t* step_inc[2];
    int64_t scale_index;
} QaPeerAllreduceArgs;
long long qa_peer_ctrl_bytes(long long n);   /* bytes of a control block (incl. the staging area) for an arena of n floats */
/* K8's update pass alone: workspace (sum of squares of grad * grad_scale, fp64) and *step are taken as they are */
int qa_adam_apply(const QaClipAdamArgs* a, void* stream);
int qa_peer_allreduce(const QaPeerAllreduceArgs* a, void* stream);
/* exportable device buffers for the above: cudaMalloc + zero fill / cudaFree / cudaIpcGetMemHandle (64 bytes) /
 * cudaIpcOpenMemHandle (lazy peer access) / cudaIpcCloseMemHandle */
int qa_ipc_alloc(void** ptr, uint64_t bytes);
int qa_ipc_free(void* ptr);
int qa_ipc_get_handle(const void* ptr, uint8_t* handle64);
int qa_ipc_open_handle(const uint8_t* handle64, void** ptr);
int qa_ipc_close_handle(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* QA_B200_H_ */
