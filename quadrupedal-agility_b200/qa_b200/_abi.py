"""ctypes mirror of `include/qa_b200.h` and the loader of `libqa_b200.so`.

The library is the product: there is NO fallback.  `load()` raises if the shared object
has not been built (run `python -c "import __graft_entry__ as g; g.build()"` or
`make -C quadrupedal-agility_b200`), and every op wrapper raises `RuntimeError` on a
non-zero return code.
"""
import ctypes as C
import os

QA_ABI_VERSION = 1
NUM_DOF, DIM_C, NUM_REWARDS = 12, 5, 14
QA_K2_BULK_STORE = 1
QA_K2_TILED = 2
QA_K2_PDL = 4
MAX_NOISE_LANES = 64

f32p = C.POINTER(C.c_float)
vp = C.c_void_p   # every device pointer travels as a plain address


class QaActionPushArgs(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("delay", C.c_int32), ("clip", C.c_float),
                ("actions_in", vp), ("action_history_buf", vp), ("actions_out", vp)]


class QaTorqueArgs(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("action_scale", C.c_float), ("hip_scale_reduction", C.c_float),
                ("actions", vp), ("dof_state", vp), ("motor_strength", vp), ("p_gains", vp), ("d_gains", vp),
                ("default_dof_pos", vp), ("torque_limits", vp), ("torques", vp), ("torques_org", vp)]


class QaTerrain(C.Structure):
    _fields_ = [("height_samples", vp), ("rows", C.c_int32), ("cols", C.c_int32), ("border_size", C.c_float),
                ("horizontal_scale", C.c_float), ("vertical_scale", C.c_float)]


class QaHeightScanArgs(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("num_points", C.c_int32), ("root_states", vp), ("height_points", vp),
                ("terrain", QaTerrain), ("measured_heights", vp)]


class QaMocapTable(C.Structure):
    _fields_ = [("frames", vp), ("clip_start", vp), ("clip_nframes", vp), ("clip_len_s", vp),
                ("clip_frame_dur", vp), ("mode_offset", vp), ("mode_clips", vp), ("mode_cdf", vp),
                ("num_clips", C.c_int32), ("num_frames", C.c_int32)]


class QaMocapBlendArgs(C.Structure):
    _fields_ = [("num", C.c_int32), ("table", QaMocapTable), ("clip_idx", vp), ("time_u", vp),
                ("time_between_frames", C.c_double), ("disc_obs_len", C.c_int32), ("frames_out", vp)]


class QaBbcConst(C.Structure):
    _fields_ = [
        ("num_bodies", C.c_int32), ("feet_indices", C.c_int32 * 4),
        ("termination_body_mask", C.c_uint32), ("penalised_body_mask", C.c_uint32),
        ("default_dof_pos", C.c_float * 12), ("dof_pos_lower", C.c_float * 12), ("dof_pos_upper", C.c_float * 12),
        ("dof_vel_limits", C.c_float * 12), ("torque_limits", C.c_float * 12), ("hip_dof_mask", C.c_uint32),
        ("reward_scale", C.c_float * NUM_REWARDS), ("dt", C.c_float), ("tracking_sigma", C.c_float),
        ("soft_dof_vel_limit", C.c_float), ("soft_torque_limit", C.c_float), ("jump_goal", C.c_float),
        ("jump_height_lo", C.c_float), ("only_positive_rewards", C.c_int32),
        ("max_episode_length", C.c_float), ("resample_period", C.c_int32), ("episode_length_s", C.c_float),
        ("lin_vel_x", (C.c_float * 2) * DIM_C), ("lin_vel_y", (C.c_float * 2) * DIM_C),
        ("ang_vel_yaw", (C.c_float * 2) * DIM_C),
        ("jump_h_lo", C.c_float), ("jump_h_span", C.c_float), ("loco_h_lo", C.c_float), ("loco_h_span", C.c_float),
        ("lin_vel_x_clip", C.c_float), ("lin_vel_y_clip", C.c_float), ("ang_vel_yaw_clip", C.c_float),
        ("prior_cdf", C.c_float * DIM_C),
        ("s_lin_vel", C.c_float), ("s_ang_vel", C.c_float), ("s_dof_pos", C.c_float), ("s_dof_vel", C.c_float),
        ("s_key_pos", C.c_float), ("s_foot_contact", C.c_float), ("s_lin_vel_dist", C.c_float),
        ("s_ang_vel_dist", C.c_float), ("clip_obs", C.c_float), ("add_noise", C.c_int32),
        ("root_height_obs", C.c_int32), ("measure_heights", C.c_int32),
        ("center_px", C.c_float), ("center_py", C.c_float), ("max_push_vel_xy", C.c_float),
        ("time_between_frames", C.c_double), ("disc_obs_len", C.c_int32),
        ("num_noise", C.c_int32), ("noise_idx", C.c_int32 * MAX_NOISE_LANES), ("noise_scale", C.c_float * MAX_NOISE_LANES),
    ]


class QaBbcStepArgs(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("do_push", C.c_int32), ("obs_pitch", C.c_int32),
        ("contact_ring_head", C.c_int32), ("contact_ring_len", C.c_int32), ("flags", C.c_uint32),
        ("rng_seed", C.c_uint64), ("rng_step", C.c_uint64),
        ("root_states", vp), ("dof_state", vp), ("rigid_body_state", vp), ("contact_forces", vp),
        ("motor_strength", vp), ("mass_params", vp), ("friction_coeffs", vp), ("env_origins", vp),
        ("noise_scale_vec", vp), ("terrain", QaTerrain), ("mocap", QaMocapTable),
        ("episode_length_buf", vp), ("last_contacts", vp), ("commands", vp), ("latent_eps", vp),
        ("latent_c", vp), ("actions", vp), ("last_actions", vp), ("torques_org", vp),
        ("last_torques_org", vp), ("last_dof_vel", vp), ("last_root_vel", vp), ("action_history_buf", vp),
        ("obs_history_buf", vp), ("episode_sums", vp), ("feet_air_time", vp), ("contact_buf", vp),
        ("contact_force_buf", vp),
        ("obs_buf", vp), ("privileged_obs_buf", vp), ("obs_disc_buf", vp), ("rew_buf", vp), ("reset_buf", vp),
        ("time_out_buf", vp), ("base_lin_vel", vp), ("base_ang_vel", vp), ("projected_gravity", vp), ("rpy", vp),
        ("feet_forces", vp), ("contact_filt", vp), ("root_h", vp),
        ("episode_rew_means", vp), ("time_outs_latched", vp), ("num_resets", vp), ("workspace", vp),
        ("step_state", vp), ("push_interval", C.c_int32),
        ("noise_u", vp), ("rs_eps_u", vp), ("rs_c_idx", vp), ("rs_cmd_u", vp), ("rt_eps_u", vp),
        ("rt_c_idx", vp), ("rt_cmd_u", vp), ("push_u", vp), ("mocap_clip_idx", vp), ("mocap_time_u", vp), ("prior_cdf", vp),
    ]


class QaCompactArgs(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("reset_buf", vp), ("prev_obs_disc_buf", vp), ("reset_env_ids", vp),
                ("reset_env_ids_i32", vp), ("terminal_disc_states", vp), ("count", vp)]


class QaGaeArgs(C.Structure):
    _fields_ = [("num_steps", C.c_int32), ("num_envs", C.c_int32), ("gamma", C.c_float), ("lam", C.c_float),
                ("rewards", vp), ("values", vp), ("dones", vp), ("last_values", vp), ("returns", vp),
                ("advantages", vp), ("workspace", vp)]


GATHER_MAX = 16


class QaGatherArgs(C.Structure):
    _fields_ = [("num_rows", C.c_int64), ("num_tensors", C.c_int32), ("indices", vp), ("src", vp * GATHER_MAX),
                ("dst", vp * GATHER_MAX), ("width", C.c_int32 * GATHER_MAX), ("dst_pitch", C.c_int32 * GATHER_MAX),
                ("src_pitch", C.c_int32 * GATHER_MAX), ("src_col0", C.c_int32 * GATHER_MAX), ("dst_col0", C.c_int32 * GATHER_MAX)]


class QaClipAdamArgs(C.Structure):
    _fields_ = [("numel", C.c_int64), ("params", vp), ("grads", vp), ("exp_avg", vp), ("exp_avg_sq", vp),
                ("lr", vp), ("step", vp), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("max_grad_norm", C.c_float), ("grad_scale", C.c_float), ("grad_norm_out", vp), ("workspace", vp),
                ("weight_decay", C.c_float)]


class QaLinearArgs(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("act", C.c_int32), ("x", vp),
                ("x_pitch", C.c_int64), ("w", vp), ("w_pitch", C.c_int64), ("bias", vp), ("y", vp),
                ("y_pitch", C.c_int64), ("x_col0", C.c_int32), ("y_col0", C.c_int32)]


class QaActBwdArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int32), ("act", C.c_int32), ("gy", vp), ("gy_pitch", C.c_int64),
                ("y", vp), ("y_pitch", C.c_int64), ("gz", vp), ("gz_pitch", C.c_int64), ("db", vp),
                ("zero_db", C.c_int32), ("addend", vp), ("addend_pitch", C.c_int64), ("addend_scale", vp)]


class QaHeadFwdArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int32), ("Kh", C.c_int32), ("h", vp), ("h_pitch", C.c_int64), ("w", vp),
                ("w_pitch", C.c_int64), ("bias", vp), ("y", vp), ("y_pitch", C.c_int64)]


class QaPolicySampleArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("A", C.c_int32), ("mu", vp), ("mu_pitch", C.c_int64), ("std", vp), ("noise", vp),
                ("rng_seed", C.c_uint64), ("rng_step", C.c_uint64), ("step_state", vp), ("actions", vp), ("logp", vp),
                ("actions_st", vp), ("logp_st", vp), ("mu_st", vp), ("sigma_st", vp)]


class QaDiscPrepareArgs(C.Structure):
    _fields_ = [("B", C.c_int64), ("width", C.c_int32), ("obs_dim", C.c_int32), ("replay_states", vp), ("replay_eps", vp),
                ("replay_c", vp), ("expert_lb", vp), ("expert_label", vp), ("expert_ulb", vp), ("idx_pi", vp), ("idx_lb", vp),
                ("idx_ulb", vp), ("task_obs_weight_decay", C.c_int32), ("task_obs_weight", vp), ("obs_disc_weight_step", C.c_float),
                ("norm_mean", vp), ("norm_std", vp), ("norm_clip", C.c_float), ("x", vp), ("x_pitch", C.c_int64), ("tgt_eps", vp),
                ("tgt_c", vp), ("tgt_label", vp)]


class QaDiscHeadsArgs(C.Structure):
    _fields_ = [("B", C.c_int64), ("h2", vp), ("h2_pitch", C.c_int64), ("w_d", vp), ("b_d", vp), ("w_eps", vp), ("b_eps", vp),
                ("w_c", vp), ("w_c_pitch", C.c_int64), ("b_c", vp), ("tgt_eps", vp), ("tgt_c", vp), ("tgt_label", vp),
                ("ss_coef", C.c_float), ("disc_coef", C.c_float), ("us_coef", C.c_float), ("info_max_coef", vp), ("gz2", vp),
                ("gz2_pitch", C.c_int64), ("v2", vp), ("v2_pitch", C.c_int64), ("dw_d", vp), ("db_d", vp), ("dw_eps", vp),
                ("db_eps", vp), ("dw_c", vp), ("dw_c_pitch", C.c_int64), ("db_c", vp), ("db2", vp), ("stats", vp), ("prior_batch", vp)]


class QaDiscGpArgs(C.Structure):
    _fields_ = [("B", C.c_int64), ("width", C.c_int32), ("coef", C.c_float), ("g", vp), ("g_pitch", C.c_int64), ("stats", vp)]


class QaDiscRegArgs(C.Structure):
    _fields_ = [("params", vp), ("grads", vp), ("seg_off", C.c_int64 * 3), ("seg_len", C.c_int64 * 3),
                ("logit_reg_coef", C.c_float), ("weight_decay_coef", C.c_float), ("stats", vp)]


class QaNormMomentsArgs(C.Structure):
    _fields_ = [("B", C.c_int64), ("width", C.c_int32), ("num_batches", C.c_int32), ("x", vp), ("x_pitch", C.c_int64), ("moments", vp)]


class QaNormMergeArgs(C.Structure):
    _fields_ = [("B", C.c_int64), ("width", C.c_int32), ("num_batches", C.c_int32), ("world_size", C.c_int32), ("moments", vp),
                ("mean", vp), ("var", vp), ("count", vp), ("mean32", vp), ("std32", vp), ("epsilon", C.c_double), ("prior", vp),
                ("prior_batch", vp), ("prior_soft_coef", C.c_float), ("std", vp), ("min_std", vp), ("num_std", C.c_int32)]


QA_PEER_MAX_RANKS = 8
QA_ADAM_CHAIN_MAX = 8


class QaAdamChainOp(C.Structure):
    _fields_ = [("lo", C.c_int64), ("hi", C.c_int64), ("exp_avg", vp), ("exp_avg_sq", vp), ("lr", vp), ("step", vp),
                ("weight_decay", C.c_float)]


class QaAdamChainArgs(C.Structure):
    _fields_ = [("params", vp), ("grads", vp), ("num_ops", C.c_int32), ("ops", QaAdamChainOp * QA_ADAM_CHAIN_MAX),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("grad_scale", C.c_float), ("ticket", vp)]


class QaPeerAllreduceArgs(C.Structure):
    _fields_ = [("world_size", C.c_int32), ("rank", C.c_int32), ("n", C.c_int64), ("seg_split", C.c_int64), ("norm_end", C.c_int64),
                ("arena", vp * QA_PEER_MAX_RANKS), ("ctrl", vp * QA_PEER_MAX_RANKS), ("sumsq_out", vp * 2),
                ("grad_scale", C.c_float), ("step_inc", vp * 2), ("scale_index", C.c_int64)]


class QaHeadBwdArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int32), ("Kh", C.c_int32), ("act", C.c_int32), ("gz_scale", C.c_float),
                ("gz", vp), ("gz_pitch", C.c_int64), ("h", vp), ("h_pitch", C.c_int64), ("w", vp), ("w_pitch", C.c_int64),
                ("gz_prev", vp), ("gz_prev_pitch", C.c_int64), ("dw", vp), ("dw_pitch", C.c_int64), ("db", vp),
                ("db_prev", vp)]


class QaPpoLossArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("mu", vp), ("mu_pitch", C.c_int64), ("std", vp), ("value", vp),
                ("value_pitch", C.c_int64), ("actions", vp), ("old_logp", vp), ("advantages", vp), ("returns", vp),
                ("target_values", vp), ("old_mu", vp), ("old_sigma", vp), ("clip", C.c_float), ("c_surr", C.c_float),
                ("c_value", C.c_float), ("c_bound", C.c_float), ("c_entropy", C.c_float),
                ("use_clipped_value_loss", C.c_int32), ("dmu", vp), ("dvalue", vp), ("dstd", vp), ("stats", vp)]


class QaLinearBwdArgs(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("gz", vp), ("gz_pitch", C.c_int64), ("x", vp),
                ("x_pitch", C.c_int64), ("w", vp), ("w_pitch", C.c_int64), ("dx", vp), ("dx_pitch", C.c_int64),
                ("dw", vp), ("dw_pitch", C.c_int64), ("act_prev", C.c_int32), ("y_prev", vp), ("y_prev_pitch", C.c_int64),
                ("db_prev", vp), ("db_accumulate", C.c_int32), ("x_col0", C.c_int32), ("w_col0", C.c_int32)]


class QaHistEncArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("hist", vp), ("hist_pitch", C.c_int64), ("w0", vp), ("w0_pitch", C.c_int64),
                ("b0", vp), ("w1", vp), ("b1", vp), ("w2", vp), ("b2", vp), ("w3", vp), ("w3_pitch", C.c_int64),
                ("b3", vp), ("out", vp), ("out_pitch", C.c_int64)]


class QaRowLossArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("W", C.c_int32), ("mode", C.c_int32), ("a", vp), ("a_pitch", C.c_int64),
                ("b", vp), ("b_pitch", C.c_int64), ("da", vp), ("da_pitch", C.c_int64), ("loss", vp)]


class QaPpoScalarsArgs(C.Structure):
    _fields_ = [("ppo_stats", vp), ("std", vp), ("num_actions", C.c_int32), ("priv_reg_loss", vp),
                ("estimator_loss", vp), ("kl", vp), ("desired_kl", C.c_float), ("lr_min", C.c_float),
                ("lr_max", C.c_float), ("lr", vp), ("stats_accum", vp)]


class QaDepthArgs(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("crop_top", C.c_int32),
                ("crop_left", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32), ("buffer_len", C.c_int32),
                ("image_ptrs", vp), ("images", vp), ("image_stride", C.c_int64), ("episode_length_buf", vp),
                ("near_clip", C.c_float), ("far_clip", C.c_float), ("depth_noise", C.c_float), ("clip_span", C.c_float),
                ("noise_scale_u", vp),
                ("offset_u", vp), ("pixel_u", vp), ("rng_seed", C.c_uint64), ("rng_step", C.c_uint64),
                ("depth_buffer", vp)]


class QaPpoLossTscArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("logits", vp), ("logits_pitch", C.c_int64), ("mu", vp), ("mu_pitch", C.c_int64),
                ("std", vp), ("value", vp), ("value_pitch", C.c_int64), ("actions", vp), ("actions_pitch", C.c_int64),
                ("old_logp_d", vp), ("old_logp_c", vp), ("advantages", vp), ("returns", vp), ("target_values", vp),
                ("old_mu", vp), ("old_sigma", vp), ("clip", C.c_float), ("c_value", C.c_float), ("c_entropy", C.c_float),
                ("use_clipped_value_loss", C.c_int32), ("dlogits", vp), ("dlogits_pitch", C.c_int64), ("dmu", vp),
                ("dmu_pitch", C.c_int64), ("dvalue", vp), ("dstd", vp), ("stats", vp)]


TSC_NUM_REWARDS = 8


class QaTscConst(C.Structure):
    _fields_ = [("num_bodies", C.c_int32), ("feet_indices", C.c_int32 * 4), ("termination_body_mask", C.c_uint32),
                ("penalised_body_mask", C.c_uint32), ("dt", C.c_float), ("max_episode_length", C.c_float),
                ("episode_length_s", C.c_float), ("next_goal_threshold", C.c_float), ("leave_goal_threshold", C.c_float),
                ("reach_goal_delay_steps", C.c_float), ("num_goals_total", C.c_int32), ("num_goals_per_obstacle", C.c_int32),
                ("last_goal_repeat", C.c_int32), ("num_obstacle_types", C.c_int32), ("update_interval", C.c_int32),
                ("use_camera", C.c_int32), ("root_height_obs", C.c_int32), ("only_positive_rewards", C.c_int32),
                ("target_lin_vel", C.c_float), ("reward_scale", C.c_float * TSC_NUM_REWARDS),
                ("default_dof_pos", C.c_float * 12), ("base_init_state", C.c_float * 13), ("rand_yaw_range", C.c_float),
                ("rand_x_range", C.c_float), ("rand_y_range", C.c_float), ("frame_ang0", C.c_float),
                ("seesaw_dof_pos", C.c_float), ("s_lin_vel", C.c_float), ("s_ang_vel", C.c_float), ("s_dof_pos", C.c_float),
                ("s_dof_vel", C.c_float), ("s_key_pos", C.c_float), ("s_foot_contact", C.c_float),
                ("s_lin_vel_dist", C.c_float), ("s_ang_vel_dist", C.c_float), ("clip_obs", C.c_float),
                ("num_height_points", C.c_int32), ("hl_hist_len", C.c_int32), ("hl_action_dim", C.c_int32),
                ("contact_ring_len", C.c_int32)]


class QaTscStepArgs(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("global_counter", C.c_int64), ("root_states", vp), ("dof_state", vp),
                ("rigid_body_state", vp), ("contact_forces", vp), ("obst_dof_state", vp), ("seesaw_dof_index", vp),
                ("num_obst_dofs", C.c_int64), ("terrain", QaTerrain), ("x_edge_mask", vp), ("height_points", vp),
                ("env_goals", vp), ("obstacle_types", vp), ("mass_params", vp), ("friction_coeffs", vp),
                ("motor_strength", vp), ("episode_length_buf", vp), ("last_root_vel_in", vp), ("last_contacts", vp),
                ("reach_goal_timer", vp), ("cur_goal_idx", vp), ("cur_goals", vp), ("next_goals", vp), ("actions", vp),
                ("torques_org", vp), ("last_actions", vp), ("last_dof_vel", vp), ("last_torques_org", vp),
                ("last_root_vel", vp), ("commands", vp), ("latent_eps", vp), ("latent_c", vp),
                ("action_hl_history_buf", vp), ("episode_sums", vp), ("feet_air_time", vp), ("obs_history_buf", vp),
                ("action_history_buf", vp), ("contact_buf", vp), ("contact_ring_head", C.c_int32), ("measured_heights", vp),
                ("delta_yaw", vp), ("delta_next_yaw", vp), ("base_lin_vel", vp), ("base_ang_vel", vp),
                ("projected_gravity", vp), ("base_lin_acc", vp), ("rpy", vp), ("contact_filt", vp), ("target_yaw", vp),
                ("next_target_yaw", vp), ("cur_obstacle_types", vp), ("reached_goal", vp), ("reach_goal_cutoff", vp),
                ("feet_at_edge", vp), ("reset_buf", vp), ("time_out_buf", vp), ("time_outs_latched", vp), ("rew_buf", vp),
                ("episode_rew_means", vp), ("num_resets", vp), ("workspace", vp), ("obs_buf", vp), ("obs_bbc_buf", vp),
                ("obs_disc_buf", vp), ("yaw_u", vp), ("x_u", vp), ("y_u", vp), ("rng_seed", C.c_uint64),
                ("rng_step", C.c_uint64)]


class QaDiscInputArgs(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("dones", vp), ("prev_disc", vp), ("next_disc", vp), ("hist_prev", vp),
                ("hist_new", vp), ("hist_next", vp), ("x_norm", vp), ("x_pitch", C.c_int64), ("norm_mean", vp),
                ("norm_std", vp), ("norm_clip", C.c_float), ("task_obs_weight_decay", C.c_int32),
                ("task_obs_weight", C.c_float), ("obs_disc_weight_step", C.c_float), ("task_obs_weight_dev", vp),
                ("rewards_in", vp), ("rewards_snap", vp), ("dones_snap", vp), ("time_outs_in", vp), ("time_outs_snap", vp),
                ("latent_eps_in", vp), ("latent_eps_out", vp), ("latent_c_in", vp), ("latent_c_out", vp)]


class QaDiscRewardArgs(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("heads", vp), ("heads_pitch", C.c_int64), ("obs", vp), ("obs_pitch", C.c_int64),
                ("obs_width", C.c_int32), ("reward_t", vp), ("dt", C.c_float), ("coef_i", C.c_float), ("coef_us", C.c_float),
                ("coef_ss", C.c_float), ("coef_t", C.c_float), ("values", vp), ("values_pitch", C.c_int64), ("time_outs", vp), ("gamma", C.c_float),
                ("dones", vp), ("rewards_out", vp), ("dones_out", vp), ("reward_terms", vp)]


# every symbol `include/qa_b200.h` declares: name -> (restype, argtypes)
SYMBOLS = {
    "qa_version": (C.c_int, []),
    "qa_build_info": (C.c_char_p, []),
    "qa_struct_size": (C.c_int, [C.c_int]),
    "qa_action_push": (C.c_int, [C.POINTER(QaActionPushArgs), vp]),
    "qa_pd_torques": (C.c_int, [C.POINTER(QaTorqueArgs), vp]),
    "qa_height_scan": (C.c_int, [C.POINTER(QaHeightScanArgs), vp]),
    "qa_mocap_blend": (C.c_int, [C.POINTER(QaMocapBlendArgs), vp]),
    "qa_post_physics_bbc": (C.c_int, [C.POINTER(QaBbcConst), C.POINTER(QaBbcStepArgs), vp]),
    "qa_compact_resets": (C.c_int, [C.POINTER(QaCompactArgs), vp]),
    "qa_gae": (C.c_int, [C.POINTER(QaGaeArgs), vp]),
    "qa_gather_minibatch": (C.c_int, [C.POINTER(QaGatherArgs), vp]),
    "qa_clip_adam": (C.c_int, [C.POINTER(QaClipAdamArgs), vp]),
    "qa_adam_apply": (C.c_int, [C.POINTER(QaClipAdamArgs), vp]),
    "qa_adam_chain": (C.c_int, [C.POINTER(QaAdamChainArgs), vp]),
    "qa_peer_allreduce": (C.c_int, [C.POINTER(QaPeerAllreduceArgs), vp]),
    "qa_peer_ctrl_bytes": (C.c_longlong, [C.c_longlong]),
    "qa_ipc_alloc": (C.c_int, [C.POINTER(vp), C.c_uint64]),
    "qa_ipc_free": (C.c_int, [vp]),
    "qa_ipc_get_handle": (C.c_int, [vp, C.c_char_p]),
    "qa_ipc_open_handle": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "qa_ipc_close_handle": (C.c_int, [vp]),
    "qa_linear_fwd": (C.c_int, [C.POINTER(QaLinearArgs), vp]),
    "qa_linear_bwd": (C.c_int, [C.POINTER(QaLinearBwdArgs), vp]),
    "qa_act_bwd": (C.c_int, [C.POINTER(QaActBwdArgs), vp]),
    "qa_hist_encoder_fwd": (C.c_int, [C.POINTER(QaHistEncArgs), vp]),
    "qa_ppo_loss": (C.c_int, [C.POINTER(QaPpoLossArgs), vp]),
    "qa_row_loss": (C.c_int, [C.POINTER(QaRowLossArgs), vp]),
    "qa_ppo_scalars": (C.c_int, [C.POINTER(QaPpoScalarsArgs), vp]),
    "qa_depth_update": (C.c_int, [C.POINTER(QaDepthArgs), vp]),
    "qa_ppo_loss_tsc": (C.c_int, [C.POINTER(QaPpoLossTscArgs), vp]),
    "qa_post_physics_tsc_pre": (C.c_int, [C.POINTER(QaTscConst), C.POINTER(QaTscStepArgs), vp]),
    "qa_post_physics_tsc_post": (C.c_int, [C.POINTER(QaTscConst), C.POINTER(QaTscStepArgs), vp]),
    "qa_disc_input": (C.c_int, [C.POINTER(QaDiscInputArgs), vp]),
    "qa_disc_reward": (C.c_int, [C.POINTER(QaDiscRewardArgs), vp]),
    "qa_zero_async": (C.c_int, [vp, C.c_uint64, vp]),
    "qa_copy_async": (C.c_int, [vp, vp, C.c_uint64, vp]),
    "qa_head_fwd": (C.c_int, [C.POINTER(QaHeadFwdArgs), vp]),
    "qa_head_bwd": (C.c_int, [C.POINTER(QaHeadBwdArgs), vp]),
    "qa_policy_sample": (C.c_int, [C.POINTER(QaPolicySampleArgs), vp]),
    "qa_disc_prepare": (C.c_int, [C.POINTER(QaDiscPrepareArgs), vp]),
    "qa_disc_heads_loss": (C.c_int, [C.POINTER(QaDiscHeadsArgs), vp]),
    "qa_disc_gp_loss": (C.c_int, [C.POINTER(QaDiscGpArgs), vp]),
    "qa_disc_reg": (C.c_int, [C.POINTER(QaDiscRegArgs), vp]),
    "qa_norm_moments": (C.c_int, [C.POINTER(QaNormMomentsArgs), vp]),
    "qa_norm_merge": (C.c_int, [C.POINTER(QaNormMergeArgs), vp]),
}

STRUCT_ORDER = [QaActionPushArgs, QaTorqueArgs, QaTerrain, QaHeightScanArgs, QaMocapTable, QaMocapBlendArgs,
                QaBbcConst, QaBbcStepArgs, QaCompactArgs, QaGaeArgs, QaGatherArgs, QaClipAdamArgs, QaLinearArgs, QaActBwdArgs, QaPpoLossArgs, QaLinearBwdArgs, QaHistEncArgs,
                QaRowLossArgs, QaPpoScalarsArgs, QaDepthArgs, QaPpoLossTscArgs, QaTscConst, QaTscStepArgs,
                QaDiscInputArgs, QaDiscRewardArgs, QaHeadFwdArgs, QaHeadBwdArgs, QaPolicySampleArgs, QaDiscPrepareArgs, QaDiscHeadsArgs,
                QaDiscGpArgs, QaDiscRegArgs, QaNormMomentsArgs, QaNormMergeArgs, QaPeerAllreduceArgs, QaAdamChainArgs]

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libqa_b200.so")


def load():
    """dlopen libqa_b200.so, bind every declared symbol, check the ABI version.  Raises loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"qa_b200: {LIB_PATH} is missing -- the CUDA library is the product and there is no fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` from the repo root.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    ver = lib.qa_version()
    if ver != QA_ABI_VERSION:
        raise RuntimeError(f"qa_b200: ABI mismatch, library reports {ver}, bindings expect {QA_ABI_VERSION}")
    for which, st in enumerate(STRUCT_ORDER):
        want = lib.qa_struct_size(which)
        if want != C.sizeof(st):
            raise RuntimeError(f"qa_b200: struct layout mismatch for {st.__name__}: library {want} B, "
                               f"bindings {C.sizeof(st)} B")
    _LIB = lib
    return lib


def check(code: int, what: str) -> None:
    if code == 0:
        return
    if code < 0:
        kind = {-1: "QA_EINVAL (null pointer / bad size)", -2: "QA_ERANGE (dimension outside compiled limits)"}.get(
            code, "argument error")
        raise RuntimeError(f"{what}: {kind} [{code}]")
    raise RuntimeError(f"{what}: CUDA error {code}")
