"""Thin torch-tensor wrappers over the C ABI of `libqa_b200.so` (include/qa_b200.h).

Each function checks dtype / contiguity / device, passes `tensor.data_ptr()` and the current CUDA
stream, and raises `RuntimeError` on a non-zero return code.  No function here computes anything
in torch: if the library is missing, `_abi.load()` raises.
"""
import ctypes as C
from typing import Optional

import torch

from . import _abi
from . import config as K


# number of libqa_b200 kernels launched through this module (bench.py reports it as `gpu_launches`)
launches = 0


def _count(n: int) -> None:
    global launches
    launches += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor], dtype=None, name="tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"qa_b200: {name} must be a CUDA tensor (no CPU fallback exists)")
    if not t.is_contiguous():
        raise RuntimeError(f"qa_b200: {name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"qa_b200: {name} must be {dtype}, got {t.dtype}")
    return t.data_ptr()


def _p_strided(t: torch.Tensor, dtype, name) -> int:
    """Device address of a 2-D tensor with unit column stride and an arbitrary row pitch (the wrapper passes the pitch)."""
    if not t.is_cuda:
        raise RuntimeError(f"qa_b200: {name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dim() != 2 or t.stride(1) != 1 or t.dtype != dtype:
        raise RuntimeError(f"qa_b200: {name} must be a 2-D {dtype} tensor with unit column stride")
    return t.data_ptr()


def _bytep(t, name):
    if t.dtype not in (torch.bool, torch.uint8):
        raise RuntimeError(f"qa_b200: {name} must be bool/uint8")
    return _p(t, None, name)


def zero_(t: torch.Tensor) -> None:
    """Stream-ordered memset of a contiguous CUDA tensor (a memset node under graph capture, not a kernel)."""
    lib = _abi.load()
    _abi.check(lib.qa_zero_async(_p(t, None, "tensor"), t.numel() * t.element_size(), _stream()), "qa_zero_async")


def copy_(dst: torch.Tensor, src: torch.Tensor) -> None:
    """Stream-ordered device-to-device copy between contiguous CUDA tensors of equal byte size."""
    lib = _abi.load()
    n = dst.numel() * dst.element_size()
    if n != src.numel() * src.element_size():
        raise RuntimeError("qa_copy_async: size mismatch")
    _abi.check(lib.qa_copy_async(_p(dst, None, "dst"), _p(src, None, "src"), n, _stream()), "qa_copy_async")


# ---- K0 ---------------------------------------------------------------------------------------
def action_push(actions_in, action_history_buf, actions_out, delay: int, clip: float) -> None:
    """LeggedRobot.step front half, legged_robot.py:84-98 (history shifted IN PLACE)."""
    lib = _abi.load()
    n = actions_in.shape[0]
    a = _abi.QaActionPushArgs(n, int(delay), float(clip), _p(actions_in, torch.float32, "actions"),
                              _p(action_history_buf, torch.float32, "action_history_buf"),
                              _p(actions_out, torch.float32, "actions_out"))
    _abi.check(lib.qa_action_push(C.byref(a), _stream()), "qa_action_push")
    _count(1)


# ---- K1 ---------------------------------------------------------------------------------------
def pd_torques(actions, dof_state, motor_strength, p_gains, d_gains, default_dof_pos, torque_limits,
               torques, torques_org, action_scale: float, hip_scale_reduction: float) -> None:
    """LeggedRobot._compute_torques, legged_robot.py:547-579."""
    lib = _abi.load()
    f = torch.float32
    a = _abi.QaTorqueArgs(actions.shape[0], float(action_scale), float(hip_scale_reduction),
                          _p(actions, f, "actions"), _p(dof_state, f, "dof_state"),
                          _p(motor_strength, f, "motor_strength"), _p(p_gains, f, "p_gains"),
                          _p(d_gains, f, "d_gains"), _p(default_dof_pos, f, "default_dof_pos"),
                          _p(torque_limits, f, "torque_limits"), _p(torques, f, "torques"),
                          _p(torques_org, f, "torques_org"))
    _abi.check(lib.qa_pd_torques(C.byref(a), _stream()), "qa_pd_torques")
    _count(1)


def terrain_struct(height_samples, border_size, horizontal_scale, vertical_scale) -> _abi.QaTerrain:
    return _abi.QaTerrain(_p(height_samples, torch.int16, "height_samples"), height_samples.shape[0],
                          height_samples.shape[1], float(border_size), float(horizontal_scale),
                          float(vertical_scale))


# ---- K3 ---------------------------------------------------------------------------------------
def height_scan(root_states, height_points, height_samples, border_size, horizontal_scale, vertical_scale,
                measured_heights) -> None:
    """LeggedRobot._get_heights, legged_robot.py:1190-1228."""
    lib = _abi.load()
    f = torch.float32
    a = _abi.QaHeightScanArgs(root_states.shape[0], height_points.shape[0], _p(root_states, f, "root_states"),
                              _p(height_points, f, "height_points"),
                              terrain_struct(height_samples, border_size, horizontal_scale, vertical_scale),
                              _p(measured_heights, f, "measured_heights"))
    _abi.check(lib.qa_height_scan(C.byref(a), _stream()), "qa_height_scan")
    _count(1)


def mocap_struct(table) -> _abi.QaMocapTable:
    """`table`: qa_b200.mocap.MocapTable whose tensors already live on the GPU."""
    return _abi.QaMocapTable(
        _p(table.frames, torch.float32, "mocap.frames"), _p(table.clip_start, torch.int32, "mocap.clip_start"),
        _p(table.clip_nframes, torch.float64, "mocap.clip_nframes"),
        _p(table.clip_len_s, torch.float64, "mocap.clip_len_s"),
        _p(table.clip_frame_dur, torch.float64, "mocap.clip_frame_dur"),
        _p(table.mode_offset, torch.int32, "mocap.mode_offset"), _p(table.mode_clips, torch.int32, "mocap.mode_clips"),
        _p(table.mode_cdf, torch.float64, "mocap.mode_cdf"), table.num_clips, int(table.frames.shape[0]))


# ---- K4 ---------------------------------------------------------------------------------------
def mocap_blend(table, clip_idx, time_u, time_between_frames: float, disc_obs_len: int, frames_out) -> None:
    """MotionLoader.get_full_frame_at_time_batch(traj_time_sample_batch(...)), motion_loader.py:333-341, 410-447."""
    lib = _abi.load()
    a = _abi.QaMocapBlendArgs(clip_idx.shape[0], mocap_struct(table), _p(clip_idx, torch.int32, "clip_idx"),
                              _p(time_u, torch.float64, "time_u"), float(time_between_frames), int(disc_obs_len),
                              _p(frames_out, torch.float32, "frames_out"))
    _abi.check(lib.qa_mocap_blend(C.byref(a), _stream()), "qa_mocap_blend")
    _count(1)


# ---- reset compaction ---------------------------------------------------------------------------
def compact_resets(reset_buf, prev_obs_disc_buf, reset_env_ids, reset_env_ids_i32, terminal_disc_states,
                   count) -> None:
    """`reset_buf.nonzero()` + `obs_disc_buf[env_ids]`, legged_robot.py:153-154, without a host sync."""
    lib = _abi.load()
    a = _abi.QaCompactArgs(reset_buf.shape[0], _bytep(reset_buf, "reset_buf"),
                           _p(prev_obs_disc_buf, torch.float32, "prev_obs_disc_buf"),
                           _p(reset_env_ids, torch.int64, "reset_env_ids"),
                           _p(reset_env_ids_i32, torch.int32, "reset_env_ids_i32"),
                           _p(terminal_disc_states, torch.float32, "terminal_disc_states"),
                           _p(count, torch.int32, "count"))
    _abi.check(lib.qa_compact_resets(C.byref(a), _stream()), "qa_compact_resets")
    _count(1)


# ---- K5 ---------------------------------------------------------------------------------------
def gae(rewards, values, dones, last_values, returns, advantages, workspace, gamma: float, lam: float) -> None:
    """RolloutStorage.compute_returns, rollout_storage.py:97-111.  Tensors are (T,N[,1])."""
    lib = _abi.load()
    f = torch.float32
    T, N = rewards.shape[0], rewards.shape[1]
    a = _abi.QaGaeArgs(T, N, float(gamma), float(lam), _p(rewards, f, "rewards"), _p(values, f, "values"),
                       _bytep(dones, "dones"), _p(last_values, f, "last_values"), _p(returns, f, "returns"),
                       _p(advantages, f, "advantages"), _p(workspace, torch.float64, "workspace"))
    _abi.check(lib.qa_gae(C.byref(a), _stream()), "qa_gae")
    _count(2)


# ---- K6 ---------------------------------------------------------------------------------------
def gather_minibatch_windows(indices, entries) -> None:
    """One launch of K6 over `entries` = [(src, src_col0, dst, dst_col0, width)]: dst[j, dst_col0:+width] =
    src[indices[j], src_col0:+width].  `src` / `dst` are 2-D fp32 with unit column stride (any row pitch)."""
    lib = _abi.load()
    a = _abi.QaGatherArgs()
    a.num_rows, a.num_tensors = indices.shape[0], len(entries)
    if len(entries) > _abi.GATHER_MAX:
        raise RuntimeError("qa_gather_minibatch: too many tensors")
    a.indices = _p(indices, torch.int64, "indices")
    for t, (s, sc, d, dc, w) in enumerate(entries):
        if d.shape[0] != indices.shape[0] or sc + w > s.shape[1] or dc + w > d.shape[1]:
            raise RuntimeError("qa_gather_minibatch: window outside the tensor")
        a.src[t], a.dst[t] = _p_strided(s, torch.float32, "src"), _p_strided(d, torch.float32, "dst")
        a.src_pitch[t], a.dst_pitch[t] = int(s.stride(0)), int(d.stride(0))
        a.src_col0[t], a.dst_col0[t], a.width[t] = int(sc), int(dc), int(w)
    _abi.check(lib.qa_gather_minibatch(C.byref(a), _stream()), "qa_gather_minibatch")
    _count(1)


def gather_minibatch(indices, srcs, dsts) -> None:
    """dst[t][j] = src[t][indices[j]] for every (src, dst) pair (2-D float32, same width), one launch.
    Replaces the per-tensor advanced indexing of mini_batch_generator, rollout_storage.py:147-155."""
    lib = _abi.load()
    a = _abi.QaGatherArgs()
    a.num_rows, a.num_tensors = indices.shape[0], len(srcs)
    if len(srcs) > _abi.GATHER_MAX:
        raise RuntimeError("qa_gather_minibatch: too many tensors")
    a.indices = _p(indices, torch.int64, "indices")
    for t, (s, d) in enumerate(zip(srcs, dsts)):
        if s.shape[1:] != d.shape[1:] or d.shape[0] != indices.shape[0]:
            raise RuntimeError("qa_gather_minibatch: shape mismatch")
        if d.dim() == 2 and d.stride(1) == 1 and d.stride(0) != d.shape[1]:      # padded rows (TMA-legal pitch)
            a.src[t], a.dst[t] = _p(s, torch.float32, "src"), _p_strided(d, torch.float32, "dst")
            a.dst_pitch[t] = int(d.stride(0))
        else:
            a.src[t], a.dst[t] = _p(s, torch.float32, "src"), _p(d, torch.float32, "dst")
            a.dst_pitch[t] = 0
        a.width[t] = int(s[0].numel())
    _abi.check(lib.qa_gather_minibatch(C.byref(a), _stream()), "qa_gather_minibatch")
    _count(1)


# ---- K8 ---------------------------------------------------------------------------------------
def clip_adam(params, grads, exp_avg, exp_avg_sq, lr, step, workspace, beta1=0.9, beta2=0.999, eps=1e-8,
              max_grad_norm=1.0, grad_scale=1.0, grad_norm_out=None, weight_decay=0.0) -> None:
    """clip_grad_norm_ + Adam.step on a flat fp32 buffer (gail.py:409-412); `lr` (f32) and `step` (i32) are
    1-element device tensors."""
    lib = _abi.load()
    f = torch.float32
    a = _abi.QaClipAdamArgs(params.numel(), _p(params, f, "params"), _p(grads, f, "grads"), _p(exp_avg, f, "exp_avg"),
                            _p(exp_avg_sq, f, "exp_avg_sq"), _p(lr, f, "lr"), _p(step, torch.int32, "step"),
                            float(beta1), float(beta2), float(eps), float(max_grad_norm), float(grad_scale),
                            _p(grad_norm_out, f, "grad_norm_out"), _p(workspace, torch.float64, "workspace"),
                            float(weight_decay))
    _abi.check(lib.qa_clip_adam(C.byref(a), _stream()), "qa_clip_adam")
    _count(2)


def adam_apply(params, grads, exp_avg, exp_avg_sq, lr, step, workspace, beta1=0.9, beta2=0.999, eps=1e-8, max_grad_norm=1.0,
               grad_scale=1.0, grad_norm_out=None, weight_decay=0.0) -> None:
    """K8's update pass alone: `workspace` already holds sum((grad * grad_scale)^2) and `step` is already incremented (K31)."""
    lib = _abi.load()
    f = torch.float32
    a = _abi.QaClipAdamArgs(params.numel(), _p(params, f, "params"), _p(grads, f, "grads"), _p(exp_avg, f, "exp_avg"),
                            _p(exp_avg_sq, f, "exp_avg_sq"), _p(lr, f, "lr"), _p(step, torch.int32, "step"),
                            float(beta1), float(beta2), float(eps), float(max_grad_norm), float(grad_scale),
                            _p(grad_norm_out, f, "grad_norm_out"), _p(workspace, torch.float64, "workspace"),
                            float(weight_decay))
    _abi.check(lib.qa_adam_apply(C.byref(a), _stream()), "qa_adam_apply")
    _count(1)


def adam_chain(params, grads, chain, ticket, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0) -> None:
    """K8c: the Adam steps of `chain` = [(lo, hi, exp_avg, exp_avg_sq, lr, step, weight_decay), ...] on the flat buffer
    `params` / `grads`, applied per element in chain order, in ONE launch (no clipping); every `step` is incremented."""
    lib = _abi.load()
    f = torch.float32
    a = _abi.QaAdamChainArgs()
    a.params, a.grads, a.num_ops = _p(params, f, "params"), _p(grads, f, "grads"), len(chain)
    for k, (lo, hi, m, v, lr, step, wd) in enumerate(chain):
        o = a.ops[k]
        o.lo, o.hi, o.exp_avg, o.exp_avg_sq = int(lo), int(hi), _p(m, f, "exp_avg"), _p(v, f, "exp_avg_sq")
        o.lr, o.step, o.weight_decay = _p(lr, f, "lr"), _p(step, torch.int32, "step"), float(wd)
    a.beta1, a.beta2, a.eps, a.grad_scale = float(beta1), float(beta2), float(eps), float(grad_scale)
    a.ticket = _p(ticket, torch.int32, "ticket")
    _abi.check(lib.qa_adam_chain(C.byref(a), _stream()), "qa_adam_chain")
    _count(1)


def peer_allreduce(world_size, rank, n, arena_ptrs, ctrl_ptrs, seg_split=0, norm_end=0, sumsq_out=(None, None), grad_scale=1.0,
                   step_inc=(None, None), scale_index=-1) -> None:
    """K31: in-place all-reduce(SUM) of every rank's arena over NVLink peer memory (+ the gradient norms K8 needs).
    `arena_ptrs[p]` / `ctrl_ptrs[p]`: device addresses of rank p's arena / control block as mapped in this process."""
    lib = _abi.load()
    a = _abi.QaPeerAllreduceArgs()
    a.world_size, a.rank, a.n, a.seg_split, a.norm_end = int(world_size), int(rank), int(n), int(seg_split), int(norm_end)
    for p in range(world_size):
        a.arena[p], a.ctrl[p] = int(arena_ptrs[p]), int(ctrl_ptrs[p])
    for k in range(2):
        a.sumsq_out[k] = None if sumsq_out[k] is None else _p(sumsq_out[k], torch.float64, "sumsq_out")
        a.step_inc[k] = None if step_inc[k] is None else _p(step_inc[k], torch.int32, "step_inc")
    a.grad_scale, a.scale_index = float(grad_scale), int(scale_index)
    _abi.check(lib.qa_peer_allreduce(C.byref(a), _stream()), "qa_peer_allreduce")
    _count(1)


# ---- K7 ---------------------------------------------------------------------------------------
ACT_ID = {None: 0, "none": 0, "elu": 1, "relu": 2}


def linear_tc_ok(x, weight) -> bool:
    """TMA constraints of qa_linear_fwd: fp32 CUDA, unit inner stride, 16 B aligned bases and row pitches."""
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 2 and
            x.stride(1) == 1 and weight.stride(1) == 1 and x.stride(0) % 4 == 0 and weight.stride(0) % 4 == 0 and
            x.data_ptr() % 16 == 0 and weight.data_ptr() % 16 == 0 and x.shape[0] > 0)


def linear_fwd(x, weight, bias, y, act, x_col0: int = 0, y_col0: int = 0) -> None:
    """y[:, y_col0:y_col0+N] = act(x[:, x_col0:x_col0+K] @ weight.T + bias) on the tcgen05 tensor cores (TF32 operands, fp32
    accumulate).  With a column offset, `x` / `y` are the WIDE row-major tensors (16-byte aligned base and pitch) and K / N come
    from `weight`."""
    lib = _abi.load()
    M = x.shape[0]
    N, K = weight.shape
    if y.stride(1) != 1 or y.dtype != torch.float32 or not y.is_cuda:
        raise RuntimeError("qa_linear_fwd: bad output tensor")
    if (x_col0 or y_col0) and (x.shape[1] < x_col0 + K or y.shape[1] < y_col0 + N):
        raise RuntimeError("qa_linear_fwd: column window outside the tensor")
    a = _abi.QaLinearArgs(M, N, K, ACT_ID[act], x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0),
                          None if bias is None else _p(bias, torch.float32, "bias"), y.data_ptr(), y.stride(0),
                          int(x_col0), int(y_col0))
    _abi.check(lib.qa_linear_fwd(C.byref(a), _stream()), "qa_linear_fwd")
    _count(1)


def linear_bwd(gz, x, weight, dx=None, dw=None, act_prev=None, y_prev=None, db_prev=None, db_accumulate=False,
               x_col0: int = 0, w_col0: int = 0, K: int = None) -> None:
    """dx = gz @ weight[:, w_col0:w_col0+K] (overwritten) and/or dw += gz.T @ x[:, x_col0:x_col0+K] (accumulated) on the tcgen05
    kernel, MN-major operands.  With `act_prev` the dx epilogue also applies the previous layer's activation derivative (from
    its output `y_prev`) and reduces that layer's bias gradient into `db_prev` (zeroed first unless `db_accumulate`).
    `K` defaults to the width of `dx` / `dw` / `weight` / `x`, in that order."""
    lib = _abi.load()
    M, N = gz.shape
    if K is None:
        K = (dx.shape[1] if dx is not None else dw.shape[1] if dw is not None else
             weight.shape[1] if weight is not None else x.shape[1])
    a = _abi.QaLinearBwdArgs(M, N, K, gz.data_ptr(), gz.stride(0),
                             None if x is None else x.data_ptr(), 0 if x is None else x.stride(0),
                             None if weight is None else weight.data_ptr(), 0 if weight is None else weight.stride(0),
                             None if dx is None else dx.data_ptr(), 0 if dx is None else dx.stride(0),
                             None if dw is None else dw.data_ptr(), 0 if dw is None else dw.stride(0),
                             ACT_ID[act_prev], None if y_prev is None else y_prev.data_ptr(),
                             0 if y_prev is None else y_prev.stride(0),
                             None if db_prev is None else _p(db_prev, torch.float32, "db_prev"), int(bool(db_accumulate)),
                             int(x_col0), int(w_col0))
    _abi.check(lib.qa_linear_bwd(C.byref(a), _stream()), "qa_linear_bwd")
    _count((dx is not None) + (dw is not None))


def linear_bwd_ok(gz, x, weight) -> bool:
    ok = lambda t: (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and   # noqa: E731
                    t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0)
    return ok(gz) and ok(x) and ok(weight) and gz.shape[0] > 0


# ---- K11 --------------------------------------------------------------------------------------
def hist_encoder_fwd(hist, enc, out) -> None:
    """StateHistoryEncoder.forward (tsteps=10, ELU) in one kernel.  hist: (M,570) view; enc: the module."""
    lib = _abi.load()
    f = torch.float32
    w0, w3 = enc.encoder[0].weight, enc.linear_output[0].weight
    w1, w2 = enc.conv_layers[0].weight, enc.conv_layers[2].weight
    if hist.stride(1) != 1 or out.stride(1) != 1 or w0.stride(1) != 1 or w3.stride(1) != 1:
        raise RuntimeError("qa_hist_encoder_fwd: unit inner strides required")
    a = _abi.QaHistEncArgs(hist.shape[0], hist.data_ptr(), hist.stride(0), w0.data_ptr(), w0.stride(0),
                           _p(enc.encoder[0].bias, f, "b0"), _p(w1, f, "w1"), _p(enc.conv_layers[0].bias, f, "b1"),
                           _p(w2, f, "w2"), _p(enc.conv_layers[2].bias, f, "b2"), w3.data_ptr(), w3.stride(0),
                           _p(enc.linear_output[0].bias, f, "b3"), out.data_ptr(), out.stride(0))
    _abi.check(lib.qa_hist_encoder_fwd(C.byref(a), _stream()), "qa_hist_encoder_fwd")
    _count(1)


# ---- K9 ---------------------------------------------------------------------------------------
def act_bwd(gy, y, act, gz=None, db=None, zero_db=True, addend=None, addend_scale=None) -> None:
    """gz = (gy [+ addend_scale * addend]) * act'(y) (ELU' from the saved output: 1 if y > 0 else y + 1) and/or db = gz.sum(0);
    `addend_scale` is a 1-element device tensor."""
    lib = _abi.load()
    M, N = gy.shape
    for t in (gy, y, gz, addend):
        if t is not None and (t.stride(1) != 1 or t.dtype != torch.float32 or not t.is_cuda):
            raise RuntimeError("qa_act_bwd: operands must be fp32 CUDA with unit inner stride")
    a = _abi.QaActBwdArgs(M, N, ACT_ID[act], gy.data_ptr(), gy.stride(0), None if y is None else y.data_ptr(),
                          0 if y is None else y.stride(0), None if gz is None else gz.data_ptr(),
                          0 if gz is None else gz.stride(0), None if db is None else _p(db, torch.float32, "db"),
                          int(zero_db), None if addend is None else addend.data_ptr(),
                          0 if addend is None else addend.stride(0),
                          None if addend_scale is None else _p(addend_scale.reshape(1), torch.float32, "addend_scale"))
    _abi.check(lib.qa_act_bwd(C.byref(a), _stream()), "qa_act_bwd")
    _count(1)


# ---- K20 / K21 ----------------------------------------------------------------------------------
def head_fwd(h, weight, bias, y) -> None:
    """y = h @ weight.T + bias for a narrow output layer (N <= 16, Kh in {32, 64, 128}), CUDA cores, full fp32."""
    lib = _abi.load()
    f = torch.float32
    N, Kh = weight.shape
    a = _abi.QaHeadFwdArgs(h.shape[0], N, Kh, _p_strided(h, f, "h"), h.stride(0), _p_strided(weight, f, "weight"),
                           weight.stride(0), None if bias is None else _p(bias, f, "bias"), _p_strided(y, f, "y"), y.stride(0))
    _abi.check(lib.qa_head_fwd(C.byref(a), _stream()), "qa_head_fwd")
    _count(1)


def head_bwd(gz, h, weight, act, gz_prev=None, dw=None, db=None, db_prev=None, gz_scale: float = 1.0) -> None:
    """Backward of a narrow output layer in one pass over its input `h` (see include/qa_b200.h K21); dw / db / db_prev are
    ACCUMULATED.  `gz` may be (M,) for a single output column."""
    lib = _abi.load()
    f = torch.float32
    N, Kh = weight.shape
    if gz.dim() == 1:
        gz = gz.unsqueeze(1)
    a = _abi.QaHeadBwdArgs(h.shape[0], N, Kh, ACT_ID[act], float(gz_scale), _p_strided(gz, f, "gz"), gz.stride(0),
                           _p_strided(h, f, "h"), h.stride(0), _p_strided(weight, f, "weight"), weight.stride(0),
                           None if gz_prev is None else _p_strided(gz_prev, f, "gz_prev"),
                           0 if gz_prev is None else gz_prev.stride(0),
                           None if dw is None else _p_strided(dw, f, "dw"), 0 if dw is None else dw.stride(0),
                           None if db is None else _p(db, f, "db"), None if db_prev is None else _p(db_prev, f, "db_prev"))
    _abi.check(lib.qa_head_bwd(C.byref(a), _stream()), "qa_head_bwd")
    _count(1)


# ---- K22 ----------------------------------------------------------------------------------------
def policy_sample(mu, std, actions, noise=None, rng_seed=0, rng_step=0, step_state=None, logp=None, actions_st=None, logp_st=None,
                  mu_st=None, sigma_st=None) -> None:
    """actions = mu + std * n and log_prob(actions).sum(-1) (gail.py:186-196) with the storage writes fused; `noise` = the
    standard-normal draw in parity mode, else in-kernel Philox (keyed by seed / the device step counter)."""
    lib = _abi.load()
    f = torch.float32
    a = _abi.QaPolicySampleArgs(mu.shape[0], mu.shape[1], _p_strided(mu, f, "mu"), mu.stride(0), _p(std, f, "std"),
                                _p(noise, f, "noise"), int(rng_seed) & 0xFFFFFFFFFFFFFFFF, int(rng_step),
                                _p(step_state, torch.int64, "step_state"), _p(actions, f, "actions"), _p(logp, f, "logp"),
                                _p(actions_st, f, "actions_st"), _p(logp_st, f, "logp_st"), _p(mu_st, f, "mu_st"),
                                _p(sigma_st, f, "sigma_st"))
    _abi.check(lib.qa_policy_sample(C.byref(a), _stream()), "qa_policy_sample")
    _count(1)


# ---- K10 --------------------------------------------------------------------------------------
def ppo_loss(mu, std, value, actions, old_logp, advantages, returns, target_values, old_mu, old_sigma, dmu, dvalue,
             dstd, stats, clip, c_surr, c_value, c_bound, c_entropy, use_clipped_value_loss) -> None:
    """Forward + backward of the PPO loss terms of gail.py:367-408 in one kernel (see include/qa_b200.h)."""
    lib = _abi.load()
    f = torch.float32
    a = _abi.QaPpoLossArgs(mu.shape[0], mu.data_ptr(), mu.stride(0), _p(std, f, "std"), value.data_ptr(), value.stride(0),
                           _p(actions, f, "actions"), _p(old_logp, f, "old_logp"), _p(advantages, f, "advantages"),
                           _p(returns, f, "returns"), _p(target_values, f, "target_values"), _p(old_mu, f, "old_mu"),
                           _p(old_sigma, f, "old_sigma"), float(clip), float(c_surr), float(c_value), float(c_bound),
                           float(c_entropy), int(use_clipped_value_loss), _p(dmu, f, "dmu"), _p(dvalue, f, "dvalue"),
                           _p(dstd, f, "dstd"), _p(stats, f, "stats"))
    _abi.check(lib.qa_ppo_loss(C.byref(a), _stream()), "qa_ppo_loss")
    _count(1)


# ---- K2 constants ---------------------------------------------------------------------------------
def bbc_const(cfg: "K.BbcEnvConfig", prior_parameters=None) -> _abi.QaBbcConst:
    """Flattens the task configuration into the POD the fused kernel takes by value."""
    import math
    c = _abi.QaBbcConst()
    c.num_bodies = cfg.num_bodies
    for j in range(4):
        c.feet_indices[j] = cfg.feet_indices[j]
    c.termination_body_mask = sum(1 << b for b in cfg.termination_contact_indices)
    c.penalised_body_mask = sum(1 << b for b in cfg.penalised_contact_indices)
    soft = cfg.soft_dof_pos_limits()
    for d in range(12):
        c.default_dof_pos[d] = cfg.default_dof_pos[d]
        c.dof_pos_lower[d] = float(soft[d, 0])
        c.dof_pos_upper[d] = float(soft[d, 1])
        c.dof_vel_limits[d] = cfg.dof_vel_limits[d]
        c.torque_limits[d] = cfg.torque_limits[d]
    c.hip_dof_mask = sum(1 << d for d in cfg.hip_indices)
    for k, s in enumerate(cfg.reward_scales_dt()):
        c.reward_scale[k] = s
    c.dt = cfg.dt
    c.tracking_sigma = cfg.tracking_sigma
    c.soft_dof_vel_limit = cfg.soft_dof_vel_limit
    c.soft_torque_limit = cfg.soft_torque_limit
    c.jump_goal = cfg.jump_goal
    c.jump_height_lo = cfg.jump_height[0]
    c.only_positive_rewards = int(cfg.only_positive_rewards)
    c.max_episode_length = cfg.max_episode_length
    c.resample_period = cfg.resample_period
    c.episode_length_s = cfg.episode_length_s
    for m in range(K.DIM_C):
        for j in range(2):
            c.lin_vel_x[m][j] = cfg.lin_vel_x[m][j]
            c.lin_vel_y[m][j] = cfg.lin_vel_y[m][j]
            c.ang_vel_yaw[m][j] = cfg.ang_vel_yaw[m][j]
    c.jump_h_lo = cfg.jump_height[0]
    c.jump_h_span = cfg.jump_height[1] - cfg.jump_height[0]          # (hi-lo) in double, then fp32
    c.loco_h_lo = cfg.locomotion_height[0]
    c.loco_h_span = cfg.locomotion_height[1] - cfg.locomotion_height[0]
    c.lin_vel_x_clip = cfg.lin_vel_x_clip
    c.lin_vel_y_clip = cfg.lin_vel_y_clip
    c.ang_vel_yaw_clip = cfg.ang_vel_yaw_clip
    prior = [1.0 / K.DIM_C] * K.DIM_C if prior_parameters is None else [float(x) for x in prior_parameters]
    z = [p / cfg.latent_c_temperature for p in prior]                # legged_robot.py:536-538
    mx = max(z)
    ez = [math.exp(v - mx) for v in z]
    tot, acc = sum(ez), 0.0
    for k in range(K.DIM_C):
        acc += ez[k] / tot
        c.prior_cdf[k] = acc
    c.s_lin_vel, c.s_ang_vel, c.s_dof_pos, c.s_dof_vel = cfg.s_lin_vel, cfg.s_ang_vel, cfg.s_dof_pos, cfg.s_dof_vel
    c.s_key_pos, c.s_foot_contact = cfg.s_key_pos, cfg.s_foot_contact
    c.s_lin_vel_dist, c.s_ang_vel_dist = cfg.s_lin_vel_dist, cfg.s_ang_vel_dist
    c.clip_obs = cfg.clip_observations
    c.add_noise = int(cfg.add_noise)
    c.root_height_obs = int(cfg.root_height_obs)
    c.measure_heights = int(cfg.measure_heights)
    ci = cfg.center_height_index
    ny = len(cfg.measured_points_y)
    c.center_px = cfg.measured_points_x[ci // ny]
    c.center_py = cfg.measured_points_y[ci % ny]
    c.max_push_vel_xy = cfg.max_push_vel_xy
    c.time_between_frames = cfg.dt
    c.disc_obs_len = K.DISC_OBS_LEN
    ns = cfg.noise_scale_vec() if cfg.add_noise else None
    lanes = [] if ns is None else [i for i in range(K.OBS_WIDTH) if float(ns[i]) != 0.0]
    if len(lanes) > _abi.MAX_NOISE_LANES:
        raise RuntimeError(f"qa_b200: {len(lanes)} noisy observation lanes exceed QA_MAX_NOISE_LANES")
    c.num_noise = len(lanes)
    for k, i in enumerate(lanes):
        c.noise_idx[k] = i
        c.noise_scale[k] = float(ns[i])
    return c


def post_physics_bbc(const: _abi.QaBbcConst, args: _abi.QaBbcStepArgs) -> None:
    """One fused post-physics step (K2).  `args` is a pre-built, reusable struct of device pointers."""
    lib = _abi.load()
    _abi.check(lib.qa_post_physics_bbc(C.byref(const), C.byref(args), _stream()), "qa_post_physics_bbc")
    _count(1)


# ---- K12 / K13 ----------------------------------------------------------------------------------
def row_loss(a, b, da, loss, mode: int) -> None:
    """mode 0: loss = mean((a-b)^2); mode 1: loss = mean_i ||a_i - b_i||_2; da = d loss / d a (gail.py:352-365)."""
    lib = _abi.load()
    f = torch.float32
    args = _abi.QaRowLossArgs(a.shape[0], a.shape[1], mode, _p_strided(a, f, "a"), a.stride(0), _p_strided(b, f, "b"),
                              b.stride(0), _p_strided(da, f, "da"), da.stride(0), _p(loss, f, "loss"))
    _abi.check(lib.qa_row_loss(C.byref(args), _stream()), "qa_row_loss")
    _count(1)


def ppo_scalars(ppo_stats, std, priv_reg_loss, estimator_loss, kl, desired_kl, lr, stats_accum,
                lr_min=1e-5, lr_max=1e-2) -> None:
    """Adaptive-KL learning rate (gail.py:368-379) + running sums of the seven logged statistics, on the device."""
    lib = _abi.load()
    f = torch.float32
    args = _abi.QaPpoScalarsArgs(_p(ppo_stats, f, "ppo_stats"), _p(std, f, "std"), std.numel(),
                                 _p(priv_reg_loss, f, "priv_reg_loss"), _p(estimator_loss, f, "estimator_loss"),
                                 _p(kl, f, "kl"), float(desired_kl) if desired_kl else 0.0, lr_min, lr_max,
                                 _p(lr, f, "lr"), _p(stats_accum, f, "stats_accum"))
    _abi.check(lib.qa_ppo_scalars(C.byref(args), _stream()), "qa_ppo_scalars")
    _count(1)


# ---- K14 ----------------------------------------------------------------------------------------
def depth_update(image_ptrs, images, episode_length_buf, depth_buffer, in_h, in_w, crop_top, crop_left, near_clip,
                 far_clip, depth_noise, noise_scale_u=None, offset_u=None, pixel_u=None, rng_seed=0, rng_step=0) -> None:
    """TSC update_depth_buffer for all envs (tsc/.../legged_robot.py:154-202).  `image_ptrs`: int64 device tensor of
    N camera-tensor addresses, or `images`: one (N,in_h,in_w) tensor."""
    lib = _abi.load()
    f = torch.float32
    N, L, H, W = depth_buffer.shape
    a = _abi.QaDepthArgs()
    a.num_envs, a.in_h, a.in_w, a.crop_top, a.crop_left, a.out_h, a.out_w, a.buffer_len = N, in_h, in_w, crop_top, crop_left, H, W, L
    a.image_ptrs = _p(image_ptrs, torch.int64, "image_ptrs")
    a.images = _p(images, f, "images")
    a.image_stride = in_h * in_w
    if image_ptrs is not None and image_ptrs.numel() != N:
        raise RuntimeError("qa_depth_update: one camera pointer per env required")
    a.episode_length_buf = _p(episode_length_buf, torch.int64, "episode_length_buf")
    a.near_clip, a.far_clip, a.depth_noise = float(near_clip), float(far_clip), float(depth_noise)
    a.clip_span = float(far_clip - near_clip)
    a.noise_scale_u, a.offset_u, a.pixel_u = _p(noise_scale_u, f, "noise_scale_u"), _p(offset_u, f, "offset_u"), _p(pixel_u, f, "pixel_u")
    a.rng_seed, a.rng_step = int(rng_seed), int(rng_step)
    a.depth_buffer = _p(depth_buffer, f, "depth_buffer")
    _abi.check(lib.qa_depth_update(C.byref(a), _stream()), "qa_depth_update")
    _count(1)


# ---- K15 ----------------------------------------------------------------------------------------
def ppo_loss_tsc(logits, mu, std, value, actions, old_logp_d, old_logp_c, advantages, returns, target_values, old_mu,
                 old_sigma, dlogits, dmu, dvalue, dstd, stats, clip, c_value, c_entropy, clipped_value) -> None:
    """TSC PPO loss terms + gradients w.r.t. (logits, mu, value, std) in one launch (tsc ppo.py:176-262)."""
    lib = _abi.load()
    f = torch.float32
    a = _abi.QaPpoLossTscArgs()
    a.M = mu.shape[0]
    a.logits, a.logits_pitch = _p_strided(logits, f, "logits"), logits.stride(0)
    a.mu, a.mu_pitch = _p_strided(mu, f, "mu"), mu.stride(0)
    a.std = _p(std, f, "std")
    a.value, a.value_pitch = value.data_ptr(), value.stride(0)
    a.actions, a.actions_pitch = _p_strided(actions, f, "actions"), actions.stride(0)
    a.old_logp_d, a.old_logp_c = _p(old_logp_d, f, "old_logp_d"), _p(old_logp_c, f, "old_logp_c")
    a.advantages, a.returns, a.target_values = _p(advantages, f, "advantages"), _p(returns, f, "returns"), _p(target_values, f, "target_values")
    a.old_mu, a.old_sigma = _p(old_mu, f, "old_mu"), _p(old_sigma, f, "old_sigma")
    a.clip, a.c_value, a.c_entropy, a.use_clipped_value_loss = float(clip), float(c_value), float(c_entropy), int(bool(clipped_value))
    a.dlogits, a.dlogits_pitch = _p_strided(dlogits, f, "dlogits"), dlogits.stride(0)
    a.dmu, a.dmu_pitch = _p_strided(dmu, f, "dmu"), dmu.stride(0)
    a.dvalue, a.dstd, a.stats = _p(dvalue, f, "dvalue"), _p(dstd, f, "dstd"), _p(stats, f, "stats")
    _abi.check(lib.qa_ppo_loss_tsc(C.byref(a), _stream()), "qa_ppo_loss_tsc")
    _count(1)


# ---- K16 / K17 ------------------------------------------------------------------------------------
def post_physics_tsc(const, args, which: str) -> None:
    """TSC post_physics_step halves (tsc/.../legged_robot.py:226-298): which = "pre" | "post"."""
    lib = _abi.load()
    fn = lib.qa_post_physics_tsc_pre if which == "pre" else lib.qa_post_physics_tsc_post
    _abi.check(fn(C.byref(const), C.byref(args), _stream()), f"qa_post_physics_tsc_{which}")
    _count(1)


# ---- K18 / K19 ------------------------------------------------------------------------------------
def disc_input(dones, prev_disc, next_disc, hist_prev, hist_new, hist_next, x_norm, norm_mean, norm_std, norm_clip,
               task_obs_weight_decay, task_obs_weight, obs_disc_weight_step, snapshots=None, latents=None) -> None:
    """Disc-history bookkeeping of the rollout step + the normalised discriminator input (on_policy_runner.py:163-181,
    discriminator.py:74-88).  `task_obs_weight`: float or 0-d / (1,) device tensor (read at run time: a captured graph sees the
    decayed value).  `snapshots` = (rewards_in, rewards_snap, dones_snap, time_outs_in or None, time_outs_snap or None): copies
    of the env buffers the NEXT env step overwrites, for a reward tail that runs concurrently with it.  `latents` =
    (latent_eps (N,1), latent_eps_out, latent_c (N,dim_c), latent_c_out): the replay-buffer rows that go with hist_new."""
    lib = _abi.load()
    f = torch.float32
    tow_dev = task_obs_weight if torch.is_tensor(task_obs_weight) else None
    a = _abi.QaDiscInputArgs(dones.shape[0], _bytep(dones, "dones"), _p(prev_disc, f, "prev_disc"), _p(next_disc, f, "next_disc"),
                             _p(hist_prev, f, "hist_prev"), _p(hist_new, f, "hist_new"), _p(hist_next, f, "hist_next"),
                             _p_strided(x_norm, f, "x_norm"), x_norm.stride(0), _p(norm_mean, f, "norm_mean"),
                             _p(norm_std, f, "norm_std"), float(norm_clip), int(bool(task_obs_weight_decay)),
                             1.0 if tow_dev is not None else float(task_obs_weight), float(obs_disc_weight_step))
    if tow_dev is not None:
        a.task_obs_weight_dev = _p(tow_dev.reshape(1), f, "task_obs_weight")
    if snapshots is not None:
        r_in, r_snap, d_snap, t_in, t_snap = snapshots
        a.rewards_in, a.rewards_snap, a.dones_snap = _p(r_in, f, "rewards_in"), _p(r_snap, f, "rewards_snap"), _bytep(d_snap, "dones_snap")
        if t_in is not None:
            a.time_outs_in, a.time_outs_snap = _bytep(t_in, "time_outs_in"), _bytep(t_snap, "time_outs_snap")
    if latents is not None:
        e_in, e_out, c_in, c_out = latents
        if c_in.shape[-1] != 5 or not (e_out.is_contiguous() and c_out.is_contiguous()):
            raise RuntimeError("qa_b200: disc_input latents must be (N,1) / (N,5) contiguous rows")
        a.latent_eps_in, a.latent_eps_out = _p(e_in, f, "latent_eps"), _p(e_out, f, "latent_eps_out")
        a.latent_c_in, a.latent_c_out = _p(c_in, f, "latent_c"), _p(c_out, f, "latent_c_out")
    _abi.check(lib.qa_disc_input(C.byref(a), _stream()), "qa_disc_input")
    _count(1)


def disc_reward(heads, obs, reward_t, dt, coefs, rewards_out, values=None, time_outs=None, gamma=0.0, dones=None,
                dones_out=None, reward_terms=None) -> None:
    """Style-reward tail + time-out bootstrap (discriminator.py:90-118, gail.py:199-206); coefs = (i, us, ss, t)."""
    lib = _abi.load()
    f = torch.float32
    a = _abi.QaDiscRewardArgs()
    a.num_envs = heads.shape[0]
    a.heads, a.heads_pitch = _p_strided(heads, f, "heads"), heads.stride(0)
    a.obs, a.obs_pitch, a.obs_width = _p_strided(obs, f, "obs"), obs.stride(0), obs.shape[1]
    a.reward_t = _p(reward_t, f, "reward_t")
    a.dt, (a.coef_i, a.coef_us, a.coef_ss, a.coef_t) = float(dt), (float(c) for c in coefs)
    if values is not None and values.dim() == 1:
        values = values.unsqueeze(1)
    a.values = None if values is None else _p_strided(values, f, "values")
    a.values_pitch = 1 if values is None else values.stride(0)
    a.time_outs = None if time_outs is None else _bytep(time_outs, "time_outs")
    a.gamma = float(gamma)
    a.dones = None if dones is None else _bytep(dones, "dones")
    a.rewards_out = _p(rewards_out, f, "rewards_out")
    a.dones_out = None if dones_out is None else _bytep(dones_out, "dones_out")
    a.reward_terms = None if reward_terms is None else _p(reward_terms, f, "reward_terms")
    _abi.check(lib.qa_disc_reward(C.byref(a), _stream()), "qa_disc_reward")
    _count(1)


# ---- K24 - K30: discriminator update ---------------------------------------------------------------------------------
def disc_prepare(B, replay, expert, idx_pi, idx_lb, idx_ulb, task_obs_weight_decay, task_obs_weight, obs_disc_weight_step,
                 norm_mean, norm_std, norm_clip, x, tgt_eps, tgt_c, tgt_label, obs_dim) -> None:
    """gail.py:419-452: X = normalise(weight(gather(...))) for [policy | labelled | unlabelled] rows + per-row targets.
    `replay` = ReplayBuffer (states, latent_eps, latent_c), `expert` = object with preloaded_s_lb / preloaded_label / preloaded_s_ulb."""
    lib = _abi.load()
    f, i64, i32 = torch.float32, torch.int64, torch.int32
    a = _abi.QaDiscPrepareArgs()
    a.B, a.width, a.obs_dim = int(B), int(replay.states.shape[1]), int(obs_dim)
    a.replay_states, a.replay_eps, a.replay_c = _p(replay.states, f), _p(replay.latent_eps, f), _p(replay.latent_c, f)
    a.expert_lb, a.expert_label, a.expert_ulb = _p(expert.preloaded_s_lb, f), _p(expert.preloaded_label, i64), _p(expert.preloaded_s_ulb, f)
    a.idx_pi, a.idx_lb, a.idx_ulb = _p(idx_pi, i64), _p(idx_lb, i64), _p(idx_ulb, i64)
    a.task_obs_weight_decay = int(bool(task_obs_weight_decay))
    a.task_obs_weight = None if task_obs_weight is None else _p(task_obs_weight.reshape(1), f)
    a.obs_disc_weight_step = float(obs_disc_weight_step)
    a.norm_mean, a.norm_std, a.norm_clip = _p(norm_mean, f), _p(norm_std, f), float(norm_clip)
    a.x, a.x_pitch = _p_strided(x, f, "x"), x.stride(0)
    a.tgt_eps, a.tgt_c, a.tgt_label = _p(tgt_eps, f), _p(tgt_c, i32), _p(tgt_label, i32)
    _abi.check(lib.qa_disc_prepare(C.byref(a), _stream()), "qa_disc_prepare")
    _count(1)


def disc_heads_loss(B, h2, disc, tgt_eps, tgt_c, tgt_label, ss_coef, disc_coef, us_coef, info_max_coef, gz2, v2, stats,
                    prior_batch) -> None:
    """gail.py:454-490 on the trunk output: heads forward, loss terms, gradients, prior estimate, accuracies (K25).
    `disc` = Discriminator module whose parameters live in a flat buffer (gradients are accumulated into their `.grad`)."""
    lib = _abi.load()
    f = torch.float32
    lin, eps, cls = disc.linear, disc.encoder_eps, disc.classifier
    a = _abi.QaDiscHeadsArgs()
    a.B = int(B)
    a.h2, a.h2_pitch = _p_strided(h2, f, "h2"), h2.stride(0)
    a.w_d, a.b_d = _p_strided(lin.weight, f, "w_d"), _p(lin.bias, f)
    a.w_eps, a.b_eps = _p_strided(eps.weight, f, "w_eps"), _p(eps.bias, f)
    a.w_c, a.w_c_pitch, a.b_c = _p_strided(cls.weight, f, "w_c"), cls.weight.stride(0), _p(cls.bias, f)
    a.tgt_eps, a.tgt_c, a.tgt_label = _p(tgt_eps, f), _p(tgt_c, torch.int32), _p(tgt_label, torch.int32)
    a.ss_coef, a.disc_coef, a.us_coef = float(ss_coef), float(disc_coef), float(us_coef)
    a.info_max_coef = None if info_max_coef is None else _p(info_max_coef.reshape(1), f)
    a.gz2, a.gz2_pitch = _p_strided(gz2, f, "gz2"), gz2.stride(0)
    a.v2, a.v2_pitch = (None, 0) if v2 is None else (_p_strided(v2, f, "v2"), v2.stride(0))
    a.dw_d, a.db_d = _p_strided(lin.weight.grad, f, "dw_d"), _p(lin.bias.grad, f)
    a.dw_eps, a.db_eps = _p_strided(eps.weight.grad, f, "dw_eps"), _p(eps.bias.grad, f)
    a.dw_c, a.dw_c_pitch, a.db_c = _p_strided(cls.weight.grad, f, "dw_c"), cls.weight.grad.stride(0), _p(cls.bias.grad, f)
    a.db2 = _p(disc.trunk[2].bias.grad, f)
    a.stats, a.prior_batch = _p(stats, f), _p(prior_batch, f)
    _abi.check(lib.qa_disc_heads_loss(C.byref(a), _stream()), "qa_disc_heads_loss")
    _count(1)


def disc_gp_loss(g, coef, stats) -> None:
    """Gradient penalty mean_i ||g_i||^2 (gail.py:492-502): value into stats[4], g := d loss / d g in place (K27)."""
    lib = _abi.load()
    a = _abi.QaDiscGpArgs(g.shape[0], g.shape[1], float(coef), _p_strided(g, torch.float32, "g"), g.stride(0), _p(stats, torch.float32))
    _abi.check(lib.qa_disc_gp_loss(C.byref(a), _stream()), "qa_disc_gp_loss")
    _count(1)


def disc_reg(flat, segments, logit_reg_coef, weight_decay_coef, stats) -> None:
    """Logit regulariser + weight decay (gail.py:488-490, 504-507): values into stats[5:7], gradients into flat.grad (K28).
    `segments` = [(offset, length)] of trunk.0.weight, trunk.2.weight, linear.weight in the flat buffer."""
    lib = _abi.load()
    a = _abi.QaDiscRegArgs()
    a.params, a.grads = _p(flat.data, torch.float32), _p(flat.grad, torch.float32)
    for k, (o, n) in enumerate(segments):
        a.seg_off[k], a.seg_len[k] = int(o), int(n)
    a.logit_reg_coef, a.weight_decay_coef, a.stats = float(logit_reg_coef), float(weight_decay_coef), _p(stats, torch.float32)
    _abi.check(lib.qa_disc_reg(C.byref(a), _stream()), "qa_disc_reg")
    _count(1)


def norm_moments(x, B, num_batches, moments) -> None:
    """Column mean and E[x^2] (fp64) of each of the `num_batches` row blocks of x (K29)."""
    lib = _abi.load()
    a = _abi.QaNormMomentsArgs(int(B), x.shape[1], int(num_batches), _p_strided(x, torch.float32, "x"), x.stride(0),
                               _p(moments, torch.float64))
    _abi.check(lib.qa_norm_moments(C.byref(a), _stream()), "qa_norm_moments")
    _count(1)


def norm_merge(B, num_batches, world_size, moments, mean, var, count, mean32, std32, epsilon, prior=None, prior_batch=None,
               prior_soft_coef=0.0, std=None, min_std=None) -> None:
    """Sequential Chan merge of the batch moments into the running normaliser (utils.py:63-83), prior soft update
    (gail.py:462-464), policy-std floor (:523-524), all on the device (K30)."""
    lib = _abi.load()
    f, f64 = torch.float32, torch.float64
    a = _abi.QaNormMergeArgs(int(B), mean.shape[0], int(num_batches), int(world_size), _p(moments, f64), _p(mean, f64), _p(var, f64),
                             _p(count.reshape(1), f64), _p(mean32, f), _p(std32, f), float(epsilon), _p(prior, f), _p(prior_batch, f),
                             float(prior_soft_coef), _p(std, f), _p(min_std, f), 0 if std is None else int(std.numel()))
    _abi.check(lib.qa_norm_merge(C.byref(a), _stream()), "qa_norm_merge")
    _count(1)
