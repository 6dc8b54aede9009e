"""`PpoStepPlan` -- one PPO minibatch step of `SSInfoGAIL.update_actor_critic` (bbc/rsl_rl/algorithms/gail.py:328-413) as a
STATIC SCHEDULE of libqa_b200 launches on preallocated buffers: no autograd graph, no framework kernel in the step.

What the reference does per minibatch (and what `SSInfoGAIL._forward_backward` mirrors on autograd): estimator forward + MSE +
its own optimiser step (:359-365); privileged-latent encoder -> actor -> Normal (:338-342); critic (:343); history-latent
regulariser (:352-357); KL / surrogate / value / bound losses (:367-408); backward; clip + Adam (:409-412).  The network
topology is fixed for a whole training run, so the backward pass is known ahead of time and is written out here layer by layer:

  wide layers (>= 29 outputs)   K7  `qa_linear_fwd` / `qa_linear_bwd`: tcgen05 TF32 GEMMs, bias + activation in the forward
                                    epilogue, activation derivative + bias gradient of the previous layer in the dX epilogue,
                                    split-K dW reduced straight into the flat gradient buffer
  narrow heads (1 / 4 / 12)     K20 / K21 `qa_head_fwd` / `qa_head_bwd`: fp32 CUDA-core row streams
  losses                        K10 `qa_ppo_loss`, K12 `qa_row_loss` (value + gradient in one launch)
  column windows                the 57 proprioceptive lanes are read IN PLACE out of the 671-wide observation row (TMA coordinates),
                                the encoder writes its 29 outputs straight into the actor's input row, the actor-input gradient is
                                taken only for the 16-byte aligned window that covers the latent lanes

Three independent chains (estimator | critic | encoder + actor) run on three streams; under CUDA-graph capture they become
parallel branches of one graph, so one chain's launch latency and tail hide behind another's main loop.

Minibatch data lives in `num_sets` buffer sets: `SSInfoGAIL.update` gathers every minibatch ONCE per update (the reference
re-indexes the same four slices in each of the five epochs, rollout_storage.py:125, 140-145) and replays one graph per set.
"""
import contextlib
from typing import List, Optional

import torch
import torch.nn as nn

from .. import ops


def _padded(rows: int, cols: int, device) -> torch.Tensor:
    """(rows, cols) fp32 view of a buffer whose row pitch is a multiple of 4 floats (TMA's 16-byte rule)."""
    return torch.zeros(rows, (cols + 3) // 4 * 4, device=device, dtype=torch.float32)[:, :cols]


def _linears(seq) -> Optional[List[nn.Linear]]:
    """The Linear modules of a Linear/ELU(or ReLU) stack whose every Linear is followed by the activation; None otherwise."""
    mods = list(seq)
    if len(mods) % 2 != 0:
        return None
    out = []
    for i in range(0, len(mods), 2):
        if not isinstance(mods[i], nn.Linear) or not isinstance(mods[i + 1], (nn.ELU, nn.ReLU)):
            return None
        out.append(mods[i])
    return out


class _Cur:
    """Fork / join of side streams off the current stream (event dependencies; parallel branches under graph capture)."""

    def __init__(self, cuda: bool):
        self.s = torch.cuda.current_stream() if cuda else None

    def fork(self, side):
        if self.s is not None:
            side.wait_stream(self.s)

    def join(self, side):
        if self.s is not None:
            self.s.wait_stream(side)


class _Chain:
    """A trunk of wide Linear+act layers followed by an optional narrow linear head, with its activation / gradient buffers."""

    def __init__(self, trunk: List[nn.Linear], head: Optional[nn.Linear], act: str, M: int, dev):
        self.trunk, self.head, self.act = trunk, head, act
        self.h = [_padded(M, l.out_features, dev) for l in trunk]            # layer outputs
        self.gz = [_padded(M, l.out_features, dev) for l in trunk]           # gradients w.r.t. the pre-activations
        self.out = _padded(M, head.out_features, dev) if head is not None else None


class PpoStepPlan:
    @staticmethod
    def supported(alg) -> Optional[str]:
        """None when `alg`'s networks fit the schedule, else the reason they do not (the caller keeps the autograd path)."""
        ac, est = alg.actor_critic, alg.estimator
        if torch.device(alg.device).type != "cuda":
            return "not a CUDA device"
        if not alg.fused_loss:
            return "fused_loss is off"
        if not ac.train_with_estimated_latent or isinstance(ac.priv_encoder, nn.Identity):
            return "the actor does not read the privileged-latent encoder"
        if ac.activation_name not in ("elu", "relu"):
            return f"activation {ac.activation_name}"
        for seq in (ac.priv_encoder, ac.actor_trunk, ac.critic_trunk):
            if _linears(seq) is None:
                return "a trunk is not a Linear/activation stack"
        e = list(est.estimator)
        if len(e) < 3 or not isinstance(e[-1], nn.Linear) or _linears(e[:-1]) is None:
            return "estimator is not Linear/activation ... Linear"
        for head, trunk in ((ac.actor_head, ac.actor_trunk), (ac.critic_head, ac.critic_trunk), (e[-1], None)):
            if head.out_features > 16 or head.in_features not in (32, 64, 128):
                return "head shape outside K20/K21"
        if ac.num_latent > 32 or alg.num_explicit > 32:
            return "latent wider than a warp"
        if ac.fixed_std:
            return "fixed_std"
        return None

    def __init__(self, alg, mb_size: int, num_sets: int):
        self.alg, self.M, self.num_sets = alg, mb_size, num_sets
        dev = self.dev = torch.device(alg.device)
        ac, est, st = alg.actor_critic, alg.estimator, alg.storage
        self.act = ac.activation_name
        self.p, self.e, self.l = alg.num_prop, alg.num_explicit, alg.num_latent
        self.h = alg.num_hist * alg.num_prop
        self.W = st.observations.shape[-1]
        self.Wc = st.privileged_observations.shape[-1] if st.privileged_observations is not None else self.W
        self.A = st.actions.shape[-1]
        self.n_in = ac.num_actor_obs                                    # prop + explicit + latent + command
        self.n_cmd = self.n_in - self.p - self.e - self.l
        M, z = mb_size, (lambda w: torch.zeros(mb_size, w, device=dev))
        self.sets = []
        for _ in range(num_sets):
            self.sets.append(dict(obs=_padded(M, self.W, dev), critic_obs=_padded(M, self.Wc, dev), xa=_padded(M, self.n_in, dev),
                                  lat_in=_padded(M, self.l, dev),
                                  actions=z(self.A), values=z(1), returns=z(1), old_actions_log_prob=z(1), advantages=z(1),
                                  old_mu=z(self.A), old_sigma=z(self.A), hist_latent=_padded(M, self.l, dev)))
        est_mods = list(est.estimator)
        self.c_est = _Chain(_linears(est_mods[:-1]), est_mods[-1], self.act, M, dev)
        self.c_priv = _Chain(_linears(ac.priv_encoder), None, self.act, M, dev)
        self.c_actor = _Chain(_linears(ac.actor_trunk), ac.actor_head, self.act, M, dev)
        self.c_critic = _Chain(_linears(ac.critic_trunk), ac.critic_head, self.act, M, dev)
        self.d_est = _padded(M, self.e, dev)
        self.d_reg = _padded(M, self.l, dev)
        # TMA boxes start on 16-byte boundaries: the actor-input gradient is taken for the aligned column window that covers the
        # latent lanes [p+e, p+e+l) and the activation-backward kernel reads the lanes it needs out of it
        self.w_lo = (self.p + self.e) // 4 * 4
        self.w_n = min((self.p + self.e + self.l - self.w_lo + 3) // 4 * 4, self.n_in - self.w_lo)
        self.dplat_win = _padded(M, self.w_n, dev)
        self.dplat = self.dplat_win[:, self.p + self.e - self.w_lo:self.p + self.e - self.w_lo + self.l]
        self.dmu, self.dvalue = z(self.A), torch.zeros(M, device=dev)
        self._cuda = dev.type == "cuda"           # (the host tests drive the schedule on CPU through stand-in ops: no streams)
        self.s_critic = torch.cuda.Stream(device=dev) if self._cuda else None
        self.s_est = torch.cuda.Stream(device=dev) if self._cuda else None
        self.gather_keys = ("actions", "values", "returns", "old_actions_log_prob", "advantages", "old_mu", "old_sigma")

    # ---- data movement -------------------------------------------------------------------------------------------------
    def gather(self, k: int, idx: torch.Tensor, hist_latent_all: torch.Tensor) -> None:
        """Minibatch `k` := storage rows `idx` (K6, one launch): the nine tensors of the reference's generator
        (rollout_storage.py:147-155), the pre-computed history latents, and the two static windows of the actor's input row."""
        v, s = self.alg.storage.flat_views(), self.sets[k]
        p, e, l, h = self.p, self.e, self.l, self.h
        ent = [(v["obs"], 0, s["obs"], 0, self.W), (v["critic_obs"], 0, s["critic_obs"], 0, self.Wc),
               (v["obs"], 0, s["xa"], 0, p + e), (v["obs"], p + e + l + h, s["xa"], p + e + l, self.n_cmd),
               (v["obs"], p + e, s["lat_in"], 0, l), (hist_latent_all, 0, s["hist_latent"], 0, l)]
        for key in self.gather_keys:
            ent.append((v[key], 0, s[key], 0, v[key].shape[1]))
        ops.gather_minibatch_windows(idx, ent)

    def load(self, k: int, sample) -> None:
        """Minibatch `k` := the reference generator's tuple (update_actor_critic entry point); torch copies, not on the hot path."""
        s = self.sets[k]
        obs, critic_obs, actions, target_values, advantages, returns, old_logp, old_mu, old_sigma = sample[:9]
        p, e, l, h = self.p, self.e, self.l, self.h
        s["obs"].copy_(obs)
        s["critic_obs"].copy_(critic_obs)
        s["xa"][:, :p + e].copy_(obs[:, :p + e])
        s["xa"][:, p + e + l:].copy_(obs[:, p + e + l + h:])
        s["lat_in"].copy_(obs[:, p + e:p + e + l])
        for key, val in (("actions", actions), ("values", target_values), ("advantages", advantages), ("returns", returns),
                         ("old_actions_log_prob", old_logp), ("old_mu", old_mu), ("old_sigma", old_sigma)):
            s[key].copy_(val.reshape(s[key].shape))

    # ---- the schedule --------------------------------------------------------------------------------------------------
    def _trunk_fwd(self, c: _Chain, x, x_col0=0, last_y=None, last_y_col0=0):
        """`x[:, x_col0 : x_col0 + K]` is the trunk's input (K = the first layer's fan-in); the last layer writes
        `last_y[:, last_y_col0 : ...]` instead of its own buffer when given."""
        for i, lin in enumerate(c.trunk):
            if i == len(c.trunk) - 1 and last_y is not None:
                ops.linear_fwd(x, lin.weight, lin.bias, last_y, c.act, x_col0=x_col0, y_col0=last_y_col0)
            else:
                ops.linear_fwd(x, lin.weight, lin.bias, c.h[i], c.act, x_col0=x_col0)
            x, x_col0 = c.h[i], 0
        if c.head is not None:
            ops.head_fwd(c.h[-1], c.head.weight, c.head.bias, c.out)

    def _trunk_bwd(self, c: _Chain, x_in, x_col0, k_in):
        """Backward through the trunk given the gradient w.r.t. its LAST pre-activation in c.gz[-1] (set by the head's backward
        or by the caller).  `x_in[:, x_col0 : x_col0+k_in]` is the trunk's input (no gradient flows into it here)."""
        n = len(c.trunk)
        for i in range(n - 1, 0, -1):                                   # dx chain first: it is the critical path
            lin, prev = c.trunk[i], c.trunk[i - 1]
            ops.linear_bwd(c.gz[i], None, lin.weight, dx=c.gz[i - 1], act_prev=c.act, y_prev=c.h[i - 1], db_prev=prev.bias.grad,
                           db_accumulate=True)
        for i in range(n - 1, 0, -1):
            ops.linear_bwd(c.gz[i], c.h[i - 1], None, dw=c.trunk[i].weight.grad)
        ops.linear_bwd(c.gz[0], x_in, None, dw=c.trunk[0].weight.grad, x_col0=x_col0, K=k_in)

    def _head_bwd(self, c: _Chain, gz_out, scale=1.0):
        last = c.trunk[-1]
        ops.head_bwd(gz_out, c.h[-1], c.head.weight, c.act, gz_prev=c.gz[-1], dw=c.head.weight.grad, db=c.head.bias.grad,
                     db_prev=last.bias.grad, gz_scale=scale)

    def forward_backward(self, k: int, step_estimator: bool = False) -> None:
        """Forward + backward of minibatch set `k`; leaves the gradients in the flat buffers, the PPO statistics in
        `alg._ppo_stats` (surrogate, value, bound, kl) and `alg._aux_loss` (priv_reg, estimator).  `step_estimator`: on a single
        rank the estimator's optimiser step (gail.py:361-365) is issued inside the estimator's chain (`_apply` then skips it)."""
        alg, s = self.alg, self.sets[k]
        ac = alg.actor_critic
        p, e, l = self.p, self.e, self.l
        obs, xa = s["obs"], s["xa"]
        cur = _Cur(self._cuda)
        on = (lambda st: torch.cuda.stream(st)) if self._cuda else (lambda st: contextlib.nullcontext())
        if alg._grad_arena is not None:
            ops.zero_(alg._grad_arena)
        else:
            ops.zero_(alg.ac_flat.grad)
            ops.zero_(alg.est_flat.grad)
        cur.fork(self.s_critic)
        cur.fork(self.s_est)
        # ---- estimator: forward, MSE (gail.py:359), backward -- independent of everything else --------------------------------
        with on(self.s_est):
            ce = self.c_est
            self._trunk_fwd(ce, obs)
            ops.row_loss(ce.out, obs[:, p:p + e], self.d_est, alg._aux_loss[1:2], 0)
            self._head_bwd(ce, self.d_est)
            self._trunk_bwd(ce, obs, 0, p)
            # single rank: the estimator's own optimiser step (gail.py:361-365) needs nothing from the other chains -- it runs
            # here, behind the estimator's backward, instead of at the end of the step where nothing else is left to overlap it
            self.est_stepped_in_chain = step_estimator and self._cuda and alg.world_size == 1
            if self.est_stepped_in_chain:
                alg.optim_estimator.step(1.0)
        # ---- critic forward (:343) ----------------------------------------------------------------------------------------
        with on(self.s_critic):
            self._trunk_fwd(self.c_critic, s["critic_obs"])
        # ---- privileged-latent encoder -> actor (:338-342), regulariser (:352-354) ------------------------------------------
        cp, ca, cc = self.c_priv, self.c_actor, self.c_critic
        self._trunk_fwd(cp, s["lat_in"], last_y=xa, last_y_col0=p + e)                 # 29 outputs land in xa[:, 61:90]
        plat = xa[:, p + e:p + e + l]
        ops.row_loss(plat, s["hist_latent"], self.d_reg, alg._aux_loss[0:1], 1)
        self._trunk_fwd(ca, xa)
        cur.join(self.s_critic)
        cfg = self.loss_cfg()
        ops.ppo_loss(ca.out, ac.std.detach(), cc.out, s["actions"], s["old_actions_log_prob"].view(-1), s["advantages"].view(-1),
                     s["returns"].view(-1), s["values"].view(-1), s["old_mu"], s["old_sigma"], self.dmu, self.dvalue,
                     ac.std.grad, alg._ppo_stats, cfg["clip"], cfg["c_surr"], cfg["c_value"], cfg["c_bound"], cfg["c_entropy"],
                     cfg["clipped_value"])
        # ---- critic backward --------------------------------------------------------------------------------------------------
        cur.fork(self.s_critic)
        with on(self.s_critic):
            self._head_bwd(cc, self.dvalue)
            self._trunk_bwd(cc, s["critic_obs"], 0, self.Wc)
        # ---- actor backward, then the encoder's (two upstream gradients: actor input lanes + coef * regulariser) ----------------
        self._head_bwd(ca, self.dmu)
        n = len(ca.trunk)
        for i in range(n - 1, 0, -1):
            ops.linear_bwd(ca.gz[i], None, ca.trunk[i].weight, dx=ca.gz[i - 1], act_prev=ca.act, y_prev=ca.h[i - 1],
                           db_prev=ca.trunk[i - 1].bias.grad, db_accumulate=True)
        first = ca.trunk[0]
        ops.linear_bwd(ca.gz[0], None, first.weight, dx=self.dplat_win, w_col0=self.w_lo, K=self.w_n)
        last_p = cp.trunk[-1]
        ops.act_bwd(self.dplat, plat, cp.act, gz=cp.gz[-1], db=last_p.bias.grad, zero_db=False, addend=self.d_reg,
                    addend_scale=alg._priv_reg_coef)
        m = len(cp.trunk)
        for i in range(m - 1, 0, -1):
            ops.linear_bwd(cp.gz[i], None, cp.trunk[i].weight, dx=cp.gz[i - 1], act_prev=cp.act, y_prev=cp.h[i - 1],
                           db_prev=cp.trunk[i - 1].bias.grad, db_accumulate=True)
        for i in range(n - 1, 0, -1):
            ops.linear_bwd(ca.gz[i], ca.h[i - 1], None, dw=ca.trunk[i].weight.grad)
        ops.linear_bwd(ca.gz[0], xa, None, dw=first.weight.grad, K=self.n_in)
        for i in range(m - 1, 0, -1):
            ops.linear_bwd(cp.gz[i], cp.h[i - 1], None, dw=cp.trunk[i].weight.grad)
        ops.linear_bwd(cp.gz[0], s["lat_in"], None, dw=cp.trunk[0].weight.grad, K=l)
        cur.join(self.s_critic)
        cur.join(self.s_est)
        # kl for the adaptive schedule (:367-373): a 4-byte device copy (it lives in the all-reduce arena when sharded)
        ops.copy_(alg._kl.view(1), alg._ppo_stats[3:4])

    def loss_cfg(self):
        a = self.alg
        return dict(clip=a.clip_param, c_surr=a.surrogate_loss_coef, c_value=a.value_loss_coef, c_bound=a.bounds_loss_coef,
                    c_entropy=a.entropy_coef, clipped_value=a.use_clipped_value_loss)

    def step(self, k: int) -> None:
        self.forward_backward(k, step_estimator=True)
        self.alg._apply()
