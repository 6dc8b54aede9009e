"""`DiscStepPlan` -- one discriminator minibatch step of `SSInfoGAIL.update_ss_info_gail` (bbc/rsl_rl/algorithms/gail.py:415-541)
as a static schedule of libqa_b200 launches: no autograd graph, no double-backward through framework kernels, no host round trip.

The three batches [policy | labelled expert | unlabelled expert] (B rows each) go through the trunk 98 -> 512 -> 256 (ReLU) as ONE
(3B, 98) matrix:

  K24 prepare        gathers + task weighting + normalisation (:419-452) -> X, per-row targets
  K7  trunk forward  H1 = relu(X W1^T + b1), H2 = relu(H1 W2^T + b2)                 (tcgen05, TF32 operands)
  K25 heads + losses d / eps / classifier heads, CE-on-softmax, info-max, LSGAN, L1, prior estimate, accuracies; gradient w.r.t.
                     the trunk output, head parameter gradients, trunk bias-2 gradient (:454-490, :532-538)
  K7  trunk backward dW2 += gz2^T H1, gz1 = (gz2 W2) relu'(H1) (+ db1), dW1 += gz1^T X
  gradient penalty   (:492-502) on the unlabelled rows.  For a ReLU network dD/dx = W1^T m1 (W2^T m2 w_d) with the activation
                     masks m1, m2, and the second derivative only sees the masks (relu'' = 0):
                       v2 = m2 w_d (K25) ; v1 = m1 (v2 W2) ; g = v1 W1 ; loss = mean ||g||^2 ; dg = 2 c g / B (K27)
                       dW1 += v1^T dg ; dt1 = m1 (dg W1^T) ; dW2 += v2^T dt1 ; dw_d += colsum(m2 (dt1 W2^T))
                     -- six small GEMMs (K7) and two masked reductions (K9) on a side stream, parallel to the main backward
  K28 regularisers   logit regulariser + weight decay (:488-490, :504-507)
  K8  x5             the reference's three Adam optimisers (:519-521; the trunk is stepped by all three)
  K29 / K30          batch moments of the three normalised batches, Chan merge into the running normaliser, prior soft update,
                     policy-std floor (:462-464, :523-529)

The statistics (the reference's 11-tuple) are accumulated on the device; `update_disc` reads them back once per update.
"""
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from .ppo_plan import _Cur, _linears, _padded


class DiscStepPlan:
    @staticmethod
    def supported(alg) -> Optional[str]:
        d = alg.disc
        if torch.device(alg.device).type != "cuda":
            return "not a CUDA device"
        if d is None or alg.disc_loss_function != "MSELoss":
            return "only the MSE (LSGAN) discriminator loss is scheduled"
        lins = _linears(d.trunk)
        if lins is None or len(lins) != 2 or not isinstance(d.trunk[1], nn.ReLU) or lins[1].out_features != 256:
            return "trunk is not Linear-ReLU-Linear(256)-ReLU"
        if d.classifier.out_features != 5 or d.linear.out_features != 1 or d.encoder_eps.out_features != 1:
            return "head shapes"
        if alg.disc_normalizer is None:
            return "no normaliser"
        if alg.disc_batched:
            return "QA_DISC_BATCHED selects the autograd variant"
        return None

    def __init__(self, alg, B: int):
        self.alg, self.B = alg, B
        dev = self.dev = torch.device(alg.device)
        d = alg.disc
        self.lin1, self.lin2 = d.trunk[0], d.trunk[2]
        W = self.width = self.lin1.in_features
        H1, H2 = self.lin1.out_features, self.lin2.out_features
        self.x = _padded(3 * B, W, dev)
        self.h1, self.h2 = _padded(3 * B, H1, dev), _padded(3 * B, H2, dev)
        self.gz2, self.gz1 = _padded(3 * B, H2, dev), _padded(3 * B, H1, dev)
        self.v2, self.v1 = _padded(B, H2, dev), _padded(B, H1, dev)
        self.g = _padded(B, W, dev)
        self.dv1, self.dt1, self.dv2 = _padded(B, H1, dev), _padded(B, H1, dev), _padded(B, H2, dev)
        self.tgt_eps = torch.zeros(B, device=dev)
        self.tgt_c = torch.zeros(B, device=dev, dtype=torch.int32)
        self.tgt_label = torch.zeros(B, device=dev, dtype=torch.int32)
        self.moments = torch.zeros(3, 2, W, device=dev, dtype=torch.float64)
        self.prior_batch = torch.zeros(8, device=dev)
        self._cuda = dev.type == "cuda"            # (the host tests drive the schedule on CPU through stand-in ops: no streams)
        self.s_gp = torch.cuda.Stream(device=dev) if self._cuda else None
        self.s_gw = torch.cuda.Stream(device=dev) if self._cuda else None      # the penalty's two weight-gradient GEMMs
        sl = alg.disc_flat.slices
        self.reg_segments = [sl["trunk.0.weight"], sl["trunk.2.weight"], sl["linear.weight"]]
        self.obs_dim = alg.num_disc_obs

    def step(self, expert, i_pi, i_lb, i_ulb) -> None:
        """One minibatch step on the index vectors (B,) into the replay buffer / expert sets; statistics are ADDED to
        `alg._disc_stats`."""
        alg, B, d = self.alg, self.B, self.alg.disc
        env, norm = alg.env, alg.disc_normalizer
        _, mean64, var64, count, mean32, std32 = norm._device_state(self.dev)
        flat = alg.disc_flat
        stats = alg._disc_stats
        import contextlib
        cur = _Cur(self._cuda)
        on = (lambda st: torch.cuda.stream(st)) if self._cuda else (lambda st: contextlib.nullcontext())
        ops.zero_(flat.grad)
        ops.zero_(self.prior_batch)
        ops.disc_prepare(B, alg.disc_storage, expert, i_pi, i_lb, i_ulb, env.task_obs_weight_decay,
                         alg._task_obs_weight_dev() if env.task_obs_weight_decay else None, alg.obs_disc_weight_step,
                         mean32, std32, norm.clip_obs, self.x, self.tgt_eps, self.tgt_c, self.tgt_label, self.obs_dim)
        l1, l2 = self.lin1, self.lin2
        ops.linear_fwd(self.x, l1.weight, l1.bias, self.h1, "relu")
        ops.linear_fwd(self.h1, l2.weight, l2.bias, self.h2, "relu")
        ops.disc_heads_loss(B, self.h2, d, self.tgt_eps, self.tgt_c, self.tgt_label, alg.ss_coef, alg.disc_coef, alg.us_coef,
                            alg._info_max_coef_on, self.gz2, self.v2, stats, self.prior_batch)
        # ---- gradient penalty on the unlabelled rows (side stream) -----------------------------------------------------------
        cur.fork(self.s_gp)
        with on(self.s_gp):
            h1u, h2u = self.h1[2 * B:], self.h2[2 * B:]
            ops.linear_bwd(self.v2, None, l2.weight, dx=self.v1, act_prev="relu", y_prev=h1u)          # v1 = m1 (v2 W2)
            ops.linear_bwd(self.v1, None, l1.weight, dx=self.g)                                        # g = v1 W1
            ops.disc_gp_loss(self.g, alg.disc_grad_penalty, stats)                                     # g := d loss / d g
            gp = _Cur(self._cuda)                                                                      # (current stream = s_gp)
            # the two weight-gradient GEMMs are leaves of the chain (split-K reduce-adds into the flat gradient buffer): they
            # run on a third stream next to the dv1 -> dt1 -> dv2 chain, which is the critical path of the whole step
            gp.fork(self.s_gw)
            with on(self.s_gw):
                ops.linear_bwd(self.v1, self.g, None, dw=l1.weight.grad)                               # dW1 += v1^T dg
            ops.linear_fwd(self.g, l1.weight, None, self.dv1, None)                                    # dv1 = dg W1^T
            ops.act_bwd(self.dv1, h1u, "relu", gz=self.dt1)                                            # dt1 = m1 dv1
            gp.fork(self.s_gw)
            with on(self.s_gw):
                ops.linear_bwd(self.v2, self.dt1, None, dw=l2.weight.grad)                             # dW2 += v2^T dt1
            ops.linear_fwd(self.dt1, l2.weight, None, self.dv2, None)                                  # dv2 = dt1 W2^T
            ops.act_bwd(self.dv2, h2u, "relu", gz=None, db=d.linear.weight.grad.view(-1), zero_db=False)   # dw_d += sum m2 dv2
            gp.join(self.s_gw)
        # ---- main backward -----------------------------------------------------------------------------------------------------
        ops.linear_bwd(self.gz2, None, l2.weight, dx=self.gz1, act_prev="relu", y_prev=self.h1, db_prev=l1.bias.grad,
                       db_accumulate=True)
        ops.linear_bwd(self.gz2, self.h1, None, dw=l2.weight.grad)
        ops.linear_bwd(self.gz1, self.x, None, dw=l1.weight.grad)
        cur.join(self.s_gp)
        ops.disc_reg(flat, self.reg_segments, alg.disc_logit_reg, alg.disc_weight_decay, stats)
        # ---- optimisers, normaliser, prior, std floor ------------------------------------------------------------------------
        from .. import dist as qdist
        scale = qdist.allreduce_flat_(flat.grad)                                  # env shards: ONE all-reduce, 1/W folded into K8
        alg._disc_optim_step(scale)
        ops.norm_moments(self.x, B, 3, self.moments)
        W = alg.world_size
        if W > 1:
            import torch.distributed as tdist
            tdist.all_reduce(self.moments, op=tdist.ReduceOp.SUM)
            tdist.all_reduce(self.prior_batch, op=tdist.ReduceOp.SUM)
        ac = alg.actor_critic
        floor = (not ac.fixed_std) and alg.min_std is not None
        ops.norm_merge(B, 3, W, self.moments, mean64, var64, count, mean32, std32, norm.epsilon,
                       prior=env.prior_parameters, prior_batch=self.prior_batch, prior_soft_coef=alg.prior_soft_coef,
                       std=ac.std.data if floor else None, min_std=alg.min_std.contiguous() if floor else None)
        norm.__dict__["_dev_dirty"] = True
