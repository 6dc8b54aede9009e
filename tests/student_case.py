"""The depth-student case shared by `oracle/gen_golden_student.py` (run on the reference's modules AND on this package's) and
`tests/test_tsc_student.py`: T recurrent student steps with the BYOL augmentation active, then one `update_depth_actor`
(tsc/rsl_rl/runners/on_policy_runner.py:320-403, algorithms/ppo.py:327-358).  Only attributes both implementations share are
touched (`depth_encoder`, `depth_actor`, `byol_learner.augment1[i].fn / .p`, `update_depth_actor`)."""
import random

import torch

P, A, Y, L = 65, 8, 2, 32


def student_rollout_and_update(alg, inputs, aug_p, seed, stride=997):
    enc, actor = alg.depth_encoder, alg.depth_actor
    applied = [0]

    class CountedModule(torch.nn.Module):                                      # T.GaussianBlur is a registered child module
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x):
            applied[0] += 1
            return self.inner(x)

    def counted(fn):
        if isinstance(fn, torch.nn.Module):
            return CountedModule(fn)

        def wrapper(x):
            applied[0] += 1
            return fn(x)
        return wrapper

    saved = []
    for m in enc.byol_learner.augment1:
        saved.append((m, m.fn, m.p))
        m.fn, m.p = counted(m.fn), aug_p
    random.seed(seed)
    torch.manual_seed(seed)
    cat = student_forward(alg, inputs)
    dev = next(actor.parameters()).device
    stats = alg.update_depth_actor(cat["student"], inputs["actions_teacher"].to(dev), cat["yaw_s"], cat["yaw_t"], cat["obst_s"],
                                   cat["obst_t"], cat["depth"])
    enc.detach_hidden_states()
    for m, fn, p in saved:
        m.fn, m.p = fn, p
    flat = lambda mod: torch.cat([p.detach().reshape(-1) for p in mod.parameters()])[::stride].cpu().clone()   # noqa: E731
    return {"encoder_out": cat["out"].detach().cpu(), "student_actions": cat["student"].detach().cpu(),
            "stats": torch.tensor(stats, dtype=torch.float64), "hidden": enc.hidden_states.detach().cpu().clone(),
            "actor_params": flat(actor), "encoder_params": flat(enc), "target_params": flat(enc.byol_learner.target_encoder),
            "n_aug_applied": applied[0]}


def shard_inputs(inputs, lo, hi):
    """The env shard [lo, hi) of a `synthetic.make_student_inputs` dict."""
    T, N = inputs["obs"].shape[:2]
    teacher = inputs["actions_teacher"].reshape(T, N, -1)[:, lo:hi].flatten(0, 1)
    return dict(obs=inputs["obs"][:, lo:hi], depth=inputs["depth"][:, lo:hi], delta_yaw_ok=inputs["delta_yaw_ok"][:, lo:hi],
                actions_teacher=teacher)


def student_forward(alg, inputs):
    """T recurrent student steps (on_policy_runner.py:320-371) from a fresh GRU state; returns the concatenated buffers."""
    enc, actor = alg.depth_encoder, alg.depth_actor
    enc.train()
    actor.train()
    enc.hidden_states = None
    T = inputs["obs"].shape[0]
    dev = next(actor.parameters()).device
    buf = {k: [] for k in ("out", "student", "yaw_s", "yaw_t", "obst_s", "obst_t", "depth")}
    for t in range(T):
        obs, depth, ok = inputs["obs"][t].to(dev), inputs["depth"][t].to(dev), inputs["delta_yaw_ok"][t].to(dev)
        prop = obs[:, :P].clone()
        prop[:, P - A:P] = 0
        out = enc(depth.clone(), prop)
        depth_latent, delta_yaw, obst = out[:, :L], 1.5 * out[:, L:L + Y], out[:, L + Y:]
        obs_student = obs.clone()
        obs_student[ok, P - A:P - A + Y] = delta_yaw.detach()[ok]
        obs_student[:, P - A + Y:P] = torch.nn.functional.one_hot(torch.argmax(obst.detach(), dim=-1), num_classes=obst.shape[-1]).float()
        emb = actor(obs_student, hist_encoding=True, scandots_latent=depth_latent)
        buf["student"].append(torch.cat([actor.actor_d(emb), actor.actor_c(emb)], dim=-1))
        buf["out"].append(out)
        buf["yaw_s"].append(delta_yaw)
        buf["yaw_t"].append(obs[:, P - A:P - A + Y])
        buf["obst_s"].append(obst)
        buf["obst_t"].append(obs[:, P - A + Y:P])
        buf["depth"].append(depth.clone())
    return {k: torch.cat(v, dim=0) for k, v in buf.items()}


def byol_forward_backward(alg, inputs, aug_p, seed):
    """One BYOL loss + backward on a fixed batch with the augmentation active (no optimiser step: raw gradients are robust to
    summation order, unlike post-Adam parameters)."""
    enc = alg.depth_encoder
    learner = enc.byol_learner
    saved = [(m, m.p) for m in learner.augment1]
    for m, _ in saved:
        m.p = aug_p
    random.seed(seed + 11)
    torch.manual_seed(seed + 11)
    enc.train()
    dev = next(enc.parameters()).device
    x = inputs["depth"].flatten(0, 1).to(dev)
    for p in enc.parameters():
        p.grad = None
    loss = learner(x)
    loss.backward()
    for m, p in saved:
        m.p = p
    named = dict(enc.named_parameters())
    grads = {k: named[k].grad.detach().cpu().clone() for k in ("byol_learner.online_predictor.3.weight",
                                                               "byol_learner.online_encoder.projector.0.weight",
                                                               "base_backbone.image_compression.0.weight",
                                                               "base_backbone.image_compression.8.weight")}
    for p in enc.parameters():
        p.grad = None
    out = {"byol.loss": loss.detach().cpu().reshape(1)}
    out.update({"byol.grad." + k: v for k, v in grads.items()})
    return out
