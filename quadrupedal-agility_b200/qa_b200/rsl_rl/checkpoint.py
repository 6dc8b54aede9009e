"""Checkpoint codec: flat Adam state <-> the `torch.optim.Adam.state_dict()` layout the reference writes.

The reference's checkpoints (bbc/rsl_rl/runners/on_policy_runner.py:306-339, the shipped `tsc/weights/bbc/model.pt`) hold
six `torch.optim.Adam` state dicts: `{'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [{..., 'params':
[i, ...]}]}` with one entry per parameter in the order the parameters were handed to the optimiser (gail.py:96-128).
Here the moments live in one flat fp32 buffer per optimiser slice (`FlatAdam`, rows of 2-D weights padded to 16 B), so a
checkpoint written by this package must be re-shaped on the way out and on the way in.  Both directions are plain tensor
slicing (no kernel), run once per save/load.

`groups` below is a list of dicts, one per reference param group:
    {"adam": FlatAdam, "params": [(offset, padded_count, shape), ...], "extra": {"name": "trunk", ...}}
with `offset` in the coordinates of the FlatParams buffer the FlatAdam slices.
"""
import torch

_ADAM_DEFAULTS = dict(amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False, fused=None)


def layout(flat, module, prefix=""):
    """[(offset, padded_count, shape)] of `module`'s parameters (in `module.parameters()` order) inside `flat`, whose slice
    names carry `prefix` (e.g. "history_encoder." when `flat` was built over the enclosing ActorCritic)."""
    out = []
    for name, p in module.named_parameters():
        off, cnt = flat.slices[prefix + name]
        out.append((off, cnt, tuple(p.shape)))
    return out


def _unpad(buf, cnt, shape):
    """A padded flat slice -> a dense tensor of `shape` (2-D weights: row pitch cnt/rows; others: leading numel)."""
    if len(shape) == 2:
        rows, k = shape
        return buf.view(rows, cnt // rows)[:, :k].clone()
    n = 1
    for s in shape:
        n *= s
    return buf[:n].view(shape).clone()


def _pad_into(buf, cnt, shape, value):
    buf.zero_()
    if len(shape) == 2:
        rows, k = shape
        buf.view(rows, cnt // rows)[:, :k].copy_(value)
    else:
        buf[:value.numel()].copy_(value.reshape(-1))


def to_torch_state_dict(groups):
    """`torch.optim.Adam.state_dict()` of the optimiser the reference would hold (same keys, same parameter numbering)."""
    state, param_groups, idx = {}, [], 0
    for g in groups:
        adam = g["adam"]
        step = adam.step_count.detach().to("cpu", torch.float32).reshape(())
        ids = []
        for off, cnt, shape in g["params"]:
            lo = off - adam.lo
            state[idx] = {"step": step.clone(),
                          "exp_avg": _unpad(adam.exp_avg[lo:lo + cnt], cnt, shape),
                          "exp_avg_sq": _unpad(adam.exp_avg_sq[lo:lo + cnt], cnt, shape)}
            ids.append(idx)
            idx += 1
        pg = dict(g.get("extra", {}))
        pg.update(lr=float(adam.lr.item()), betas=tuple(adam.betas), eps=adam.eps)
        pg.setdefault("weight_decay", adam.weight_decay)
        for k, v in _ADAM_DEFAULTS.items():
            pg.setdefault(k, v)
        pg["params"] = ids
        param_groups.append(pg)
    return {"state": state, "param_groups": param_groups}


def is_torch_state_dict(sd) -> bool:
    return isinstance(sd, dict) and "state" in sd and "param_groups" in sd


def from_torch_state_dict(sd, groups, load_lr=True):
    """Inverse of `to_torch_state_dict`.  Accepts the reference's dicts from any torch version (`step` int or tensor);
    parameters without an entry (never stepped, e.g. the history encoder inside `optim_ac`) get zero moments.  When two
    groups alias one FlatAdam slice only the matching parameters are touched."""
    if len(sd["param_groups"]) != len(groups):
        raise ValueError(f"optimizer state has {len(sd['param_groups'])} param groups, expected {len(groups)}")
    for pg, g in zip(sd["param_groups"], groups):
        adam = g["adam"]
        if len(pg["params"]) != len(g["params"]):
            raise ValueError(f"param group has {len(pg['params'])} parameters, expected {len(g['params'])}")
        step = 0
        for pid, (off, cnt, shape) in zip(pg["params"], g["params"]):
            lo = off - adam.lo
            st = sd["state"].get(pid)
            for key, buf in (("exp_avg", adam.exp_avg), ("exp_avg_sq", adam.exp_avg_sq)):
                dst = buf[lo:lo + cnt]
                if st is None:
                    dst.zero_()
                    continue
                v = st[key]
                if tuple(v.shape) != tuple(shape):
                    raise ValueError(f"optimizer state {pid}.{key} has shape {tuple(v.shape)}, expected {tuple(shape)}")
                _pad_into(dst, cnt, shape, v.to(dst.device, torch.float32))
            if st is not None:
                s = st["step"]
                step = max(step, int(s.item()) if torch.is_tensor(s) else int(s))
        adam.step_count.fill_(step)
        if load_lr:
            adam.lr.fill_(float(pg["lr"]))


def zero_torch_state_dict(groups, device="cpu"):
    """The dict of an Adam optimiser that has not stepped yet but whose moments are materialised (what this package writes for
    optimisers it builds lazily).  groups: [{"shapes": [shape, ...], "lr": float, "weight_decay": float, "extra": {...}}]."""
    state, param_groups, idx = {}, [], 0
    for g in groups:
        ids = []
        for shape in g["shapes"]:
            state[idx] = {"step": torch.zeros((), dtype=torch.float32), "exp_avg": torch.zeros(shape, device=device),
                          "exp_avg_sq": torch.zeros(shape, device=device)}
            ids.append(idx)
            idx += 1
        pg = dict(g.get("extra", {}))
        pg.update(lr=float(g["lr"]), betas=(0.9, 0.999), eps=1e-8)
        pg.setdefault("weight_decay", g.get("weight_decay", 0.0))
        for k, v in _ADAM_DEFAULTS.items():
            pg.setdefault(k, v)
        pg["params"] = ids
        param_groups.append(pg)
    return {"state": state, "param_groups": param_groups}
