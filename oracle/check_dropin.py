"""TEST INFRASTRUCTURE (build container; needs /root/reference).  The drop-in construction path of `qa_b200.dropin`, checked on
the reference env the parity harness builds (`oracle/ref_env.build_reference_env`: the UNMODIFIED `LeggedRobot` class of
bbc/legged_gym/envs/base/legged_robot.py with IsaacGym's tensors injected):

  * `config_from_reference(ref_env)` gives the `BbcEnvConfig` the harness was sized with -- every field equal -- out of the
    reference's own nested `Go2LocomotionCfg` and the env's asset index lists;
  * `static_from_reference(ref_env)` gives back, tensor for tensor, the per-env constants that were injected under the
    reference's attribute names;
  * `make_task_class(RefLeggedRobot)` has the reference constructor's signature (task_registry.py:66-70) and hands the env the
    reference built to `from_reference_env`.
Prints "dropin OK"."""
import dataclasses
import inspect
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "quadrupedal-agility_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from qa_b200 import dropin, synthetic  # noqa: E402
from qa_b200.config import BbcEnvConfig  # noqa: E402
import ref_env as RE  # noqa: E402


def main():
    cfg = BbcEnvConfig(num_envs=64)
    static = synthetic.make_static(cfg, seed=5)
    snap = synthetic.make_snapshot(cfg, seed=5, step=0)
    ref, env = RE.build_reference_env(cfg, static, snap, None)
    # the harness injects noise_scale_vec through the reference's own `_get_noise_scale_vec`, limits / gains from `static`
    got_cfg = dropin.config_from_reference(env)
    for f in dataclasses.fields(BbcEnvConfig):
        a, b = getattr(got_cfg, f.name), getattr(cfg, f.name)
        if f.name in ("dof_pos_lower", "dof_pos_upper"):      # URDF limits: the env carries their soft version (static), see below
            continue
        if isinstance(a, float):
            assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (f.name, a, b)   # torque / velocity limits travel through fp32 tensors
        elif isinstance(a, list) and a and isinstance(a[0], float):
            assert all(abs(x - y) <= 1e-6 * max(1.0, abs(y)) for x, y in zip(a, b)) and len(a) == len(b), (f.name, a, b)
        else:
            assert a == b, (f.name, a, b)
    st = dropin.static_from_reference(env)
    assert set(st) == set(dropin.STATIC_KEYS)
    for k in dropin.STATIC_KEYS:
        want = static[k]
        if k == "noise_scale_vec":                            # the reference's own vector (legged_robot.py:721-740) == this package's
            want = cfg.noise_scale_vec()
        assert st[k].shape == want.shape and torch.equal(st[k].to(want.dtype), want), k
    # a kernel-layout mismatch is refused, not served
    env.cfg.env.history_len = 12
    try:
        dropin.config_from_reference(env)
        raise AssertionError("a 12-slot history must be refused")
    except ValueError:
        env.cfg.env.history_len = 10
    # constructor signature of the registrable class == the reference's
    T = dropin.make_task_class(ref.LeggedRobot)
    ref_params = list(inspect.signature(ref.LeggedRobot.__init__).parameters)[1:]
    assert list(inspect.signature(T.__new__).parameters)[1:] == ref_params, (ref_params,)
    seen = {}

    class FakeRef:                                            # records what the registry-style call hands down
        def __init__(self, **kw):
            seen.update(kw)

    T.REFERENCE_CLASS = FakeRef
    orig = dropin.from_reference_env
    dropin.from_reference_env = lambda e, device=None: ("env over", e, device)
    try:
        out = T(cfg="C", sim_params="S", physics_engine="P", sim_device="cuda:0", headless=True)
    finally:
        dropin.from_reference_env = orig
    assert out[0] == "env over" and isinstance(out[1], FakeRef) and out[2] == "cuda:0"
    assert seen == dict(cfg="C", sim_params="S", physics_engine="P", sim_device="cuda:0", headless=True)
    print("dropin OK")


if __name__ == "__main__":
    main()
