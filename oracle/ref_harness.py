"""Import the UNMODIFIED reference (`/root/reference/{bbc,tsc}`) in this container.

TEST INFRASTRUCTURE ONLY -- used by `oracle/gen_golden.py` to (a) validate the oracle
restatement in `oracle/*.py` against the reference's own classes and (b) generate the
golden vectors committed under `tests/golden/`.  `/root/reference` does not exist on
the GPU box, so nothing in `tests/ -m gpu`, `bench.py` or `smoke()` imports this file.

The reference cannot be installed (isaacgym is a closed binary, python<=3.8); instead
its modules are imported with the stub packages in `oracle/stubs/` shadowing
`isaacgym`, `turtle`, `pybullet_utils` and `matplotlib`.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("QA_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bbc", "legged_gym"))


def import_reference(which: str = "bbc"):
    """Returns a namespace of the reference's hot-path modules for `which` in {bbc, tsc}.

    bbc/ and tsc/ carry their own forks of `legged_gym` and `rsl_rl`, so only one of them
    can be live in `sys.modules` at a time; switching purges the other.
    """
    assert which in ("bbc", "tsc")
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    root = os.path.join(REFERENCE_ROOT, which)
    live = sys.modules.get("legged_gym")
    if live is not None and not os.path.realpath(live.__file__).startswith(os.path.realpath(root)):
        for k in [k for k in sys.modules if k.split(".")[0] in ("legged_gym", "rsl_rl")]:
            del sys.modules[k]
    for p in (root, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # import order matters: legged_gym.envs first, as train.py:9 does
    ns = type("RefNS", (), {})()
    ns.envs = importlib.import_module("legged_gym.envs")
    ns.legged_robot = importlib.import_module("legged_gym.envs.base.legged_robot")
    ns.LeggedRobot = ns.legged_robot.LeggedRobot
    ns.torch_jit_utils = importlib.import_module("legged_gym.utils.torch_jit_utils")
    ns.rollout_storage = importlib.import_module("rsl_rl.storage.rollout_storage")
    ns.RolloutStorage = ns.rollout_storage.RolloutStorage
    ns.actor_critic = importlib.import_module("rsl_rl.modules.actor_critic")
    ns.estimator = importlib.import_module("rsl_rl.modules.estimator")
    ns.utils = importlib.import_module("rsl_rl.utils.utils")
    ns.discriminator = importlib.import_module("rsl_rl.algorithms.discriminator")
    if which == "bbc":
        ns.motion_loader = importlib.import_module("rsl_rl.datasets.motion_loader")
        ns.gail = importlib.import_module("rsl_rl.algorithms.gail")
    else:
        ns.ppo = importlib.import_module("rsl_rl.algorithms.ppo")
    ns.root = root
    return ns
