"""Drop-in boundary, seen from the reference's side: the UNMODIFIED reference trainer code runs over this package's
`ActorCritic` / `Estimator` and gives the results it gives over its own modules.  Needs the reference tree (build container);
runs in a fresh interpreter because the reference's `rsl_rl` package and this package's pickle alias share a module name."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/bbc"), reason="the reference tree only exists in the build container")
def test_reference_ssinfogail_runs_over_this_packages_actor_critic():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "check_interop.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "interop OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/tsc"), reason="the reference tree only exists in the build container")
def test_reference_tsc_ppo_runs_over_this_packages_actor_critic_tsc():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "check_interop.py"), "tsc"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "interop OK (tsc)" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/bbc"), reason="the reference tree only exists in the build container")
def test_dropin_construction_extracts_config_and_constants_from_the_reference_env():
    """`qa_b200.dropin`: BbcEnvConfig + per-env constants out of an env the UNMODIFIED reference class built; the registrable
    task class has the reference constructor's signature (bbc/legged_gym/utils/task_registry.py:66-70)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "check_dropin.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "dropin OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
