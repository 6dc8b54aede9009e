"""The `env.cfg.*` / `env.obs_scales` names the reference's trainers read off the env (bbc/rsl_rl/runners/
on_policy_runner.py:37-63, tsc/rsl_rl/runners/on_policy_runner.py:33-57, 281-296, bbc/rsl_rl/datasets/motion_loader.py:212-217)."""
from qa_b200.config import BbcEnvConfig
from qa_b200.legged_robot_tsc import TscEnvConfig


def test_bbc_cfg_env_namespace():
    e = BbcEnvConfig(num_envs=16).env
    assert (e.num_prop, e.num_explicit, e.num_latent, e.num_command, e.history_len) == (57, 4, 29, 11, 10)
    assert e.num_obs == e.num_privileged_obs == 101 and e.num_obs_disc == 49
    assert (e.disc_history_len, e.disc_obs_len, e.obs_disc_weight_step, e.frame_duration_scale) == (2, 2, 0.0, 1.0)


def test_tsc_cfg_namespaces():
    c = TscEnvConfig(num_envs=16)
    e = c.env
    assert (e.n_proprio, e.n_delta_yaw, e.n_obst_type, e.n_auxiliary, e.n_scan, e.n_priv, e.n_priv_latent) == (65, 2, 6, 8, 132, 4, 29)
    assert (e.history_len, e.num_command, e.num_observations_bbc, e.num_actions_bbc, e.num_obs_disc, e.disc_obs_len) == (10, 11, 101, 12, 49, 2)
    assert c.domain_rand.action_buf_len == 8 and c.noise.add_noise is False and c.obstacle.curriculum is False
    c.next_goal_threshold = 0.45
    assert c.env.next_goal_threshold == 0.45


def test_envs_implement_the_vec_env_interface():
    """bbc/rsl_rl/env/vec_env.py:7-36, tsc/rsl_rl/env/vec_env.py:6-30."""
    from qa_b200.legged_robot import LeggedRobot
    from qa_b200.legged_robot_tsc import LeggedRobotTSC
    from qa_b200.rsl_rl.vec_env import METHODS, VecEnv, missing_members
    assert issubclass(LeggedRobot, VecEnv) and issubclass(LeggedRobotTSC, VecEnv)
    assert all(callable(getattr(LeggedRobot, m)) for m in METHODS + ("get_disc_observations",))
    assert all(callable(getattr(LeggedRobotTSC, m)) for m in ("step", "get_observations", "get_privileged_observations",
                                                              "get_observations_bbc", "get_observations_disc", "set_commands"))
    import types
    assert "step" in missing_members(types.SimpleNamespace(num_envs=1)) and missing_members(types.SimpleNamespace(), tsc=True)
