"""The env interface the runners are typed against (bbc/rsl_rl/env/vec_env.py:7-36, tsc/rsl_rl/env/vec_env.py:6-30): what
`OnPolicyRunner` / `OnPolicyRunnerTSC` rely on.  `LeggedRobot` and `LeggedRobotTSC` are registered as virtual subclasses;
`missing_members(env)` names what an object lacks of the contract (used by the tests, and handy when wrapping another env)."""
from abc import ABC, abstractmethod
from typing import Optional

import torch

ATTRIBUTES = ("num_envs", "num_obs", "num_privileged_obs", "num_actions", "max_episode_length", "privileged_obs_buf", "obs_buf",
              "rew_buf", "reset_buf", "episode_length_buf", "extras", "device")
METHODS = ("step", "reset", "get_observations", "get_privileged_observations")


class VecEnv(ABC):
    num_envs: int
    num_obs: int
    num_privileged_obs: Optional[int]
    num_actions: int
    max_episode_length: float
    privileged_obs_buf: Optional[torch.Tensor]
    obs_buf: torch.Tensor
    rew_buf: torch.Tensor
    reset_buf: torch.Tensor
    episode_length_buf: torch.Tensor
    extras: dict
    device: torch.device

    @abstractmethod
    def step(self, actions: torch.Tensor):
        """-> (obs, privileged_obs | None, rewards, dones, extras, reset_env_ids, terminal_disc_states) in this repository's fork."""

    @abstractmethod
    def reset(self):
        """-> (obs, privileged_obs)"""

    @abstractmethod
    def get_observations(self) -> torch.Tensor: ...

    @abstractmethod
    def get_privileged_observations(self) -> Optional[torch.Tensor]: ...


def missing_members(env, tsc: bool = False):
    need = list(ATTRIBUTES) + list(METHODS) + (["get_observations_bbc", "get_observations_disc", "set_commands"] if tsc
                                               else ["get_disc_observations"])
    if tsc:
        need.remove("reset")                                        # the TSC runner never calls it (tsc on_policy_runner.py)
    return [n for n in need if not hasattr(env, n)]
