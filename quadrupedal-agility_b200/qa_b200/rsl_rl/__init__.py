"""Trainer half of the hot path: the reference's `rsl_rl` API (RolloutStorage, ActorCritic, Estimator,
Discriminator, SSInfoGAIL, OnPolicyRunner) over flat parameter buffers and the libqa_b200 kernels."""
from .modules import ActorCritic, ActorCriticBBC, Estimator, Discriminator, DiscriminatorTSC, StateHistoryEncoder  # noqa: F401
from .storage import RolloutStorage  # noqa: F401
from .utils import Normalizer, RunningMeanStd  # noqa: F401
from .algorithm import SSInfoGAIL  # noqa: F401
from .tsc import Actor, ActorCriticTSC, RolloutStorageTSC, PPO  # noqa: F401
