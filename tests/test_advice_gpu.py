"""Device regressions for the round-1 advisor findings: state that changes between iterations must reach kernels and captured
CUDA graphs through DEVICE memory, never through values frozen at capture / construction time."""
import types

import pytest
import torch

from helpers import mocap_table
from qa_b200 import synthetic
from qa_b200.config import BbcEnvConfig

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_skill_prior_update_reaches_the_fused_step_also_under_graph_replay():
    """`update_ss_info_gail` moves `env.prior_parameters` in place (gail.py:462-464) and `_resample_latent_c` re-derives
    softmax(prior / T) on every call (legged_robot.py:536-539): after `refresh_prior()` the in-kernel mode draw must follow the
    NEW prior -- eagerly and when the step is a replayed CUDA graph captured under the OLD prior."""
    from test_env_gpu import make_env
    cfg = BbcEnvConfig(num_envs=4096)
    static = synthetic.make_static(cfg, seed=21)
    snap = synthetic.make_snapshot(cfg, seed=21, step=0)
    env = make_env(cfg, static, snap, None, 10, table=mocap_table())
    env.use_device_step_counter(True)

    def step_all_reset():
        env.episode_length_buf.fill_(int(env.max_episode_length) + 1)          # every env takes the reset branch
        env.post_physics_step()

    step_all_reset()                                                           # warm-up outside the capture
    torch.cuda.synchronize()
    freq = lambda: env.latent_c.mean(dim=0).cpu()                              # noqa: E731
    uniform = freq()
    assert float((uniform - 0.2).abs().max()) < 0.04                           # initial prior is uniform (:822)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step_all_reset()
    new_prior = torch.tensor([0.5, 0.05, 0.15, 0.05, 0.25], device=DEV)
    env.prior_parameters.copy_(new_prior)                                      # what the trainer does, in place
    env.refresh_prior()
    want = torch.softmax(new_prior.double() / cfg.latent_c_temperature, 0).float().cpu()
    assert torch.allclose(env.prior_prob.cpu(), want, atol=1e-6)
    step_all_reset()
    torch.cuda.synchronize()
    assert float((freq() - want).abs().max()) < 0.04, (freq(), want)
    env.prior_parameters.copy_(torch.tensor([0.05, 0.5, 0.05, 0.35, 0.05], device=DEV))
    env.refresh_prior()
    want2 = env.prior_prob.cpu()
    g.replay()
    torch.cuda.synchronize()
    assert float((freq() - want2).abs().max()) < 0.04, (freq(), want2)
    assert float((want2 - want).abs().max()) > 0.3                             # the two priors are far apart: the test can fail


def test_discriminator_graph_follows_task_obs_weight_decay():
    """`env.task_obs_weight` decays every iteration (on_policy_runner.py:224-225) and scales discriminator-input lanes in the
    update (gail.py:425-432): a captured discriminator step must use the CURRENT weight.  Graph and eager runs through two
    updates with different weights must agree; and the second update must differ from one at the first weight."""
    from test_trainer_gpu import build

    def run(graph, weights):
        alg, env, norm = build(synthetic.make_weights(3), n_envs=64)
        alg.use_cuda_graph = graph
        env.prior_parameters = torch.full((5,), 0.2, device=DEV)
        gen = torch.Generator().manual_seed(4)
        alg.disc_storage.insert(torch.randn(900, 98, generator=gen).to(DEV), torch.rand(900, 1, generator=gen).to(DEV),
                                torch.nn.functional.one_hot(torch.randint(0, 5, (900,), generator=gen), 5).float().to(DEV))
        expert = types.SimpleNamespace(preloaded_s_lb=torch.randn(500, 98, generator=gen).to(DEV),
                                       preloaded_label=torch.randint(0, 5, (500,), generator=gen).to(DEV),
                                       preloaded_s_ulb=torch.randn(700, 98, generator=gen).to(DEV))
        out = []
        for i, wgt in enumerate(weights):
            env.task_obs_weight = wgt
            torch.manual_seed(11 + i)
            out.append(alg.update_disc(expert, num_updates=2))
        return out

    eager = run(False, (1.0, 0.25))
    graph = run(True, (1.0, 0.25))
    frozen = run(False, (1.0, 1.0))
    for a, b in zip(eager[1], graph[1]):
        assert abs(a - b) <= 2e-3 * abs(a) + 1e-5, (eager[1], graph[1])
    assert max(abs(a - b) / (abs(a) + 1e-6) for a, b in zip(eager[1][:5], frozen[1][:5])) > 1e-2   # the weight matters
