"""RolloutStorage with the reference's API (bbc/rsl_rl/storage/rollout_storage.py:7-157) on persistent
device buffers: `add_transitions` copies, `compute_returns` is ONE call of the GAE warp-scan kernel (K5)
instead of a 24-iteration Python loop, `mini_batch_generator` draws one permutation per update and reuses
the same index slices in every epoch (:125, :140-145)."""
import torch

from .. import ops


class RolloutStorage:
    class Transition:
        def __init__(self, num_envs=0, obs_shape=(0,), privileged_obs_shape=(0,), actions_shape=(0,), device='cpu'):
            z = lambda *s, **k: torch.zeros(*s, device=device, **k)            # noqa: E731
            self.observations = z(num_envs, *obs_shape)
            self.critic_observations = z(num_envs, *privileged_obs_shape)
            self.actions = z(num_envs, *actions_shape)
            self.rewards = z(num_envs)
            self.dones = z(num_envs, dtype=torch.uint8)
            self.values = z(num_envs, 1)
            self.actions_log_prob = z(num_envs)
            self.action_mean = z(num_envs, *actions_shape)
            self.action_sigma = z(num_envs, *actions_shape)
            self.hidden_states = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape,
                 device='cpu'):
        self.device = device
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        T, N = num_transitions_per_env, num_envs
        z = lambda *s, **k: torch.zeros(*s, device=device, **k)                # noqa: E731
        # observation rows live at a 16-byte row pitch (671 -> 672 floats): every slot `observations[t]` is then a legal TMA
        # operand, so the critic reads its input straight out of the storage slot and K6 gathers with a source pitch; the public
        # tensors keep the reference's shapes (T, N, width) as views
        rows = lambda shape: z(T, N, (shape[0] + 3) // 4 * 4)[..., :shape[0]] if len(shape) == 1 else z(T, N, *shape)   # noqa: E731
        self.observations = rows(obs_shape)
        self.privileged_observations = rows(privileged_obs_shape) if privileged_obs_shape[0] is not None else None
        self.rewards = z(T, N, 1)
        self.actions = z(T, N, *actions_shape)
        self.dones = z(T, N, 1, dtype=torch.uint8)
        self.actions_log_prob = z(T, N, 1)
        self.values = z(T, N, 1)
        self.returns = z(T, N, 1)
        self.advantages = z(T, N, 1)
        self.mu = z(T, N, *actions_shape)
        self.sigma = z(T, N, *actions_shape)
        self.num_transitions_per_env, self.num_envs = T, N
        self.saved_hidden_states_a = self.saved_hidden_states_c = None
        self.step = 0
        self._gae_ws = torch.zeros(8, device=device, dtype=torch.float64) if torch.device(device).type == "cuda" else None

    def add_transitions(self, transition: "RolloutStorage.Transition", rewards_dones_written: bool = False):
        """`rewards_dones_written`: rewards[step] / dones[step] were already written in place (fused reward kernel K19)."""
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        self.observations[s].copy_(transition.observations)
        if self.privileged_observations is not None:
            self.privileged_observations[s].copy_(transition.critic_observations)
        self.actions[s].copy_(transition.actions)
        if not rewards_dones_written:
            self.rewards[s].copy_(transition.rewards.view(-1, 1))
            self.dones[s].copy_(transition.dones.view(-1, 1))
        self.values[s].copy_(transition.values)
        self.actions_log_prob[s].copy_(transition.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(transition.action_mean)
        self.sigma[s].copy_(transition.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def advance(self):
        """Close a transition whose fields were written straight into slot `step` by the producing kernels (rollout_plan)."""
        self.step += 1

    def compute_returns(self, last_values, gamma, lam):
        """GAE + advantage normalisation (rollout_storage.py:97-111) through libqa_b200 (K5)."""
        if not self.rewards.is_cuda:
            # host construction (OnPolicyRunner's default device is 'cpu', as in the reference) is for plumbing and tests
            # only: the product path is the CUDA kernel and says so instead of failing deep inside ops.gae
            raise RuntimeError("RolloutStorage.compute_returns: qa_b200 computes GAE on the GPU only (K5, libqa_b200); "
                               "construct the runner / storage with a CUDA device")
        ops.gae(self.rewards, self.values, self.dones, last_values.contiguous(), self.returns, self.advantages,
                self._gae_ws, gamma, lam)

    def get_statistics(self):
        done = self.dones
        done[-1] = 1
        flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
        done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero(as_tuple=False)[:, 0]))
        trajectory_lengths = (done_indices[1:] - done_indices[:-1])
        return trajectory_lengths.float().mean(), self.rewards.mean()

    def flat_views(self):
        f = lambda t: t.flatten(0, 1)                                           # noqa: E731
        crit = self.privileged_observations if self.privileged_observations is not None else self.observations
        return dict(obs=f(self.observations), critic_obs=f(crit), actions=f(self.actions), values=f(self.values),
                    returns=f(self.returns), old_actions_log_prob=f(self.actions_log_prob),
                    advantages=f(self.advantages), old_mu=f(self.mu), old_sigma=f(self.sigma))

    def mini_batch_generator(self, num_mini_batches, num_epochs=8, indices=None):
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        if indices is None:
            indices = torch.randperm(num_mini_batches * mini_batch_size, requires_grad=False, device=self.device)
        v = self.flat_views()
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                idx = indices[i * mini_batch_size:(i + 1) * mini_batch_size]
                yield (v["obs"][idx], v["critic_obs"][idx], v["actions"][idx], v["values"][idx], v["advantages"][idx],
                       v["returns"][idx], v["old_actions_log_prob"][idx], v["old_mu"][idx], v["old_sigma"][idx],
                       (None, None), None)
