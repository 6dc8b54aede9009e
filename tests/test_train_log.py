"""Runner bookkeeping (bbc/rsl_rl/runners/on_policy_runner.py:183-206, :238-304): the device-staged `EpisodeBook` must leave
the six `deque(maxlen=100)`s exactly as the reference's per-step accounting does, and `log_bbc` must emit the reference's tags."""
import statistics
import types
from collections import deque

import numpy as np
import torch

from qa_b200.rsl_rl.train_log import EpisodeBook, ScalarLog, log_bbc


def reference_accounting(terms, dones, n_iters, T):
    """The reference's loop, restated with torch ops as written there (:190-206)."""
    N, C = terms.shape[1], terms.shape[2]
    bufs = [deque(maxlen=100) for _ in range(C)]
    len_buffer = deque(maxlen=100)
    cur = [torch.zeros(N) for _ in range(C)]
    cur_len = torch.zeros(N)
    for s in range(n_iters * T):
        for c in range(C):
            cur[c] += terms[s, :, c]
        cur_len += 1
        new_ids = (dones[s] > 0).nonzero(as_tuple=False)
        for c in range(C):
            bufs[c].extend(cur[c][new_ids][:, 0].cpu().numpy().tolist())
        len_buffer.extend(cur_len[new_ids][:, 0].cpu().numpy().tolist())
        for c in range(C):
            cur[c][new_ids] = 0
        cur_len[new_ids] = 0
    return bufs, len_buffer


def test_episode_book_matches_the_reference_deques():
    g = torch.Generator().manual_seed(0)
    N, T, iters, C = 64, 6, 5, 5
    terms = torch.randn(iters * T, N, C, generator=g)
    dones = (torch.rand(iters * T, N, generator=g) < 0.08).to(torch.uint8)
    dones[7] = 0                                                          # a step without any finished episode
    book = EpisodeBook(N, T, ("total", "i", "us", "ss", "t"), "cpu", num_episode_keys=3)
    assert book.means() == {}
    for it in range(iters):
        for t in range(T):
            s = it * T + t
            if t % 2:                                                     # both ways of handing the terms over
                book.term_slot().copy_(terms[s])
                book.record(dones[s].bool(), episode_means=torch.full((3,), float(s)))
            else:
                book.record(dones[s], terms[s], torch.full((3,), float(s)))
        book.flush()
        assert [float(r[0]) for r in book.episode_rows] == [float(it * T + t) for t in range(T)]
    bufs, len_buffer = reference_accounting(terms, dones, iters, T)
    for c, name in enumerate(book.names):
        assert list(book.buffers[name]) == list(bufs[c]), name
    assert list(book.len_buffer) == list(len_buffer) and len(len_buffer) == 100
    m = book.means()
    assert m["total"] == statistics.mean(bufs[0]) and m["episode_length"] == statistics.mean(len_buffer)


def test_log_bbc_emits_the_reference_tags():
    names = ["action_rate", "torques"]
    env = types.SimpleNamespace(num_envs=8, reward_names=names, reward_scales={"action_rate": -0.5, "torques": 2.0})
    alg = types.SimpleNamespace(actor_critic=types.SimpleNamespace(std=torch.full((12,), 0.5)), lr_ac=1e-3, lr_disc=2e-3, lr_q=3e-3)
    book = EpisodeBook(8, 2, ("total", "i", "us", "ss", "t"), "cpu", num_episode_keys=2)
    for t in range(2):
        book.record(torch.ones(8), torch.full((8, 5), 1.0 + t), torch.tensor([1.0, 4.0]))
    book.flush()
    runner = types.SimpleNamespace(env=env, alg=alg, book=book, num_steps_per_env=2)
    log = ScalarLog()
    stats = tuple(float(i) for i in range(17))
    m = log_bbc(runner, log, 3, stats, 0.125, 0.5, 1.5)
    want = {"Episode/rew_action_rate": -2.0, "Episode/rew_torques": 2.0, "Loss/surrogate_loss": 0.0, "Loss/value_loss": 1.0,
            "Loss/b_loss": 2.0, "Loss/entropy_batch": 3.0, "Loss/priv_reg_loss": 4.0, "Loss/estimator_loss": 5.0,
            "Loss/hist_latent_loss": 0.125, "Loss/ss_loss": 6.0, "Loss/info_max_loss": 7.0, "Loss/disc_loss": 8.0,
            "Loss/us_loss": 9.0, "Loss/grad_pen_loss": 10.0, "Loss/disc_logit_loss": 11.0, "Loss/disc_weight_decay": 12.0,
            "Acc/acc_lb": 13.0, "Acc/acc_pi": 14.0, "Acc/acc_exp": 15.0, "Acc/acc_ulb": 16.0, "Loss/mean_noise_std": 0.5,
            "LR/lr_ac": 1e-3, "LR/lr_disc": 2e-3, "LR/lr_q": 3e-3, "Perf/total_fps": 8.0, "Perf/collection time": 0.5,
            "Perf/learning_time": 1.5, "Train/mean_reward": 1.5, "Train/mean_reward_i": 1.5, "Train/mean_reward_us": 1.5,
            "Train/mean_reward_ss": 1.5, "Train/mean_reward_t": 1.5, "Train/mean_episode_length": 1.0}
    assert set(log.scalars) == set(want)
    for k, v in want.items():
        assert log.scalars[k] == [(3, v)] or abs(log.last(k) - v) < 1e-6, k
    assert m["total"] == 1.5
    # a PPO-only update (no expert set) logs no discriminator scalars
    log2 = ScalarLog()
    log_bbc(runner, log2, 4, stats[:6], None, 0.5, 1.5)
    assert "Loss/ss_loss" not in log2.scalars and "Loss/hist_latent_loss" not in log2.scalars


def test_episode_rows_start_with_the_first_reset():
    """`infos['episode']` does not exist before the first reset (legged_robot.py:188-189), so the reference's `ep_infos` has no
    row for earlier steps (on_policy_runner.py:185-186): the staged per-step means are filtered by a device-side latch."""
    import torch
    from qa_b200.rsl_rl.train_log import EpisodeBook
    book = EpisodeBook(4, 5, ("total",), "cpu", num_episode_keys=3)
    means = torch.zeros(3)
    resets = [0, 0, 2, 0, 1]
    for t, k in enumerate(resets):
        if k:
            means = torch.full((3,), float(t))
        book.record(torch.zeros(4, dtype=torch.uint8), torch.zeros(4, 1), means, num_resets=torch.tensor([k], dtype=torch.int32))
    book.flush()
    assert [float(r[0]) for r in book.episode_rows] == [2.0, 2.0, 4.0]
    # the latch survives the flush: the next iteration's rows are all valid
    book.record(torch.zeros(4, dtype=torch.uint8), torch.zeros(4, 1), means, num_resets=torch.tensor([0], dtype=torch.int32))
    book.flush()
    assert len(book.episode_rows) == 1


def test_normalizer_load_moments_keeps_device_tensors():
    import numpy as np
    import torch
    from qa_b200.rsl_rl.utils import Normalizer
    n = Normalizer(6)
    mean32, std32 = n.device_moments("cpu")
    n.update_torch(torch.randn(50, 6))
    st = n._device_state("cpu")
    n.load_moments(np.arange(6.0), np.full(6, 4.0), 123.0)
    assert n._device_state("cpu")[1] is st[1] and n.device_moments("cpu")[0] is st[4]
    assert torch.allclose(st[4], torch.arange(6.0)) and torch.allclose(st[5], torch.sqrt(torch.full((6,), 4.0 + n.epsilon)))
    assert float(st[3]) == 123.0 and n.count == 123.0
    x = torch.randn(3, 6)
    assert torch.allclose(n.normalize_torch(x, "cpu"), torch.clamp((x - torch.arange(6.0)) / st[5], -n.clip_obs, n.clip_obs))
