"""CPU: the oracle restatement reproduces the golden vectors that the UNMODIFIED reference produced
(oracle/gen_golden.py), i.e. the pin of the oracle travels with the repo."""
import numpy as np
import pytest
import torch

import bbc_env as O
import trainer as OT
from helpers import GOLD, load_env_golden, mocap_table, assert_close
from qa_b200 import config as C


@pytest.mark.parametrize("name", ["n64a", "n64b_push"])
def test_env_oracle_matches_reference_golden(name):
    torch.set_num_threads(1)
    cfg, static, snap, draws, ref, meta = load_env_golden(name)
    table = mocap_table()
    o = O.post_physics_step(cfg, static, snap, draws, table, int(meta["counter_before"]) + 1)
    exact = ["reset_buf", "time_out_buf", "reset_env_ids", "episode_length_buf", "contact_filt", "last_contacts",
             "latent_c"]
    for k in exact:
        assert torch.equal(o[k].to(torch.int64), ref[k].to(torch.int64)), k
    for k in ["rew_buf", "obs_buf", "privileged_obs_buf", "obs_disc_buf", "obs_history_buf", "commands", "latent_eps",
              "root_states", "dof_state", "terminal_disc_states", "last_actions", "last_dof_vel", "last_root_vel",
              "last_torques_org", "action_history_buf", "feet_air_time", "base_lin_vel", "base_ang_vel",
              "projected_gravity", "roll", "pitch", "yaw", "feet_forces", "episode_sums", "measured_heights"]:
        assert torch.equal(o[k], ref[k]), f"{name}: oracle != golden on {k} (bit-exact on CPU expected)"
    if "episode_rew_means" in ref:
        assert torch.equal(o["episode_rew_means"], ref["episode_rew_means"])
    tq, tq_org = O.compute_torques(cfg, {**static, **snap}, snap["actions"].clone())
    assert torch.equal(tq, ref["torques"]) and torch.equal(tq_org, ref["torques_org"])
    hist, act = O.action_push(cfg, snap["action_history_buf"], snap["actions"], delay=1)
    assert torch.equal(hist, ref["act_hist_pushed"]) and torch.equal(act, ref["actions_clipped"])


def test_mocap_table_fixture_structure():
    t = mocap_table()
    assert t.frames.shape == (1196, 49) and t.num_clips == 17
    assert t.clip_label.tolist() == [3, 3, 3, 3, 4, 4, 4, 4, 1, 1, 1, 2, 2, 2, 0, 0, 0]
    assert int(t.mode_offset[-1]) == 17
    q = t.frames[:, 3:7]
    assert torch.allclose(q.norm(dim=-1), torch.ones(1196), atol=1e-5) and bool((q[:, 3] >= 0).all())


@pytest.mark.parametrize("name", ["gae_t24_n64", "gae_t24_n100", "gae_t5_n33"])
def test_gae_oracle_matches_reference_golden(name):
    torch.set_num_threads(1)
    z = np.load(f"{GOLD}/trainer_{name}.npz")
    t = lambda k: torch.from_numpy(z[k])          # noqa: E731
    ret, adv = OT.compute_returns(t("rewards"), t("values"), t("dones"), t("last_values"), float(z["gamma"]),
                                  float(z["lam"]))
    assert torch.equal(ret, t("ref_returns")) and torch.equal(adv, t("ref_advantages"))


def test_reward_order_is_alphabetical():
    assert list(C.REWARD_NAMES) == sorted(C.REWARD_NAMES) and len(C.REWARD_NAMES) == 14


def test_policy_oracle_matches_reference_golden():
    """oracle/trainer.py act / predict_disc_reward / ppo_losses vs the reference's recorded outputs."""
    from qa_b200 import synthetic
    torch.set_num_threads(1)
    z = np.load(f"{GOLD}/trainer_policy_seed3.npz")
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    w = synthetic.make_weights(int(g["in.weights_seed"]))
    for he in (False, True):
        o = OT.act(w["ac"], w["est"], g["in.obs"], g["in.obs"], g["in.draw"], hist_encoding=he)
        for k in ("actions", "values", "actions_log_prob", "action_mean", "action_sigma"):
            assert torch.allclose(o[k], g[f"act{int(he)}.{k}"], rtol=1e-6, atol=1e-6), k
    r = OT.predict_disc_reward(w["disc"], g["in.rew_t"], g["in.obs"], g["in.disc_hist"], w["norm_mean"], w["norm_var"],
                               0.02, float(g["in.task_obs_weight"]))
    for k, v in zip(("rewards", "reward_i", "reward_us", "reward_ss", "reward_t"), r):
        assert v.dtype == g[f"disc.{k}"].dtype and torch.allclose(v, g[f"disc.{k}"], rtol=1e-6, atol=1e-7), k
    batch = dict(obs=g["in.obs"], critic_obs=g["in.obs"], actions=g["in.batch.actions"],
                 target_values=g["in.batch.target_values"], advantages=g["in.batch.advantages"],
                 returns=g["in.batch.returns"], old_actions_log_prob=g["in.batch.old_actions_log_prob"],
                 old_mu=g["in.batch.old_mu"], old_sigma=g["in.batch.old_sigma"])
    L = OT.ppo_losses(w["ac"], w["est"], batch, priv_reg_coef=OT.priv_reg_coef(int(g["in.priv_reg_counter"])))
    for k in ("surrogate_loss", "value_loss", "b_loss", "entropy", "priv_reg_loss", "estimator_loss"):
        assert torch.allclose(L[k], g[f"ppo.{k}"], rtol=1e-5, atol=1e-7), k
    assert abs(OT.adaptive_lr(1e-3, float(L["kl_mean"])) - float(g["ppo.lr_new"])) < 1e-12
