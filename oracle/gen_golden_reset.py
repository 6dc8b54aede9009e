"""Golden vectors for `LeggedRobot.reset()` and a 3-step `reset(); step(); step()` chain FROM THE UNMODIFIED REFERENCE
(bbc/legged_gym/envs/base/legged_robot.py:67-76 -> reset_idx(all envs) + step(zero actions)).

Build container only (needs /root/reference):   python oracle/gen_golden_reset.py

The reference env is driven by oracle/ref_env.py (LeggedRobot.__new__ + injected synthetic state, IsaacGym calls land in the
stub's attribute sink, i.e. "physics" leaves the simulator tensors as reset_idx wrote them); the random sources return the
dense per-env draws the CUDA kernels consume in parity mode.  Writes tests/golden/bbc_env_reset_n64.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

from qa_b200 import config as C  # noqa: E402
from qa_b200.config import BbcEnvConfig  # noqa: E402
from qa_b200.mocap import MocapTable  # noqa: E402
from qa_b200 import synthetic  # noqa: E402
import ref_env  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(1)

STATE = ("commands", "latent_eps", "latent_c", "root_states", "dof_state", "obs_buf", "privileged_obs_buf", "obs_disc_buf",
         "obs_history_buf", "episode_length_buf", "last_actions", "last_dof_vel", "last_root_vel", "last_torques_org",
         "action_history_buf", "feet_air_time", "base_lin_vel", "base_ang_vel", "projected_gravity", "feet_forces", "contact_filt",
         "last_contacts", "rew_buf", "reset_buf", "time_out_buf", "torques_org", "contact_buf", "contact_force_buf")


def capture(env):
    out = {k: getattr(env, k).clone() for k in STATE}
    out["episode_sums"] = torch.stack([env.episode_sums[k] for k in C.REWARD_NAMES]).clone()
    out["extras_time_outs"] = env.extras["time_outs"].clone()
    out["extras_episode"] = torch.stack([env.extras["episode"]["rew_" + k] for k in C.REWARD_NAMES]).clone()
    return out


def main(N=64, seed=13, counter_before=37):
    files = ref_env.labelled_clip_files()
    table = MocapTable.from_json_files(files)
    cfg = BbcEnvConfig(num_envs=N)
    static = synthetic.make_static(cfg, seed=seed, terrain_cells=1600)
    snap = synthetic.make_snapshot(cfg, seed=seed, step=0, reset_frac=0.2, plant_frac=0.05)
    draws = [synthetic.make_rng_draws(cfg, seed=seed, step=t) for t in range(3)]
    for d in draws:
        d["mocap_clip_idx"] = table.sample_clip(d["rt_c_idx"], d["mocap_clip_u"])
    import glob
    from ref_harness import REFERENCE_ROOT
    ulb = sorted(glob.glob(os.path.join(REFERENCE_ROOT, "bbc", "mocap_data", "mocap_all_ulb", "*.json")))[:1]
    ref, env = ref_env.build_reference_env(cfg, static, snap, None)
    ML = ref.motion_loader

    class _Loader(ML.MotionLoader):            # as in gen_golden.py: one unlabelled clip, which mocap_state_init=True never reads
        def get_full_frame_batch(self, num_frames, latent_c_idx=None):
            ref_env.CTX.choice_calls = 0
            ref_env.CTX.latent_c_idx = latent_c_idx.cpu()
            return super().get_full_frame_batch(num_frames, latent_c_idx)

    env.motion_loader = _Loader(motion_files_lb=files, motion_files_ulb=ulb, mocap_category=env.mocap_category,
                                time_between_frames=env.dt, mocap_state_init=True, device="cpu")
    env.common_step_counter = counter_before
    env.global_counter = 5
    env.delay = torch.tensor(0.0)
    env.cfg.domain_rand.action_curr_step = []           # keep the delay fixed (the schedule pops at global step 0 otherwise)
    out = {}
    # reset(): the reset_idx(all) and the zero-action step consume the SAME draw set (one dict per public call here)
    ref_env.CTX.draws = draws[0]
    obs, priv = env.reset()
    assert obs is env.obs_buf and priv is env.privileged_obs_buf
    for k, v in capture(env).items():
        out[f"reset.{k}"] = v
    g = torch.Generator().manual_seed(seed)
    for t in (1, 2):
        ref_env.CTX.draws = draws[t]
        a = torch.randn(N, 12, generator=g)
        out[f"step{t}.actions_in"] = a
        ret = env.step(a.clone())
        for k, v in capture(env).items():
            out[f"step{t}.{k}"] = v
        out[f"step{t}.reset_env_ids"] = ret[5].clone()
        out[f"step{t}.terminal_disc_states"] = ret[6].clone()
    n0, n1, n2 = (int(out[f"{s}.reset_buf"].sum()) for s in ("reset", "step1", "step2"))
    print(f"  reset chain N={N}: resets in the reset() step {n0}, step1 {n1}, step2 {n2}; counter {env.common_step_counter}")
    assert env.common_step_counter == counter_before + 3
    save = {}
    for k, v in static.items():
        if k != "height_samples" and isinstance(v, torch.Tensor):
            save["static." + k] = v.numpy()
    for k, v in snap.items():
        if isinstance(v, torch.Tensor):
            save["snap." + k] = v.numpy()
    for t, d in enumerate(draws):
        for k, v in d.items():
            if isinstance(v, torch.Tensor):
                save[f"draws{t}.{k}"] = v.numpy()
    for k, v in out.items():
        save["ref." + k] = v.numpy()
    save["meta.seed"], save["meta.num_envs"], save["meta.counter_before"] = np.array(seed), np.array(N), np.array(counter_before)
    save["meta.global_counter"] = np.array(5)
    np.savez_compressed(os.path.join(GOLD, "bbc_env_reset_n64.npz"), **save)
    print("  wrote tests/golden/bbc_env_reset_n64.npz")


if __name__ == "__main__":
    main()
