"""Generate the golden vectors in tests/golden/ FROM THE REFERENCE ITSELF, and pin the oracle.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz, asserts oracle == reference

For every case the UNMODIFIED reference classes (imported by oracle/ref_harness.py, driven by
oracle/ref_env.py) run on a seeded synthetic state; the oracle restatement (oracle/bbc_env.py,
oracle/trainer.py) runs on the same inputs and must agree (bit-exact on CPU where the op
sequence is identical, else within the stated tolerance); inputs + reference outputs are then
saved as small fixtures so that the GPU box -- which has no /root/reference -- can check the
CUDA path against what the reference really produced.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "quadrupedal-agility_b200"))

import glob  # noqa: E402

from qa_b200 import config as C  # noqa: E402
from qa_b200.config import BbcEnvConfig  # noqa: E402
from qa_b200.mocap import MocapTable  # noqa: E402
from qa_b200 import synthetic  # noqa: E402
import bbc_env as O  # noqa: E402
import ref_env  # noqa: E402
from ref_harness import REFERENCE_ROOT  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(1)          # deterministic reductions while generating


def _np(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.detach().cpu().numpy()
    return out


def max_rel(a, b):
    a = a.double()
    b = b.double()
    return float(((a - b).abs() / (b.abs() + 1e-6)).max()) if a.numel() else 0.0


def gen_mocap_table():
    files = ref_env.labelled_clip_files()
    table = MocapTable.from_json_files(files)
    import json
    raw_w = [float(json.load(open(f))["MotionWeight"]) for f in files]
    table.save_npz(os.path.join(GOLD, "mocap_lb_table.npz"), raw_w)
    back = MocapTable.from_npz(os.path.join(GOLD, "mocap_lb_table.npz"))
    assert torch.equal(back.frames, table.frames) and torch.equal(back.clip_len_s, table.clip_len_s)
    assert torch.equal(back.mode_cdf, table.mode_cdf)
    return files, table


def check_table_against_reference(files, table, loader):
    """qa_b200.mocap.MocapTable vs the reference MotionLoader's own parsed clips."""
    for k in range(table.num_clips):
        s, n = int(table.clip_start[k]), int(table.clip_nframes[k])
        assert torch.equal(table.frames[s:s + n], loader.mocap_trajectory_full_lb[k]), files[k]
    assert np.array_equal(table.clip_len_s.numpy(), loader.mocap_lens_lb)
    assert np.array_equal(table.clip_nframes.numpy(), loader.mocap_num_frames_lb)
    assert np.array_equal(table.clip_label.numpy(), loader.mocap_label)
    assert np.allclose(table.clip_weight.numpy(), loader.mocap_weights_lb, rtol=0, atol=0)
    print("  mocap table == reference MotionLoader clips: OK "
          f"({table.num_clips} clips, {table.frames.shape[0]} frames)")


def run_env_case(name, N, seed, counter_before, files, table, save=True, reset_frac=0.015, plant_frac=0.003):
    cfg = BbcEnvConfig(num_envs=N)
    static = synthetic.make_static(cfg, seed=seed, terrain_cells=1600)
    snap = synthetic.make_snapshot(cfg, seed=seed, step=0, reset_frac=reset_frac, plant_frac=plant_frac)
    draws = synthetic.make_rng_draws(cfg, seed=seed, step=0)
    draws["mocap_clip_idx"] = table.sample_clip(draws["rt_c_idx"], draws["mocap_clip_u"])

    ulb = sorted(glob.glob(os.path.join(REFERENCE_ROOT, "bbc", "mocap_data", "mocap_all_ulb", "*.json")))[:1]
    ref, env = ref_env.build_reference_env(cfg, static, snap, None)
    env.motion_loader = None
    # real loader (labelled clips + one unlabelled clip, which mocap_state_init=True never reads)
    RefLoader = type(env).__mro__  # noqa: F841
    ML = ref.motion_loader

    class _Loader(ML.MotionLoader):
        def get_full_frame_batch(self, num_frames, latent_c_idx=None):
            ref_env.CTX.choice_calls = 0
            ref_env.CTX.latent_c_idx = latent_c_idx.cpu()
            return super().get_full_frame_batch(num_frames, latent_c_idx)

    env.motion_loader = _Loader(motion_files_lb=files, motion_files_ulb=ulb, mocap_category=env.mocap_category,
                                time_between_frames=env.dt, mocap_state_init=True, device="cpu")
    if name.endswith("a"):
        check_table_against_reference(files, table, env.motion_loader)
    ref_env.CTX.draws = draws

    # ---- a2: PD torques (reference) -------------------------------------------------------
    tq_ref = env._compute_torques(snap["actions"].clone())
    tq_org_ref = env.torques_org.clone()
    tq_o, tq_org_o = O.compute_torques(cfg, {**static, **snap}, snap["actions"].clone())
    assert torch.equal(tq_ref, tq_o) and torch.equal(tq_org_ref, tq_org_o), "torques oracle != reference"
    env.torques_org = snap["torques_org"].clone()          # the snapshot's own torques_org feeds the rewards

    # ---- a1: action push (reference lines executed verbatim through step()'s front half) --
    hist_o, act_o = O.action_push(cfg, snap["action_history_buf"], snap["actions"], delay=1)

    # ---- a3..a11: post_physics_step (reference) ---------------------------------------------
    env.common_step_counter = counter_before
    env_ids, terminal = env.post_physics_step()
    o = O.post_physics_step(cfg, static, snap, draws, table, counter_before + 1)

    def eq(label, a, b, exact=True):
        a = a.float() if a.dtype == torch.bool else a
        b = b.float() if b.dtype == torch.bool else b
        if exact:
            ok = torch.equal(a, b)
        else:
            ok = torch.allclose(a, b, rtol=1e-6, atol=1e-7)
        if not ok:
            d = (a.double() - b.double()).abs()
            raise AssertionError(f"{name}: oracle != reference on {label}: max abs {float(d.max()):.3e} "
                                 f"at {int(d.argmax())}")

    eq("reset_buf", env.reset_buf, o["reset_buf"])
    eq("time_out_buf", env.time_out_buf, o["time_out_buf"])
    eq("reset_env_ids", env_ids, o["reset_env_ids"])
    eq("terminal_disc_states", terminal, o["terminal_disc_states"])
    eq("rew_buf", env.rew_buf, o["rew_buf"])
    eq("measured_heights", env.measured_heights, o["measured_heights"])
    eq("commands", env.commands, o["commands"])
    eq("latent_eps", env.latent_eps, o["latent_eps"])
    eq("latent_c", env.latent_c, o["latent_c"])
    eq("root_states", env.root_states, o["root_states"])
    eq("dof_state", env.dof_state, o["dof_state"])
    eq("obs_buf", env.obs_buf, o["obs_buf"])
    eq("privileged_obs_buf", env.privileged_obs_buf, o["privileged_obs_buf"])
    eq("obs_disc_buf", env.obs_disc_buf, o["obs_disc_buf"])
    eq("obs_history_buf", env.obs_history_buf, o["obs_history_buf"])
    eq("episode_length_buf", env.episode_length_buf, o["episode_length_buf"])
    eq("last_actions", env.last_actions, o["last_actions"])
    eq("last_dof_vel", env.last_dof_vel, o["last_dof_vel"])
    eq("last_root_vel", env.last_root_vel, o["last_root_vel"])
    eq("last_torques_org", env.last_torques_org, o["last_torques_org"])
    eq("action_history_buf", env.action_history_buf, o["action_history_buf"])
    eq("feet_air_time", env.feet_air_time, o["feet_air_time"])
    eq("base_lin_vel", env.base_lin_vel, o["base_lin_vel"])
    eq("projected_gravity", env.projected_gravity, o["projected_gravity"])
    eq("contact_filt", env.contact_filt, o["contact_filt"])
    eq("last_contacts", env.last_contacts, o["last_contacts"])
    es_ref = torch.stack([env.episode_sums[k] for k in C.REWARD_NAMES])
    assert list(env.episode_sums.keys()) == list(C.REWARD_NAMES), list(env.episode_sums.keys())
    assert env.reward_names == list(C.REWARD_NAMES)
    eq("episode_sums", es_ref, o["episode_sums"])
    if len(env_ids):
        means_ref = torch.stack([env.extras["episode"]["rew_" + k] for k in C.REWARD_NAMES])
        eq("episode_rew_means", means_ref, o["episode_rew_means"])
        eq("extras.time_outs", env.extras["time_outs"], o["time_out_buf"])
    n_reset = int(env.reset_buf.sum())
    n_to = int(env.time_out_buf.sum())
    n_rs = int((env.episode_length_buf % 300 == 0).sum())
    print(f"  {name}: N={N} resets={n_reset} timeouts={n_to} push={o['do_push']} "
          f"resampled={n_rs} rew>0={int(((o['rew_buf']>0)).sum())}: oracle == reference (bit-exact on CPU)")

    if save:
        out = {}
        for k, v in _np(static).items():
            if k == "height_samples":
                continue                                   # regenerated from the seed (5 MB)
            out["static." + k] = v
        for k, v in _np(snap).items():
            out["snap." + k] = v
        for k, v in _np(draws).items():
            out["draws." + k] = v
        ref_out = dict(
            torques=tq_ref, torques_org=tq_org_ref, act_hist_pushed=hist_o, actions_clipped=act_o,
            reset_buf=env.reset_buf, time_out_buf=env.time_out_buf, reset_env_ids=env_ids,
            terminal_disc_states=terminal, rew_buf=env.rew_buf, measured_heights=env.measured_heights,
            commands=env.commands, latent_eps=env.latent_eps, latent_c=env.latent_c,
            root_states=env.root_states, dof_state=env.dof_state, obs_buf=env.obs_buf,
            privileged_obs_buf=env.privileged_obs_buf, obs_disc_buf=env.obs_disc_buf,
            obs_history_buf=env.obs_history_buf, episode_length_buf=env.episode_length_buf,
            last_actions=env.last_actions, last_dof_vel=env.last_dof_vel, last_root_vel=env.last_root_vel,
            last_torques_org=env.last_torques_org, action_history_buf=env.action_history_buf,
            feet_air_time=env.feet_air_time, base_lin_vel=env.base_lin_vel, base_ang_vel=env.base_ang_vel,
            projected_gravity=env.projected_gravity, roll=env.roll, pitch=env.pitch, yaw=env.yaw,
            feet_forces=env.feet_forces, contact_filt=env.contact_filt, last_contacts=env.last_contacts,
            episode_sums=es_ref)
        if len(env_ids):
            ref_out["episode_rew_means"] = means_ref
        for k, v in _np(ref_out).items():
            out["ref." + k] = v
        out["meta.seed"] = np.array(seed)
        out["meta.counter_before"] = np.array(counter_before)
        out["meta.num_envs"] = np.array(N)
        out["meta.reset_frac"] = np.array(reset_frac)
        out["meta.plant_frac"] = np.array(plant_frac)
        np.savez_compressed(os.path.join(GOLD, f"bbc_env_{name}.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    print("[gen_golden] mocap table")
    files, table = gen_mocap_table()
    print("[gen_golden] BBC env cases (reference LeggedRobot vs oracle)")
    run_env_case("n64a", 64, 7, counter_before=12, files=files, table=table, reset_frac=0.25, plant_frac=0.06)
    run_env_case("n64b_push", 64, 8, counter_before=399, files=files, table=table, reset_frac=0.25, plant_frac=0.06)
    run_env_case("n4096_check", 4096, 1234, counter_before=3, files=files, table=table, save=False)
    # a step on which nothing resets, nothing times out and nothing is resampled: reset_idx's empty-set early return and the
    # untouched extras (SURVEY a')
    run_env_case("zero_reset_check", 64, 11, counter_before=5, files=files, table=table, save=False, reset_frac=0.0, plant_frac=0.0)
    import gen_golden_trainer
    gen_golden_trainer.main()


if __name__ == "__main__":
    main()
