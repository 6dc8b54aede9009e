"""`IsaacGymPhysics` against recording stand-ins for `gym` / `gymtorch`: the adapter must make the reference's calls in the
reference's order (bbc/legged_gym/envs/base/legged_robot.py:100-106, 129-131, 594-596, 632-634, 687, 747-770)."""
import torch

from qa_b200.isaacgym_backend import IsaacGymPhysics


class _Gym:
    def __init__(self, n, b):
        self.calls = []
        self.mem = dict(root=torch.zeros(n, 13), dof=torch.zeros(n * 12, 2), rb=torch.zeros(n * b, 13), cf=torch.zeros(n * b, 3))

    def __getattr__(self, name):
        def call(sim, *args):
            self.calls.append((name,) + tuple(a if isinstance(a, (int, bool, str)) else id(a) for a in args))
            return {"acquire_actor_root_state_tensor": "root", "acquire_dof_state_tensor": "dof",
                    "acquire_rigid_body_state_tensor": "rb", "acquire_net_contact_force_tensor": "cf"}.get(name)
        return call


class _GymTorch:
    def __init__(self, gym):
        self.gym = gym

    def wrap_tensor(self, handle):
        return self.gym.mem[handle]

    @staticmethod
    def unwrap_tensor(t):
        return t


def test_adapter_makes_the_reference_calls_in_order():
    n, b = 4, 19
    gym = _Gym(n, b)
    ph = IsaacGymPhysics(gym, "sim", n, gymtorch=_GymTorch(gym))
    assert [c[0] for c in gym.calls] == ["acquire_actor_root_state_tensor", "acquire_dof_state_tensor",
                                         "acquire_rigid_body_state_tensor", "acquire_net_contact_force_tensor"]
    assert ph.root_states is gym.mem["root"] and ph.dof_state is gym.mem["dof"] and ph.rigid_body_state is gym.mem["rb"]
    assert tuple(ph.contact_forces.shape) == (n, b, 3) and ph.contact_forces.data_ptr() == gym.mem["cf"].data_ptr()   # a view
    gym.calls.clear()
    tq = torch.zeros(n, 12)
    ph.set_dof_actuation_force(tq)
    ph.simulate()
    assert gym.calls == [("set_dof_actuation_force_tensor", id(tq)), ("simulate",), ("fetch_results", True), ("refresh_dof_state_tensor",)]
    gym.calls.clear()
    ph.refresh()
    assert [c[0] for c in gym.calls] == ["refresh_actor_root_state_tensor", "refresh_net_contact_force_tensor",
                                         "refresh_rigid_body_state_tensor"]
    gym.calls.clear()
    ids = torch.tensor([1, 3, 0, 0], dtype=torch.int32)
    ph.set_states_indexed(ids, torch.tensor([0]))
    assert gym.calls == []                                           # no reset: no simulator call (the reference guards on len(env_ids))
    ph.set_states_indexed(ids, torch.tensor([2]))
    assert gym.calls == [("set_dof_state_tensor_indexed", id(ph.dof_state), id(ids), 2),
                         ("set_actor_root_state_tensor_indexed", id(ph.root_states), id(ids), 2)]
    gym.calls.clear()
    ph.set_root_states_all()
    assert gym.calls == [("set_actor_root_state_tensor", id(ph.root_states))]


def test_tsc_adapter_slices_obstacle_actors_and_resets_them_with_the_robots():
    """tsc/legged_gym/envs/base/legged_robot.py:976-998 (slices), :822-838 / :886-899 (indexed setters), :131-133 (torques)."""
    from qa_b200.isaacgym_backend import IsaacGymPhysicsTSC
    n, b, per_env = 4, 17, 3
    lay = dict(num_obst=n * per_env, num_border=n, num_obst_links=n * per_env * 2, num_obst_joints=n)
    gym = _Gym(n, b)
    gym.mem = dict(root=torch.arange((n + lay["num_obst"] + n) * 13.).view(-1, 13), dof=torch.zeros(n * 12 + n, 2),
                   rb=torch.zeros(n * b + lay["num_obst_links"] + n, 13), cf=torch.zeros(n * b + lay["num_obst_links"] + n, 3))
    seesaw = torch.tensor([n + 1, n + 4, n + 7, n + 10], dtype=torch.int32)          # actor index of each env's seesaw
    gains = (torch.full((n,), 2.0), torch.full((n,), 0.5), torch.full((n,), 0.3))
    ph = IsaacGymPhysicsTSC(gym, "sim", n, lay, seesaw, gains, gymtorch=_GymTorch(gym))
    assert ph.root_states.shape == (n, 13) and ph.obst_root_states.shape == (n * per_env, 13) and ph.border_root_states.shape == (n, 13)
    assert ph.dof_state.shape == (n * 12, 2) and ph.obst_dof_state.shape == (n, 2)
    assert ph.rigid_body_state.shape == (n * b, 13) and ph.contact_forces.shape == (n, b, 3)
    assert ph.root_states.data_ptr() == gym.mem["root"].data_ptr() and ph.obst_dof_state.data_ptr() == gym.mem["dof"][n * 12:].data_ptr()
    seen = {}
    gym_call = gym.__getattr__

    def spy(name):
        inner = gym_call(name)

        def call(sim, *args):
            seen[name] = args
            return inner(sim, *args)
        return call
    gym.set_dof_actuation_force_tensor = spy("set_dof_actuation_force_tensor")
    gym.set_dof_state_tensor_indexed = spy("set_dof_state_tensor_indexed")
    gym.set_actor_root_state_tensor_indexed = spy("set_actor_root_state_tensor_indexed")
    ph.obst_dof_state[:, 0], ph.obst_dof_state[:, 1] = 0.1, 0.2
    ph.set_dof_actuation_force(torch.ones(n, 12))
    (tq,) = seen["set_dof_actuation_force_tensor"]
    assert tq.shape == (n * 12 + n,) and torch.allclose(tq[-n:], torch.full((n,), 2.0 * (0.3 - 0.1) - 0.5 * 0.2))
    gym.calls.clear()
    ph.refresh()
    assert [c[0] for c in gym.calls][-1] == "refresh_force_sensor_tensor" and len(gym.calls) == 4
    ids = torch.tensor([2, 0, 0, 0], dtype=torch.int32)
    ph.set_states_indexed(ids, torch.tensor([2]))
    dof_t, dof_ids, dof_n = seen["set_dof_state_tensor_indexed"]
    assert dof_t is ph._dof_all and dof_n == 4 and dof_ids.tolist() == [2, 0, n + 7, n + 1] and dof_ids.dtype == torch.int32
    root_t, root_ids, root_n = seen["set_actor_root_state_tensor_indexed"]
    assert root_t is ph._root_all and root_n == 6 and root_ids.tolist() == [2, 0, 2 + n, 0 + n, 2 + n + n * per_env, 0 + n + n * per_env]
