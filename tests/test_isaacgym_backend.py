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
