"""`IsaacGymPhysics`: the `PhysicsBackend` the envs need, over IsaacGym's tensor API -- the calls the reference makes, in its
order (bbc/legged_gym/envs/base/legged_robot.py:100-106, 129-131, 594-596, 632-634, 687, 747-770).

IsaacGym (Preview 4) is a closed python <= 3.8 binary and cannot be imported in the build image, so `gym`, `sim` and the
`gymtorch` module are handed in by the caller (`from isaacgym import gymtorch`); `tests/test_isaacgym_backend.py` drives the
class with recording stand-ins and checks the call sequence against the reference's.  IsaacGym stays the physics backend; this
class is the whole contact surface with it.
"""
import torch

from .legged_robot import PhysicsBackend


class IsaacGymPhysics(PhysicsBackend):
    def __init__(self, gym, sim, num_envs: int, gymtorch=None, num_robot_dofs: int = 12):
        if gymtorch is None:
            from isaacgym import gymtorch                                    # noqa: F811  (only where IsaacGym exists)
        self.gym, self.sim, self.gymtorch, self.num_envs = gym, sim, gymtorch, num_envs
        wrap = gymtorch.wrap_tensor
        # :747-770 -- views of simulator memory; the env kernels write resets into them in place
        self.root_states = wrap(gym.acquire_actor_root_state_tensor(sim))
        self.dof_state = wrap(gym.acquire_dof_state_tensor(sim))
        self.rigid_body_state = wrap(gym.acquire_rigid_body_state_tensor(sim))
        self.contact_forces = wrap(gym.acquire_net_contact_force_tensor(sim)).view(num_envs, -1, 3)
        self.num_robot_dofs = num_robot_dofs

    def set_dof_actuation_force(self, torques: torch.Tensor) -> None:       # :103
        self.gym.set_dof_actuation_force_tensor(self.sim, self.gymtorch.unwrap_tensor(torques))

    def simulate(self) -> None:                                              # :104-106
        self.gym.simulate(self.sim)
        self.gym.fetch_results(self.sim, True)
        self.gym.refresh_dof_state_tensor(self.sim)

    def refresh(self) -> None:                                               # :129-131
        self.gym.refresh_actor_root_state_tensor(self.sim)
        self.gym.refresh_net_contact_force_tensor(self.sim)
        self.gym.refresh_rigid_body_state_tensor(self.sim)

    def set_states_indexed(self, env_ids_i32: torch.Tensor, count) -> None:  # :594-596, :632-634
        n = int(count.item()) if torch.is_tensor(count) else int(count)      # the one host sync IsaacGym's indexed setters force
        if n == 0:
            return
        ids = self.gymtorch.unwrap_tensor(env_ids_i32)
        self.gym.set_dof_state_tensor_indexed(self.sim, self.gymtorch.unwrap_tensor(self.dof_state), ids, n)
        self.gym.set_actor_root_state_tensor_indexed(self.sim, self.gymtorch.unwrap_tensor(self.root_states), ids, n)

    def set_root_states_all(self) -> None:                                   # :687 (push_robots)
        self.gym.set_actor_root_state_tensor(self.sim, self.gymtorch.unwrap_tensor(self.root_states))
