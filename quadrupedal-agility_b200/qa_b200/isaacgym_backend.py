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


class IsaacGymPhysicsTSC(IsaacGymPhysics):
    """The TSC course adds obstacle and border actors BEHIND the robots in every simulator tensor
    (tsc/legged_gym/envs/base/legged_robot.py:976-998): the env sees the robot slices, the indexed setters of a reset also touch
    the env's obstacle / border actors (:822-838, :886-899), the torque tensor carries the obstacle joints' PD torques behind
    the robots' (:131-133, :792-794), and `refresh` also refreshes the force sensors (:233).

    layout: dict(num_obst, num_border, num_obst_links, num_obst_joints) -- actor / link / joint counts of the course;
    seesaw_actor_index: (N,) int32, actor index of env e's seesaw (`nonzero(obst_idx == 3)[e] + num_envs`, :827-831);
    obst_gains: (stiffness, damping, target) tensors of the obstacle joints ((J,), (J,), (J,)).
    """

    def __init__(self, gym, sim, num_envs, layout, seesaw_actor_index, obst_gains, gymtorch=None, border_height_range=None):
        if gymtorch is None:
            from isaacgym import gymtorch                                    # noqa: F811
        self.gym, self.sim, self.gymtorch, self.num_envs = gym, sim, gymtorch, num_envs
        wrap = gymtorch.wrap_tensor
        no, nb, nl, nj = layout["num_obst"], layout["num_border"], layout["num_obst_links"], layout["num_obst_joints"]
        self._root_all = wrap(gym.acquire_actor_root_state_tensor(sim))
        self._dof_all = wrap(gym.acquire_dof_state_tensor(sim))
        rb_all = wrap(gym.acquire_rigid_body_state_tensor(sim))
        cf_all = wrap(gym.acquire_net_contact_force_tensor(sim))
        self.root_states = self._root_all[:-(no + nb)]
        self.obst_root_states, self.border_root_states = self._root_all[-(no + nb):-nb], self._root_all[-nb:]
        self.dof_state, self.obst_dof_state = self._dof_all[:-nj], self._dof_all[-nj:]
        self.rigid_body_state = rb_all[:-(nl + nb)]
        self.contact_forces = cf_all[:-(nl + nb)].view(num_envs, -1, 3)
        self.num_obst = no
        self.seesaw_actor_index = seesaw_actor_index.to(torch.int32)
        self.obst_stiffness, self.obst_damping, self.target_obst_dof_pos = obst_gains
        self.border_height_range = border_height_range

    def set_dof_actuation_force(self, torques: torch.Tensor) -> None:       # :131-133 with the obstacle PD of :792-794
        obst = self.obst_stiffness * (self.target_obst_dof_pos - self.obst_dof_state[:, 0]) - self.obst_damping * self.obst_dof_state[:, 1]
        self.gym.set_dof_actuation_force_tensor(self.sim, self.gymtorch.unwrap_tensor(torch.cat([torques.flatten(), obst.flatten()])))

    def refresh(self) -> None:                                               # :230-233
        super().refresh()
        self.gym.refresh_force_sensor_tensor(self.sim)

    def set_states_indexed(self, env_ids_i32: torch.Tensor, count) -> None:  # :822-838, :886-899
        n = int(count.item()) if torch.is_tensor(count) else int(count)
        if n == 0:
            return
        ids = env_ids_i32[:n]
        un = self.gymtorch.unwrap_tensor
        dof_ids = torch.cat([ids, self.seesaw_actor_index[ids.long()]]).to(torch.int32)
        self.gym.set_dof_state_tensor_indexed(self.sim, un(self._dof_all), un(dof_ids), 2 * n)
        if self.border_height_range is not None:                             # :888-890 (camera runs randomise the border height)
            lo, hi = self.border_height_range
            self.border_root_states[:, 2] = -((hi - lo) * torch.rand(self.border_root_states.shape[0], device=ids.device) + lo)
        N = self.root_states.shape[0]
        root_ids = torch.cat([ids, ids + N, ids + N + self.num_obst]).to(torch.int32)
        self.gym.set_actor_root_state_tensor_indexed(self.sim, un(self._root_all), un(root_ids), 3 * n)
