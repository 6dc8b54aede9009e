"""TSC student vision modules behind the reference's names (tsc/rsl_rl/modules/depth_backbone.py:7-109, modules/byol.py):

  `DepthOnlyFCBackbone58x87`  (58,87) depth image -> 32-d scan-dot latent: Conv2d(1,32,5) -> MaxPool2d(2) -> ELU ->
                              Conv2d(32,64,3) -> ELU -> Flatten -> Linear(62400,128) -> ELU -> Linear(128,out) -> ELU/Tanh
  `RecurrentDepthBackbone`    backbone latent + proprioception -> Linear(97,128)/ELU/Linear(128,32) -> GRU(32,512) ->
                              Linear(512, 32 + n_delta_yaw + n_obst_type); the obstacle-type lanes are soft-maxed
  `DepthBYOL`                 the self-supervised learner the reference attaches to the backbone (`byol_learner`)

SURVEY 8(f)-3 keeps conv / GRU / batch-norm on cuDNN: these are plain `torch.nn` modules, written so that their
`state_dict()` keys and `parameters()` order equal the reference's (`base_backbone.image_compression.0.weight`,
`byol_learner.online_encoder.projector.1.running_mean`, `rnn.weight_ih_l0`, ...) -- a reference student checkpoint
(`depth_encoder_state_dict`, on_policy_runner.py:617-619) loads key for key.  The per-step depth PREPROCESSING is the
CUDA part of this path (K14, `qa_b200/depth.py`).

Random augmentation (byol.py:196-201) draws from python's `random` (apply / skip) and torch's global generator (noise)
in the reference's order, so that under equal seeds both implementations see the same augmented images -- that is how
`oracle/gen_golden_student.py` pins this file against the reference.
"""
import copy
import random

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


class DepthOnlyFCBackbone58x87(nn.Module):
    def __init__(self, prop_dim, scandots_output_dim, hidden_state_dim, output_activation=None, num_frames=1):
        super().__init__()
        self.num_frames = num_frames
        act = nn.ELU()
        self.image_compression = nn.Sequential(
            nn.Conv2d(num_frames, 32, kernel_size=5),                 # (1,58,87) -> (32,54,83)
            nn.MaxPool2d(kernel_size=2, stride=2),                    # -> (32,27,41)
            act,
            nn.Conv2d(32, 64, kernel_size=3),                         # -> (64,25,39)
            act,
            nn.Flatten(),
            nn.Linear(64 * 25 * 39, 128),
            act,
            nn.Linear(128, scandots_output_dim))
        self.output_activation = nn.Tanh() if output_activation == "tanh" else act
        self.augment = None                                            # set to the BYOL augmentation by the runner (:98)

    def forward(self, images: torch.Tensor):
        if self.augment:
            images = self.augment(images.clone())
        return self.output_activation(self.image_compression(images.unsqueeze(1)))


# ---- BYOL (byol.py; a port of the public byol-pytorch recipe to single-channel depth images) ---------------------------------
class _MaybeApply(nn.Module):
    """fn with probability p, decided by python's `random` (byol.py:57-65)."""

    def __init__(self, fn, p):
        super().__init__()
        self.fn, self.p = fn, p

    def forward(self, x):
        return x if random.random() > self.p else self.fn(x)


def add_background_noise(x):
    """Overwrites one random rectangle (< 1/4 of each side) of every image of the batch with either uniform noise in
    [-0.5, 0.5) or one constant from that range (byol.py:230-248).  In place, like the reference."""
    h, w = x.shape[1], x.shape[2]
    rh = torch.randint(1, h // 4, (1,)).item()
    rw = torch.randint(1, w // 4, (1,)).item()
    top = torch.randint(0, h - rh, (1,)).item()
    left = torch.randint(0, w - rw, (1,)).item()
    if torch.rand(1) < 0.5:
        patch = torch.rand((rh, rw)) - 0.5
    else:
        patch = torch.zeros((rh, rw)) + (torch.rand(1).item() - 0.5)
    x[:, top:top + rh, left:left + rw] = patch
    return x


def _default_augmentation():
    from torchvision import transforms as T
    return nn.Sequential(
        _MaybeApply(add_background_noise, p=0.1),
        _MaybeApply(lambda x: x + torch.randn_like(x) * 0.02, p=0.1),
        _MaybeApply(lambda x: x * (torch.rand_like(x) > 0.05).float(), p=0.05),
        _MaybeApply(T.GaussianBlur((3, 3), (0.5, 1.5)), p=0.1))


def _batchnorm():
    return nn.SyncBatchNorm if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else nn.BatchNorm1d


def _projection_mlp(dim, out, hidden):
    return nn.Sequential(nn.Linear(dim, hidden), _batchnorm()(hidden), nn.ReLU(inplace=True), nn.Linear(hidden, out))


class _Encoder(nn.Module):
    """`net` followed by a projection head; keys `net.*`, `projector.*` (byol.py:110-176 with hidden_layer = -1: the
    representation is the network's output)."""

    def __init__(self, net, projection_size, hidden_size):
        super().__init__()
        self.net = net
        self.projector = None
        self.projection_size, self.hidden_size = projection_size, hidden_size

    def forward(self, x, return_projection=True):
        rep = self.net(x)
        if not return_projection:
            return rep
        if self.projector is None:                                     # sized by the first batch, like the reference
            self.projector = _projection_mlp(rep.shape[1], self.projection_size, self.hidden_size).to(rep)
        return self.projector(rep), rep


def _byol_loss(x, y):
    return 2 - 2 * (F.normalize(x, dim=-1, p=2) * F.normalize(y, dim=-1, p=2)).sum(dim=-1)


class DepthBYOL(nn.Module):
    def __init__(self, net, image_size, projection_size=64, projection_hidden_size=1024, moving_average_decay=0.99):
        super().__init__()
        self.net = net
        self.augment1 = _default_augmentation()
        self.augment2 = self.augment1
        self.online_encoder = _Encoder(net, projection_size, projection_hidden_size)
        self.target_encoder = None
        self.beta = moving_average_decay
        self.online_predictor = _projection_mlp(projection_size, projection_size, projection_hidden_size)
        dev = next(net.parameters()).device
        self.to(dev)
        self.forward(torch.randn(2, image_size[0], image_size[1], device=dev))   # instantiates projector + target (:229)

    def _target(self):
        if self.target_encoder is None:
            t = copy.deepcopy(self.online_encoder)
            for p in t.parameters():
                p.requires_grad = False
            self.target_encoder = t
        return self.target_encoder

    def reset_moving_average(self):
        self.target_encoder = None

    @torch.no_grad()
    def update_moving_average(self):
        assert self.target_encoder is not None, "target encoder has not been created yet"
        for cur, ma in zip(self.online_encoder.parameters(), self.target_encoder.parameters()):
            ma.data = ma.data * self.beta + (1 - self.beta) * cur.data

    def forward(self, x, return_embedding=False, return_projection=True):
        assert not (self.training and x.shape[0] == 1), "batch norm in the projection head needs more than one sample"
        if return_embedding:
            return self.online_encoder(x, return_projection=return_projection)
        images = torch.cat((self.augment1(x.clone()), self.augment2(x.clone())), dim=0)
        pred_one, pred_two = self.online_predictor(self.online_encoder(images)[0]).chunk(2, dim=0)
        with torch.no_grad():
            proj_one, proj_two = self._target()(images)[0].detach().chunk(2, dim=0)
        return (_byol_loss(pred_one, proj_two) + _byol_loss(pred_two, proj_one)).mean()


BYOL = DepthBYOL


class RecurrentDepthBackbone(nn.Module):
    def __init__(self, base_backbone, n_depth_latent, env_cfg) -> None:
        super().__init__()
        e = env_cfg.env if hasattr(env_cfg, "env") else env_cfg
        self.n_delta_yaw, self.n_obst_type, self.n_depth_latent = e.n_delta_yaw, e.n_obst_type, n_depth_latent
        self.tanh = nn.Tanh()
        self.softmax = nn.Softmax(dim=-1)
        self.base_backbone = base_backbone
        self.byol_learner = DepthBYOL(base_backbone, image_size=(58, 87))
        self.combination_mlp = nn.Sequential(nn.Linear(n_depth_latent + e.n_proprio, 128), nn.ELU(),
                                             nn.Linear(128, n_depth_latent))
        self.rnn = nn.GRU(input_size=n_depth_latent, hidden_size=512, batch_first=True)
        self.output_mlp = nn.Sequential(nn.Linear(512, n_depth_latent + self.n_delta_yaw + self.n_obst_type))
        self.hidden_states = None

    def forward(self, depth_image, proprioception):
        """-> (N, latent + n_delta_yaw + n_obst_type); the GRU state is carried across calls WITH its graph until
        `detach_hidden_states()` (the distillation loss back-propagates through the whole rollout, :331-403)."""
        x = self.base_backbone(depth_image)
        x = self.combination_mlp(torch.cat((x, proprioception), dim=-1))
        x, self.hidden_states = self.rnn(x[:, None, :], self.hidden_states)
        x = self.output_mlp(x.squeeze(1))
        k = self.n_depth_latent + self.n_delta_yaw
        return torch.cat([x[:, :k], self.softmax(x[:, k:])], dim=-1)

    def detach_hidden_states(self):
        self.hidden_states = self.hidden_states.detach().clone()
