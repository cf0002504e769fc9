"""Host-side mirrors of the reference's network containers (url_benchmark/agent/fb_modules.py:43-230) and the
small helpers of url_benchmark/utils.py the agent API exposes.

These nn.Modules exist for the *interface*: `agent.actor`, `agent.forward_net`, `agent.backward_net` ... must be
callables with `.parameters()` in the reference's registration order (init_from / checkpoints / goals.py:560 /
demo).  Their parameters are views into the flat fp32 segments the CUDA step trains in place, so the two always
agree.  The gradient step itself never goes through these modules (see engine.py / csrc/).
"""
from __future__ import annotations

import math
import re
import typing as tp

import numpy as np
import torch
import torch.nn.functional as F
from torch import distributions as pyd
from torch import nn
from torch.distributions.utils import _standard_normal


def schedule(schdl: tp.Any, step: int) -> float:
    """utils.schedule (utils.py:235-255): constant, linear(a,b,n) or step_linear(a,b,n,c,m)."""
    try:
        return float(schdl)
    except ValueError:
        match = re.match(r'linear\((.+),(.+),(.+)\)', schdl)
        if match:
            init, final, duration = [float(g) for g in match.groups()]
            mix = np.clip(step / duration, 0.0, 1.0)
            return float((1.0 - mix) * init + mix * final)
        match = re.match(r'step_linear\((.+),(.+),(.+),(.+),(.+)\)', schdl)
        if match:
            init, final1, duration1, final2, duration2 = [float(g) for g in match.groups()]
            if step <= duration1:
                mix = np.clip(step / duration1, 0.0, 1.0)
                return float((1.0 - mix) * init + mix * final1)
            mix = np.clip((step - duration1) / duration2, 0.0, 1.0)
            return float((1.0 - mix) * final1 + mix * final2)
    raise NotImplementedError(schdl)


def weight_init(m: nn.Module) -> None:
    """utils.weight_init (utils.py:81-87) for the Linear layers of the FB networks."""
    if isinstance(m, nn.Linear):
        nn.init.orthogonal_(m.weight.data)
        if m.bias is not None:
            m.bias.data.fill_(0.0)


def soft_update_params(net: nn.Module, target_net: nn.Module, tau: float) -> None:
    for param, target_param in zip(net.parameters(), target_net.parameters()):
        target_param.data.copy_(tau * param.data + (1 - tau) * target_param.data)


def hard_update_params(net: nn.Module, target_net: nn.Module) -> None:
    for param, target_param in zip(net.parameters(), target_net.parameters()):
        target_param.data.copy_(param.data)


class TruncatedNormal(pyd.Normal):
    """utils.TruncatedNormal (utils.py:164-185): clipped noise, value clamp with straight-through gradient."""

    def __init__(self, loc: torch.Tensor, scale: torch.Tensor, low: float = -1.0, high: float = 1.0, eps: float = 1e-6) -> None:
        super().__init__(loc, scale, validate_args=False)
        self.low, self.high, self.eps = low, high, eps

    def _clamp(self, x: torch.Tensor) -> torch.Tensor:
        clamped_x = torch.clamp(x, self.low + self.eps, self.high - self.eps)
        return x - x.detach() + clamped_x.detach()

    def sample(self, clip: tp.Optional[float] = None, sample_shape: torch.Size = torch.Size()) -> torch.Tensor:  # type: ignore
        shape = self._extended_shape(sample_shape)
        eps = _standard_normal(shape, dtype=self.loc.dtype, device=self.loc.device)
        eps *= self.scale
        if clip is not None:
            eps = torch.clamp(eps, -clip, clip)
        return self._clamp(self.loc + eps)


def mlp(*layers: tp.Union[int, str]) -> nn.Sequential:
    """fb_modules.mlp (fb_modules.py:60-78): ints are Linear widths, strings name the non-linearity."""
    assert len(layers) >= 2 and isinstance(layers[0], int)
    seq: tp.List[nn.Module] = []
    prev = layers[0]
    for layer in layers[1:]:
        if isinstance(layer, str):
            if layer == "irelu":
                seq.append(nn.ReLU(inplace=True))
            elif layer == "relu":
                seq.append(nn.ReLU())
            elif layer == "ntanh":
                seq.extend([nn.LayerNorm(prev), nn.Tanh()])
            else:
                raise ValueError(f"Unknown non-linearity {layer}")
        else:
            seq.append(nn.Linear(prev, layer))
            prev = layer
    return nn.Sequential(*seq)


class Actor(nn.Module):
    """fb_modules.Actor with preprocess=True (fb_modules.py:81-126)."""

    def __init__(self, obs_dim: int, z_dim: int, action_dim: int, feature_dim: int, hidden_dim: int, add_trunk: bool = False,
                 preprocess: bool = True) -> None:
        super().__init__()
        self.obs_dim, self.z_dim, self.action_dim, self.preprocess = obs_dim, z_dim, action_dim, preprocess
        if preprocess:
            self.obs_net = mlp(obs_dim, hidden_dim, "ntanh", feature_dim, "irelu")
            self.obs_z_net = mlp(obs_dim + z_dim, hidden_dim, "ntanh", feature_dim, "irelu")
            self.trunk: nn.Module = mlp(2 * feature_dim, hidden_dim, "irelu") if add_trunk else nn.Identity()
            head_in = hidden_dim if add_trunk else 2 * feature_dim
        else:   # fb_modules.py:102-104
            self.trunk = mlp(obs_dim + z_dim, hidden_dim, "ntanh", hidden_dim, "irelu", hidden_dim, "irelu")
            head_in = hidden_dim
        self.policy = mlp(head_in, hidden_dim, "irelu", action_dim)
        self.apply(weight_init)

    def forward(self, obs: torch.Tensor, z: torch.Tensor, std: float) -> TruncatedNormal:
        assert z.shape[-1] == self.z_dim
        if self.preprocess:
            obs_z = self.obs_z_net(torch.cat([obs, z], dim=-1))
            o = self.obs_net(obs)
            h = torch.cat([o, obs_z], dim=-1)
        else:
            h = torch.cat([obs, z], dim=-1)
        mu = torch.tanh(self.policy(self.trunk(h)))
        return TruncatedNormal(mu, torch.ones_like(mu) * std)


class ForwardMap(nn.Module):
    """fb_modules.ForwardMap with preprocess=True (fb_modules.py:154-199)."""

    def __init__(self, obs_dim: int, z_dim: int, action_dim: int, feature_dim: int, hidden_dim: int, add_trunk: bool = False,
                 preprocess: bool = True) -> None:
        super().__init__()
        self.obs_dim, self.z_dim, self.action_dim, self.preprocess = obs_dim, z_dim, action_dim, preprocess
        if preprocess:
            self.obs_action_net = mlp(obs_dim + action_dim, hidden_dim, "ntanh", feature_dim, "irelu")
            self.obs_z_net = mlp(obs_dim + z_dim, hidden_dim, "ntanh", feature_dim, "irelu")
            self.trunk: nn.Module = mlp(2 * feature_dim, hidden_dim, "irelu") if add_trunk else nn.Identity()
            head_in = hidden_dim if add_trunk else 2 * feature_dim
        else:   # fb_modules.py:175-177
            self.trunk = mlp(obs_dim + z_dim + action_dim, hidden_dim, "ntanh", hidden_dim, "irelu", hidden_dim, "irelu")
            head_in = hidden_dim
        self.F1 = mlp(head_in, hidden_dim, "irelu", z_dim)
        self.F2 = mlp(head_in, hidden_dim, "irelu", z_dim)
        self.apply(weight_init)

    def forward(self, obs: torch.Tensor, z: torch.Tensor, action: torch.Tensor) -> tp.Tuple[torch.Tensor, torch.Tensor]:
        assert z.shape[-1] == self.z_dim
        if self.preprocess:
            oa = self.obs_action_net(torch.cat([obs, action], dim=-1))
            oz = self.obs_z_net(torch.cat([obs, z], dim=-1))
            h = torch.cat([oa, oz], dim=-1)
        else:
            h = torch.cat([obs, z, action], dim=-1)
        h = self.trunk(h)
        return self.F1(h), self.F2(h)


class BackwardMap(nn.Module):
    """fb_modules.BackwardMap (fb_modules.py:211-230)."""

    def __init__(self, obs_dim: int, z_dim: int, hidden_dim: int, norm_z: bool = True) -> None:
        super().__init__()
        self.obs_dim, self.z_dim, self.norm_z = obs_dim, z_dim, norm_z
        self.B = mlp(obs_dim, hidden_dim, "ntanh", hidden_dim, "relu", z_dim)
        self.apply(weight_init)

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        b = self.B(obs)
        if self.norm_z:
            b = math.sqrt(self.z_dim) * F.normalize(b, dim=1)
        return b


def adopt_flat(module: nn.Module, views: tp.Mapping[str, torch.Tensor]) -> None:
    """Move a freshly initialised CPU module onto the flat device segment: copy each parameter into its view and
    re-point the parameter at the view, keeping nn.Module registration order == flat layout order."""
    named = list(module.named_parameters())
    assert [n for n, _ in named] == list(views.keys()), ([n for n, _ in named], list(views.keys()))
    for name, p in named:
        v = views[name]
        assert tuple(p.shape) == tuple(v.shape), (name, tuple(p.shape), tuple(v.shape))
        v.copy_(p.data.to(v.device))
        p.data = v
