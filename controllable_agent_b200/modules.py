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
from torch import nn
from torch.distributions.utils import _standard_normal


_SCHEDULE_CALL = re.compile(r"^\s*(linear|step_linear)\((.*)\)\s*$")


def _schedule_knots(spec: str) -> tp.Tuple[tp.List[float], tp.List[float]]:
    """(steps, values) of the piecewise-linear curve a schedule string describes: `linear(a,b,n)` goes from a to b over n steps,
    `step_linear(a,b,n,c,m)` continues from b to c over the m steps after n (the grammar of utils.schedule, utils.py:235-255)."""
    m = _SCHEDULE_CALL.match(spec)
    if m is None:
        raise NotImplementedError(spec)
    args = [float(x) for x in m.group(2).split(",")]
    if m.group(1) == "linear" and len(args) == 3:
        a, b, n = args
        return [0.0, n], [a, b]
    if m.group(1) == "step_linear" and len(args) == 5:
        a, b, n, c, k = args
        return [0.0, n, n + k], [a, b, c]
    raise NotImplementedError(spec)


def schedule(schdl: tp.Any, step: int) -> float:
    """Value of a schedule spec at `step`: a number (or numeric string) is constant, otherwise piecewise-linear between the knots of
    the spec and flat outside them."""
    try:
        return float(schdl)
    except (TypeError, ValueError):
        xs, ys = _schedule_knots(str(schdl))
        return float(np.interp(float(step), xs, ys))


def weight_init(m: nn.Module) -> None:
    """Initialisation of the FB networks' Linear layers: orthogonal weight, zero bias (what utils.weight_init does for nn.Linear,
    utils.py:81-87; one orthogonal_ draw per layer, in registration order, so seeded construction matches the reference)."""
    if isinstance(m, nn.Linear):
        nn.init.orthogonal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


def _param_pairs(net: nn.Module, target_net: nn.Module) -> tp.Tuple[tp.List[torch.Tensor], tp.List[torch.Tensor]]:
    src, dst = [p.data for p in net.parameters()], [p.data for p in target_net.parameters()]
    if len(src) != len(dst):
        raise ValueError("networks with different parameter lists")
    return src, dst


def soft_update_params(net: nn.Module, target_net: nn.Module, tau: float) -> None:
    """target <- tau * net + (1 - tau) * target, parameter by parameter in registration order (host-side twin of the lerp fused into
    k_adam; utils.py:66-69).  The agent's step does this on the device; this exists for callers that poke the modules."""
    src, dst = _param_pairs(net, target_net)
    if not src:   # parameter-free modules (the states-only agent's nn.Identity encoder)
        return
    with torch.no_grad():   # two rounded products and one rounded sum per element (no fused multiply-add): bit-identical to the reference
        mixed = torch._foreach_add(torch._foreach_mul(src, tau), torch._foreach_mul(dst, 1.0 - tau))
        torch._foreach_copy_(dst, mixed)


def hard_update_params(net: nn.Module, target_net: nn.Module) -> None:
    """target <- net (utils.py:72-74), what init_from uses to adopt another agent's networks."""
    src, dst = _param_pairs(net, target_net)
    if not src:
        return
    with torch.no_grad():
        torch._foreach_copy_(dst, src)


class TruncatedNormal:
    """The action distribution the actor returns: N(loc, scale) whose samples are clipped noise around loc, clamped into
    (low, high) with a straight-through gradient (the contract of utils.TruncatedNormal, utils.py:164-185, which the CUDA step
    implements in k_actor_out).  Only what callers of `agent.actor(...)` use is provided: mean / loc / scale / stddev, sample(clip),
    log_prob."""

    def __init__(self, loc: torch.Tensor, scale: torch.Tensor, low: float = -1.0, high: float = 1.0, eps: float = 1e-6) -> None:
        self.loc, self.scale = loc, scale
        self._lo, self._hi = low + eps, high - eps

    @property
    def mean(self) -> torch.Tensor:
        return self.loc

    @property
    def stddev(self) -> torch.Tensor:
        return self.scale

    def sample(self, clip: tp.Optional[float] = None, sample_shape: torch.Size = torch.Size()) -> torch.Tensor:
        shape = torch.Size(sample_shape) + self.loc.shape
        noise = _standard_normal(shape, dtype=self.loc.dtype, device=self.loc.device) * self.scale   # one draw per element, like the reference
        if clip is not None:
            noise = noise.clamp(-clip, clip)
        x = self.loc + noise
        inside = x.clamp(self._lo, self._hi).detach()
        return inside + (x - x.detach())   # value: the clamped sample, exactly; gradient w.r.t. loc: identity

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        var = self.scale ** 2
        return -((value - self.loc) ** 2) / (2 * var) - torch.log(self.scale) - 0.5 * math.log(2 * math.pi)


class SquashedNormal:
    """tanh of a Normal(loc, scale) draw: the action distribution of the DiagGaussianActor (cfg.boltzmann; the contract of
    utils.SquashedNormal, utils.py:188-233).  `mean` is tanh(loc) (the transform applied to the base mean, not the true mean);
    log_prob(y) = Normal.log_prob(x) - log|d tanh / dx|(x) with x = atanh(y) and the Jacobian in its softplus form, which is what
    k_actor_out_bz evaluates from the pre-tanh sample it keeps."""

    def __init__(self, loc: torch.Tensor, scale: torch.Tensor) -> None:
        self.loc, self.scale = loc, scale

    @property
    def mean(self) -> torch.Tensor:
        return torch.tanh(self.loc)

    def rsample(self, sample_shape: torch.Size = torch.Size()) -> torch.Tensor:
        shape = torch.Size(sample_shape) + self.loc.shape
        return torch.tanh(self.loc + self.scale * _standard_normal(shape, dtype=self.loc.dtype, device=self.loc.device))

    def sample(self, sample_shape: torch.Size = torch.Size()) -> torch.Tensor:
        shape = torch.Size(sample_shape) + self.loc.shape
        with torch.no_grad():   # Normal.sample: torch.normal on the expanded parameters
            return torch.tanh(torch.normal(self.loc.expand(shape), self.scale.expand(shape)))

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        x = 0.5 * (torch.log1p(value) - torch.log1p(-value))
        base = -((x - self.loc) ** 2) / (2 * self.scale ** 2) - torch.log(self.scale) - 0.5 * math.log(2 * math.pi)
        return base - 2.0 * (math.log(2.0) - x - F.softplus(-2.0 * x))


_ACTIVATIONS: tp.Dict[str, tp.Callable[[int], tp.List[nn.Module]]] = {
    "irelu": lambda width: [nn.ReLU(inplace=True)],
    "relu": lambda width: [nn.ReLU()],
    "ntanh": lambda width: [nn.LayerNorm(width), nn.Tanh()],
}


def mlp(*layers: tp.Union[int, str]) -> nn.Sequential:
    """Sequential stack from a width / activation list in the reference's notation (fb_modules.py:60-78): the first entry is the
    input width, every further int adds a Linear to that width, every string an activation from _ACTIVATIONS ("ntanh" = LayerNorm
    + Tanh, "irelu" = in-place ReLU).  Module indices — and therefore parameter names like `0.weight`, `1.bias`, `3.weight` — come
    out as the reference's, which is what the flat parameter layout of the library is keyed on."""
    if len(layers) < 2 or not isinstance(layers[0], int):
        raise ValueError("mlp(in_width, ...) needs an input width and at least one layer")
    width = layers[0]
    stack: tp.List[nn.Module] = []
    for item in layers[1:]:
        if isinstance(item, int):
            stack.append(nn.Linear(width, item))
            width = item
        elif item in _ACTIVATIONS:
            stack.extend(_ACTIVATIONS[item](width))
        else:
            raise ValueError(f"Unknown non-linearity {item}")
    return nn.Sequential(*stack)


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


class DiagGaussianActor(nn.Module):
    """fb_modules.DiagGaussianActor (cfg.boltzmann, fb_modules.py:129-151): one stack on [obs | z] whose output is
    [mu | raw log-std]; the log-std is squashed into log_std_bounds with a tanh."""

    def __init__(self, obs_dim: int, z_dim: int, action_dim: int, hidden_dim: int, log_std_bounds: tp.Tuple[float, float]) -> None:
        super().__init__()
        self.obs_dim, self.z_dim, self.action_dim, self.log_std_bounds = obs_dim, z_dim, action_dim, tuple(log_std_bounds)
        self.policy = mlp(obs_dim + z_dim, hidden_dim, "ntanh", hidden_dim, "relu", 2 * action_dim)
        self.apply(weight_init)

    def forward(self, obs: torch.Tensor, z: torch.Tensor) -> SquashedNormal:
        assert z.shape[-1] == self.z_dim
        mu, raw = self.policy(torch.cat([obs, z], dim=-1)).chunk(2, dim=-1)
        lo, hi = self.log_std_bounds
        return SquashedNormal(mu, torch.exp(lo + 0.5 * (hi - lo) * (torch.tanh(raw) + 1.0)))


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


class IdentityMap(nn.Module):
    """fb_modules.IdentityMap (cfg.debug, fb_modules.py:202-208): the backward representation is the goal itself."""

    def __init__(self) -> None:
        super().__init__()
        self.B = nn.Identity()

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return self.B(obs)


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
