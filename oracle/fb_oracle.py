"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference FB-DDPG gradient step.

This file is the *checker*, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The product path
(`controllable_agent_b200`) never imports `oracle/` and has no CPU route.

It restates, in plain functional torch-CPU fp32 (torch/numpy are the libraries the reference itself
computes with; requirements.txt:2-3, unpinned), the algorithm of

    url_benchmark/agent/fb_ddpg.py:224-232   sample_z
    url_benchmark/agent/fb_ddpg.py:291-387   update_fb
    url_benchmark/agent/fb_ddpg.py:389-421   update_actor
    url_benchmark/agent/fb_ddpg.py:427-520   update (RNG consumption order, z mixing)
    url_benchmark/agent/fb_modules.py:43-230 mlp / Actor / ForwardMap / BackwardMap
    url_benchmark/utils.py:66-69,81-87,164-185,235-255  soft update, init, TruncatedNormal, schedule
    url_benchmark/in_memory_replay_buffer.py:104-190    add / sample

Parity pinning: `tests/test_oracle_golden.py` checks every function here against fixtures in
`tests/golden/` produced by the UNMODIFIED reference (run under `oracle/ref_shim.py` by
`oracle/make_golden.py`), and — in the build container, where `/root/reference` exists — against the
live reference on fresh seeds.  The reference's own tests hold no numeric known-answer vectors for
this path, so those reference-generated fixtures are the pin.
"""
from __future__ import annotations

import collections
import dataclasses
import math
import re
import typing as tp

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = tp.Dict[str, Tensor]

LN_EPS = 1e-5          # nn.LayerNorm default, fb_modules.py:49-50
NORMALIZE_EPS = 1e-12  # F.normalize default, fb_modules.py:39
CLAMP_EPS = 1e-6       # TruncatedNormal eps, utils.py:165


# ------------------------------------------------------------------------------------------------
# dimensions and parameter inventories (registration order == nn.Module.parameters() order,
# which is what init_from / soft_update_params zip over: fb_ddpg.py:168-172, utils.py:66-74)
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class Dims:
    obs_dim: int = 24
    action_dim: int = 6
    z_dim: int = 50
    goal_dim: int = 24            # == obs_dim when goal_space is None (fb_ddpg.py:112-114)
    hidden_dim: int = 1024
    feature_dim: int = 512
    backward_hidden_dim: int = 526


def _embed_spec(prefix: str, in_dim: int, d: Dims) -> tp.List[tp.Tuple[str, tp.Tuple[int, ...]]]:
    """mlp(in, hidden, "ntanh", feature, "irelu")  (fb_modules.py:92-93,165-166)."""
    return [(f"{prefix}.0.weight", (d.hidden_dim, in_dim)), (f"{prefix}.0.bias", (d.hidden_dim,)),
            (f"{prefix}.1.weight", (d.hidden_dim,)), (f"{prefix}.1.bias", (d.hidden_dim,)),
            (f"{prefix}.3.weight", (d.feature_dim, d.hidden_dim)), (f"{prefix}.3.bias", (d.feature_dim,))]


def _head_spec(prefix: str, out_dim: int, d: Dims, add_trunk: bool = False) -> tp.List[tp.Tuple[str, tp.Tuple[int, ...]]]:
    """mlp(2*feature, hidden, "irelu", out), or mlp(hidden, hidden, "irelu", out) behind a trunk  (fb_modules.py:96-107,169-185)."""
    in_dim = d.hidden_dim if add_trunk else 2 * d.feature_dim
    return [(f"{prefix}.0.weight", (d.hidden_dim, in_dim)), (f"{prefix}.0.bias", (d.hidden_dim,)),
            (f"{prefix}.2.weight", (out_dim, d.hidden_dim)), (f"{prefix}.2.bias", (out_dim,))]


def _trunk_spec(d: Dims, add_trunk: bool) -> tp.List[tp.Tuple[str, tp.Tuple[int, ...]]]:
    """add_trunk=True: trunk = mlp(2*feature, hidden, "irelu"), registered between the embeds and the heads (fb_modules.py:99-100,172-173)."""
    return [("trunk.0.weight", (d.hidden_dim, 2 * d.feature_dim)), ("trunk.0.bias", (d.hidden_dim,))] if add_trunk else []


def _deep_trunk_spec(in_dim: int, d: Dims) -> tp.List[tp.Tuple[str, tp.Tuple[int, ...]]]:
    """preprocess=False: trunk = mlp(in, hidden, "ntanh", hidden, "irelu", hidden, "irelu")  (fb_modules.py:102-104,175-177)."""
    h = d.hidden_dim
    return [("trunk.0.weight", (h, in_dim)), ("trunk.0.bias", (h,)), ("trunk.1.weight", (h,)), ("trunk.1.bias", (h,)),
            ("trunk.3.weight", (h, h)), ("trunk.3.bias", (h,)), ("trunk.5.weight", (h, h)), ("trunk.5.bias", (h,))]


def forward_map_spec(d: Dims, add_trunk: bool = False, preprocess: bool = True) -> tp.List[tp.Tuple[str, tp.Tuple[int, ...]]]:
    if not preprocess:
        return (_deep_trunk_spec(d.obs_dim + d.z_dim + d.action_dim, d)
                + _head_spec("F1", d.z_dim, d, True) + _head_spec("F2", d.z_dim, d, True))
    return (_embed_spec("obs_action_net", d.obs_dim + d.action_dim, d)
            + _embed_spec("obs_z_net", d.obs_dim + d.z_dim, d) + _trunk_spec(d, add_trunk)
            + _head_spec("F1", d.z_dim, d, add_trunk) + _head_spec("F2", d.z_dim, d, add_trunk))


def boltzmann_actor_spec(d: Dims) -> tp.List[tp.Tuple[str, tp.Tuple[int, ...]]]:
    """DiagGaussianActor (boltzmann=True): policy = mlp(obs + z, hidden, "ntanh", hidden, "relu", 2 * action)  (fb_modules.py:129-139)."""
    h = d.hidden_dim
    return [("policy.0.weight", (h, d.obs_dim + d.z_dim)), ("policy.0.bias", (h,)), ("policy.1.weight", (h,)), ("policy.1.bias", (h,)),
            ("policy.3.weight", (h, h)), ("policy.3.bias", (h,)), ("policy.5.weight", (2 * d.action_dim, h)), ("policy.5.bias", (2 * d.action_dim,))]


def actor_spec(d: Dims, add_trunk: bool = False, preprocess: bool = True) -> tp.List[tp.Tuple[str, tp.Tuple[int, ...]]]:
    if not preprocess:
        return _deep_trunk_spec(d.obs_dim + d.z_dim, d) + _head_spec("policy", d.action_dim, d, True)
    return (_embed_spec("obs_net", d.obs_dim, d) + _embed_spec("obs_z_net", d.obs_dim + d.z_dim, d) + _trunk_spec(d, add_trunk)
            + _head_spec("policy", d.action_dim, d, add_trunk))


def backward_map_spec(d: Dims) -> tp.List[tp.Tuple[str, tp.Tuple[int, ...]]]:
    """mlp(goal, Hb, "ntanh", Hb, "relu", z)  (fb_modules.py:220)."""
    hb = d.backward_hidden_dim
    return [("B.0.weight", (hb, d.goal_dim)), ("B.0.bias", (hb,)), ("B.1.weight", (hb,)), ("B.1.bias", (hb,)),
            ("B.3.weight", (hb, hb)), ("B.3.bias", (hb,)), ("B.5.weight", (d.z_dim, hb)), ("B.5.bias", (d.z_dim,))]


def init_params(spec: tp.Sequence[tp.Tuple[str, tp.Tuple[int, ...]]],
                generator: tp.Optional[torch.Generator] = None) -> Params:
    """utils.weight_init (utils.py:81-87): orthogonal Linear weights, zero biases; LayerNorm keeps
    its default affine (ones / zeros).  A 1-D '.weight' is a LayerNorm scale."""
    out: Params = collections.OrderedDict()
    for name, shape in spec:
        if len(shape) == 2:
            w = torch.empty(shape, dtype=torch.float32)
            torch.nn.init.orthogonal_(w, generator=generator)
            out[name] = w
        elif name.endswith(".weight"):
            out[name] = torch.ones(shape, dtype=torch.float32)
        else:
            out[name] = torch.zeros(shape, dtype=torch.float32)
    return out


# ------------------------------------------------------------------------------------------------
# network forwards (fb_modules.py)
# ------------------------------------------------------------------------------------------------
def _embed(p: Params, prefix: str, x: Tensor) -> Tensor:
    """Linear -> LayerNorm -> Tanh -> Linear -> ReLU."""
    h = F.linear(x, p[f"{prefix}.0.weight"], p[f"{prefix}.0.bias"])
    h = F.layer_norm(h, (h.shape[-1],), p[f"{prefix}.1.weight"], p[f"{prefix}.1.bias"], LN_EPS)
    h = torch.tanh(h)
    return torch.relu(F.linear(h, p[f"{prefix}.3.weight"], p[f"{prefix}.3.bias"]))


def _head(p: Params, prefix: str, h: Tensor) -> Tensor:
    """Linear -> ReLU -> Linear."""
    h = torch.relu(F.linear(h, p[f"{prefix}.0.weight"], p[f"{prefix}.0.bias"]))
    return F.linear(h, p[f"{prefix}.2.weight"], p[f"{prefix}.2.bias"])


def _deep_trunk(p: Params, x: Tensor) -> Tensor:
    """preprocess=False trunk: Linear -> LayerNorm -> Tanh -> Linear -> ReLU -> Linear -> ReLU (fb_modules.py:102-104,175-177)."""
    h = F.linear(x, p["trunk.0.weight"], p["trunk.0.bias"])
    h = torch.tanh(F.layer_norm(h, (h.shape[-1],), p["trunk.1.weight"], p["trunk.1.bias"], LN_EPS))
    h = torch.relu(F.linear(h, p["trunk.3.weight"], p["trunk.3.bias"]))
    return torch.relu(F.linear(h, p["trunk.5.weight"], p["trunk.5.bias"]))


def _trunk(p: Params, h: Tensor) -> Tensor:
    """The optional trunk Linear -> ReLU (add_trunk=True), the identity otherwise (fb_modules.py:96-100,118-119)."""
    if "trunk.0.weight" in p:
        return torch.relu(F.linear(h, p["trunk.0.weight"], p["trunk.0.bias"]))
    return h


def forward_map(p: Params, obs: Tensor, z: Tensor, action: Tensor) -> tp.Tuple[Tensor, Tensor]:
    """ForwardMap.forward (fb_modules.py:187-199); the parameter names tell the preprocess=False variant apart."""
    if "obs_action_net.0.weight" not in p:
        h = _deep_trunk(p, torch.cat([obs, z, action], dim=-1))
        return _head(p, "F1", h), _head(p, "F2", h)
    oa = _embed(p, "obs_action_net", torch.cat([obs, action], dim=-1))
    oz = _embed(p, "obs_z_net", torch.cat([obs, z], dim=-1))
    h = _trunk(p, torch.cat([oa, oz], dim=-1))
    return _head(p, "F1", h), _head(p, "F2", h)


def actor_mean(p: Params, obs: Tensor, z: Tensor) -> Tensor:
    """Actor.forward up to mu = tanh(policy(h)) (fb_modules.py:110-122); preprocess=False is told apart by the parameter names."""
    if "obs_net.0.weight" not in p:
        return torch.tanh(_head(p, "policy", _deep_trunk(p, torch.cat([obs, z], dim=-1))))
    oz = _embed(p, "obs_z_net", torch.cat([obs, z], dim=-1))
    o = _embed(p, "obs_net", obs)
    return torch.tanh(_head(p, "policy", _trunk(p, torch.cat([o, oz], dim=-1))))


def diag_gaussian_actor(p: Params, obs: Tensor, z: Tensor, log_std_bounds: tp.Tuple[float, float] = (-5, 2)) -> tp.Tuple[Tensor, Tensor]:
    """DiagGaussianActor.forward (boltzmann=True, fb_modules.py:141-151): (mu, std) of the pre-tanh Normal."""
    h = F.linear(torch.cat([obs, z], dim=-1), p["policy.0.weight"], p["policy.0.bias"])
    h = torch.tanh(F.layer_norm(h, (h.shape[-1],), p["policy.1.weight"], p["policy.1.bias"], LN_EPS))
    h = torch.relu(F.linear(h, p["policy.3.weight"], p["policy.3.bias"]))
    mu, log_std = F.linear(h, p["policy.5.weight"], p["policy.5.bias"]).chunk(2, dim=-1)
    lo, hi = log_std_bounds
    log_std = lo + 0.5 * (hi - lo) * (torch.tanh(log_std) + 1)
    return mu, log_std.exp()


def squashed_normal_log_prob(x: Tensor, mu: Tensor, std: Tensor) -> Tensor:
    """SquashedNormal.log_prob(tanh(x)) from the pre-tanh value x (what TanhTransform's cache hands back, utils.py:188-233):
    Normal(mu, std).log_prob(x) - log|d tanh(x)/dx|, the Jacobian in the stable form of utils.py:212-215."""
    base = -((x - mu) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
    return base - 2.0 * (math.log(2.0) - x - F.softplus(-2.0 * x))


def l2_project(x: Tensor, z_dim: int) -> Tensor:
    """sqrt(z_dim) * F.normalize(x, dim=1)."""
    return math.sqrt(z_dim) * F.normalize(x, dim=1, eps=NORMALIZE_EPS)


def backward_map(p: Params, goal: Tensor, z_dim: int, norm_z: bool = True) -> Tensor:
    """BackwardMap.forward (fb_modules.py:223-230).  A parameter-free map is the IdentityMap of cfg.debug (fb_modules.py:202-208,
    fb_ddpg.py:128-130): B(goal) = goal, no projection."""
    if len(p) == 0:
        return goal
    h = F.linear(goal, p["B.0.weight"], p["B.0.bias"])
    h = torch.tanh(F.layer_norm(h, (h.shape[-1],), p["B.1.weight"], p["B.1.bias"], LN_EPS))
    h = torch.relu(F.linear(h, p["B.3.weight"], p["B.3.bias"]))
    b = F.linear(h, p["B.5.weight"], p["B.5.bias"])
    return l2_project(b, z_dim) if norm_z else b


def truncated_normal_sample(mu: Tensor, noise: Tensor, std: float, clip: tp.Optional[float]) -> Tensor:
    """TruncatedNormal.sample (utils.py:176-185) with the N(0,1) draw `noise` made explicit.
    Value is clamp(mu + clamp(noise*std, +-clip), +-(1-1e-6)); gradient w.r.t. mu is the identity
    (straight-through, utils.py:171-174)."""
    eps = noise * std
    if clip is not None:
        eps = torch.clamp(eps, -clip, clip)
    x = mu + eps
    clamped = torch.clamp(x, -1.0 + CLAMP_EPS, 1.0 - CLAMP_EPS)
    return x - x.detach() + clamped.detach()


def normal_log_prob(x: Tensor, mu: Tensor, std: float) -> Tensor:
    """pyd.Normal.log_prob — only feeds the `actor_logprob` metric (fb_ddpg.py:399,418)."""
    var = std * std
    return -((x - mu) ** 2) / (2 * var) - math.log(std) - math.log(math.sqrt(2 * math.pi))


def schedule(schdl: tp.Any, step: int) -> float:
    """utils.schedule (utils.py:235-255)."""
    try:
        return float(schdl)
    except ValueError:
        m = re.match(r"linear\((.+),(.+),(.+)\)", schdl)
        if m:
            init, final, duration = (float(g) for g in m.groups())
            mix = float(np.clip(step / duration, 0.0, 1.0))
            return (1.0 - mix) * init + mix * final
        m = re.match(r"step_linear\((.+),(.+),(.+),(.+),(.+)\)", schdl)
        if m:
            init, final1, duration1, final2, duration2 = (float(g) for g in m.groups())
            if step <= duration1:
                mix = float(np.clip(step / duration1, 0.0, 1.0))
                return (1.0 - mix) * init + mix * final1
            mix = float(np.clip((step - duration1) / duration2, 0.0, 1.0))
            return (1.0 - mix) * final1 + mix * final2
    raise NotImplementedError(schdl)


def sample_z(batch: int, z_dim: int, generator: tp.Optional[torch.Generator] = None, norm_z: bool = True) -> Tensor:
    """sample_z (fb_ddpg.py:224-232): uniform on the sqrt(z_dim)-sphere; norm_z=False scales every coordinate of the
    direction by its own U(0,1) draw (made after the normal draw, on the same generator)."""
    g = F.normalize(torch.randn((batch, z_dim), dtype=torch.float32, generator=generator), dim=1)
    if norm_z:
        return math.sqrt(z_dim) * g
    u = torch.rand((batch, z_dim), dtype=torch.float32, generator=generator)
    return np.sqrt(z_dim) * u * g


# ------------------------------------------------------------------------------------------------
# losses (fb_ddpg.py:303-348, 389-406)
# ------------------------------------------------------------------------------------------------
def q_loss_term(F1: Tensor, F2: Tensor, Bm: Tensor, tF1: Tensor, tF2: Tensor, discount: Tensor, z: Tensor) -> Tensor:
    """The optional Q loss of update_fb (q_loss=True, fb_ddpg.py:330-340): F_k . z regressed on the implicit
    reward B cov^-1 z plus the discounted target Q; everything on the target side is under no_grad."""
    with torch.no_grad():
        next_Q = torch.min(torch.einsum("sd, sd -> s", tF1, z), torch.einsum("sd, sd -> s", tF2, z))
        cov = torch.matmul(Bm.T, Bm) / Bm.shape[0]
        inv_cov = torch.inverse(cov)
        implicit_reward = (torch.matmul(Bm, inv_cov) * z).sum(dim=1)
        target_Q = implicit_reward.detach() + discount.squeeze(1) * next_Q
    Q1, Q2 = (torch.einsum("sd, sd -> s", Fi, z) for Fi in (F1, F2))
    return F.mse_loss(Q1, target_Q) + F.mse_loss(Q2, target_Q)


def fb_loss_terms(F1: Tensor, F2: Tensor, Bm: Tensor, tF1: Tensor, tF2: Tensor, tB: Tensor,
                  discount: Tensor, ortho_coef: float, z: tp.Optional[Tensor] = None,
                  q_loss_coef: tp.Optional[float] = None) -> tp.Dict[str, Tensor]:
    """The batch x batch successor-measure loss and orthonormality regulariser, written the way the
    reference writes it (einsum + boolean off-diagonal mask).  `q_loss_coef` not None = cfg.q_loss
    (needs `z`): adds q_loss_coef * q_loss_term(...) to fb_loss and reports it as "q_loss"."""
    target_M = torch.min(torch.einsum("sd, td -> st", tF1, tB), torch.einsum("sd, td -> st", tF2, tB))
    M1 = torch.einsum("sd, td -> st", F1, Bm)
    M2 = torch.einsum("sd, td -> st", F2, Bm)
    eye = torch.eye(*M1.size(), device=M1.device)
    off_diag = ~eye.bool()
    fb_offdiag = 0.5 * sum((M - discount * target_M)[off_diag].pow(2).mean() for M in (M1, M2))
    fb_diag = -sum(M.diag().mean() for M in (M1, M2))
    cov = torch.matmul(Bm, Bm.T)
    orth_diag = -2 * cov.diag().mean()
    orth_offdiag = cov[off_diag].pow(2).mean()
    orth = orth_offdiag + orth_diag
    fb_loss = fb_offdiag + fb_diag
    out: tp.Dict[str, Tensor] = {}
    if q_loss_coef is not None:   # added before the orthonormality term, like the reference (fb_ddpg.py:341,348)
        assert z is not None
        out["q_loss"] = q_loss_term(F1, F2, Bm, tF1, tF2, discount, z)
        fb_loss = fb_loss + q_loss_coef * out["q_loss"]
    out.update({"fb_loss": fb_loss + ortho_coef * orth, "fb_offdiag": fb_offdiag, "fb_diag": fb_diag,
                "orth_loss": orth, "orth_loss_diag": orth_diag, "orth_loss_offdiag": orth_offdiag,
                "target_M": target_M, "M1": M1})
    return out


def fb_metrics(terms: tp.Dict[str, Tensor], F1: Tensor, Bm: Tensor, z: Tensor) -> tp.Dict[str, float]:
    """The metric block of update_fb (fb_ddpg.py:356-377)."""
    eye_diff = torch.matmul(Bm.T, Bm) / Bm.shape[0] - torch.eye(Bm.shape[1], dtype=Bm.dtype)
    extra = {"q_loss": terms["q_loss"].item()} if "q_loss" in terms else {}
    return {**extra, "target_M": terms["target_M"].mean().item(), "M1": terms["M1"].mean().item(),
            "F1": F1.mean().item(), "B": Bm.mean().item(),
            "B_norm": torch.norm(Bm, dim=-1).mean().item(), "z_norm": torch.norm(z, dim=-1).mean().item(),
            "fb_loss": terms["fb_loss"].item(), "fb_diag": terms["fb_diag"].item(),
            "fb_offdiag": terms["fb_offdiag"].item(), "orth_loss": terms["orth_loss"].item(),
            "orth_loss_diag": terms["orth_loss_diag"].item(),
            "orth_loss_offdiag": terms["orth_loss_offdiag"].item(),
            "orth_linf": torch.max(torch.abs(eye_diff)).item(),
            "orth_l2": eye_diff.norm().item() / math.sqrt(Bm.shape[1])}


def _with_grad(p: Params) -> Params:
    return collections.OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())


def fb_loss_and_grads(fwd: Params, bwd: Params, fwd_tgt: Params, bwd_tgt: Params, actor: Params,
                      obs: Tensor, action: Tensor, discount: Tensor, next_obs: Tensor, next_goal: Tensor,
                      z: Tensor, noise: Tensor, std: float, clip: tp.Optional[float], ortho_coef: float,
                      z_dim: int, q_loss_coef: tp.Optional[float] = None, norm_z: bool = True,
                      boltzmann: bool = False) -> tp.Dict[str, tp.Any]:
    """update_fb up to (not including) the optimizer step: loss terms, metrics, grads of every
    forward_net / backward_net tensor, and the intermediates a kernel test wants to look at."""
    with torch.no_grad():
        if boltzmann:   # dist.sample() of the SquashedNormal (fb_ddpg.py:304-306): tanh of a Normal draw, no clipping
            mu_b, std_b = diag_gaussian_actor(actor, next_obs, z)
            next_action = torch.tanh(mu_b + std_b * noise)
        else:
            next_action = truncated_normal_sample(actor_mean(actor, next_obs, z), noise, std, clip)
        tF1, tF2 = forward_map(fwd_tgt, next_obs, z, next_action)
        tB = backward_map(bwd_tgt, next_goal, z_dim, norm_z)
    f, b = _with_grad(fwd), _with_grad(bwd)
    F1, F2 = forward_map(f, obs, z, action)
    Bm = backward_map(b, next_goal, z_dim, norm_z)
    if not Bm.requires_grad:   # the identity map of cfg.debug: B is the input itself; its gradient is still reported (dL/dB)
        Bm = Bm.detach().clone().requires_grad_(True)
    F1.retain_grad(), F2.retain_grad(), Bm.retain_grad()
    terms = fb_loss_terms(F1, F2, Bm, tF1, tF2, tB, discount, ortho_coef, z, q_loss_coef)
    terms["fb_loss"].backward()
    return {"terms": {k: v.detach() for k, v in terms.items()},
            "metrics": fb_metrics({k: v.detach() for k, v in terms.items()}, F1.detach(), Bm.detach(), z),
            "grads_forward": collections.OrderedDict((k, v.grad) for k, v in f.items()),
            "grads_backward": collections.OrderedDict((k, v.grad) for k, v in b.items()),
            "next_action": next_action, "tF1": tF1, "tF2": tF2, "tB": tB,
            "F1": F1.detach(), "F2": F2.detach(), "B": Bm.detach(),
            "dF1": F1.grad, "dF2": F2.grad, "dB": Bm.grad}


def actor_loss_and_grads(actor: Params, fwd: Params, obs: Tensor, z: Tensor, noise: Tensor, std: float,
                         clip: tp.Optional[float], boltzmann: bool = False, temp: float = 1.0) -> tp.Dict[str, tp.Any]:
    """update_actor up to the optimizer step (fb_ddpg.py:389-409).  "q1_success": additional_metric (fb_ddpg.py:403-404).
    boltzmann: DiagGaussianActor + SquashedNormal.rsample, loss = mean(temp * log_prob - Q)."""
    a = _with_grad(actor)
    if boltzmann:
        mu, std_b = diag_gaussian_actor(a, obs, z)
        x = mu + std_b * noise
        action = torch.tanh(x)
        log_prob = squashed_normal_log_prob(x, mu, std_b).sum(-1, keepdim=True)
    else:
        mu = actor_mean(a, obs, z)
        action = truncated_normal_sample(mu, noise, std, clip)
        log_prob = normal_log_prob(action, mu, std).sum(-1, keepdim=True)
    F1, F2 = forward_map(fwd, obs, z, action)
    Q1 = torch.einsum("sd, sd -> s", F1, z)
    Q2 = torch.einsum("sd, sd -> s", F2, z)
    Q = torch.min(Q1, Q2)
    loss = (temp * log_prob - Q).mean() if boltzmann else -Q.mean()
    loss.backward()
    return {"actor_loss": loss.detach(), "q": Q.mean().detach(), "actor_logprob": log_prob.mean().detach(),
            "q1_success": (Q1 > Q2).float().mean().detach(),
            "grads_actor": collections.OrderedDict((k, v.grad) for k, v in a.items()),
            "action": action.detach(), "mu": mu.detach(), "Q1": Q1.detach(), "Q2": Q2.detach()}


# ------------------------------------------------------------------------------------------------
# optimizer + target tracking
# ------------------------------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam single-tensor update (amsgrad=False, weight_decay=0, maximize=False), the
    optimizer of fb_ddpg.py:146-151; `step` is the 1-based count AFTER this call."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def soft_update(net: Params, target: Params, tau: float) -> None:
    """utils.soft_update_params (utils.py:66-69)."""
    for k in net:
        target[k].copy_(tau * net[k] + (1 - tau) * target[k])


# ------------------------------------------------------------------------------------------------
# replay buffer (in_memory_replay_buffer.py)
# ------------------------------------------------------------------------------------------------
TIMESTEP_FIELDS = ("step_type", "reward", "discount", "observation", "physics", "goal", "action")


def sample_indices(n_episodes: int, episodes_length: np.ndarray, batch_size: int, future: float,
                   fixed_length: bool = True, rng: tp.Any = np.random
                   ) -> tp.Tuple[np.ndarray, np.ndarray, tp.Optional[np.ndarray]]:
    """Index draws of ReplayBuffer.sample (in_memory_replay_buffer.py:146-161), same RNG calls in
    the same order so that a shared numpy seed gives identical indices."""
    if fixed_length:
        ep_idx = rng.randint(0, n_episodes, size=batch_size)
    else:
        prob = episodes_length / episodes_length.sum()
        ep_idx = rng.choice(np.arange(len(episodes_length)), size=batch_size, p=prob)
    eps_lengths = episodes_length[ep_idx]
    step_idx = rng.randint(0, eps_lengths) + 1
    future_idx = None
    if future < 1:
        future_idx = step_idx + rng.geometric(p=(1 - future), size=batch_size)
        future_idx = np.clip(future_idx, 0, eps_lengths)
    return ep_idx, step_idx, future_idx


def gather_batch(storage: tp.Mapping[str, np.ndarray], ep_idx: np.ndarray, step_idx: np.ndarray,
                 future_idx: tp.Optional[np.ndarray], discount: float) -> tp.Dict[str, tp.Any]:
    """The fancy-index gathers of ReplayBuffer.sample (in_memory_replay_buffer.py:162-190)."""
    out: tp.Dict[str, tp.Any] = {
        "meta": {k: v[ep_idx, step_idx - 1] for k, v in storage.items() if k not in TIMESTEP_FIELDS},
        "obs": storage["observation"][ep_idx, step_idx - 1],
        "action": storage["action"][ep_idx, step_idx],
        "next_obs": storage["observation"][ep_idx, step_idx],
        "reward": storage["reward"][ep_idx, step_idx],
        "discount": discount * storage["discount"][ep_idx, step_idx],
        "goal": None, "next_goal": None, "future_obs": None, "future_goal": None}
    if "goal" in storage:
        out["goal"] = storage["goal"][ep_idx, step_idx - 1]
        out["next_goal"] = storage["goal"][ep_idx, step_idx]
        if future_idx is not None:
            out["future_goal"] = storage["goal"][ep_idx, future_idx - 1]
    if future_idx is not None:
        out["future_obs"] = storage["observation"][ep_idx, future_idx - 1]
    return out


class OracleReplay:
    """Episode-major host storage `[max_episodes, T+1, dim]` per field, filled episode by episode
    (what ReplayBuffer.add produces, in_memory_replay_buffer.py:104-133); fixed-length or ragged."""

    def __init__(self, max_episodes: int, discount: float, future: float) -> None:
        self.max_episodes, self.discount, self.future = max_episodes, discount, future
        self.storage: tp.Dict[str, np.ndarray] = {}
        self.episodes_length = np.zeros(max_episodes, dtype=np.int32)
        self.idx, self.full, self.fixed_length = 0, False, True

    def __len__(self) -> int:
        return self.max_episodes if self.full else self.idx

    def add_episode(self, episode: tp.Mapping[str, np.ndarray]) -> None:
        """`episode[name]` is `[len+1, dim]` including the dummy first row."""
        for name, values in episode.items():
            values = np.asarray(values, dtype=np.float32)
            if name not in self.storage:
                self.storage[name] = np.empty((self.max_episodes,) + values.shape, dtype=np.float32)
            self.storage[name][self.idx][:len(values)] = values
        n = len(episode["discount"]) - 1
        self.episodes_length[self.idx] = n
        prev = self.episodes_length[self.idx - 1]
        if n != prev and prev != 0:
            self.fixed_length = False
        self.idx = (self.idx + 1) % self.max_episodes
        self.full = self.full or self.idx == 0

    def sample(self, batch_size: int, rng: tp.Any = np.random) -> tp.Dict[str, tp.Any]:
        if self.fixed_length:
            idx = sample_indices(len(self), self.episodes_length, batch_size, self.future, True, rng)
        else:
            idx = sample_indices(len(self), self.episodes_length, batch_size, self.future, False, rng)
        out = gather_batch(self.storage, *idx, self.discount)
        out["_indices"] = idx
        return out


# ------------------------------------------------------------------------------------------------
# the whole agent step (fb_ddpg.py:427-520), for trajectory checks and as the timed CPU baseline
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class OracleConfig:
    dims: Dims = Dims()
    batch_size: int = 1024
    lr: float = 1e-4
    lr_coef: float = 1.0
    fb_target_tau: float = 0.01
    stddev_schedule: str = "0.2"
    stddev_clip: float = 0.3
    ortho_coef: float = 1.0
    mix_ratio: float = 0.5
    future_ratio: float = 0.0     # hindsight z (fb_ddpg.py:488-491)
    use_goal: bool = False        # goal_space is not None
    metrics: bool = True          # use_tb or use_wandb or use_hiplog
    q_loss: bool = False          # fb_ddpg.py:330-341
    q_loss_coef: float = 0.01
    additional_metric: bool = False   # q1_success (fb_ddpg.py:403-404,416-417)
    preprocess: bool = True       # False: one deep trunk on the concatenated inputs instead of the two embeds (fb_modules.py:102-104)
    boltzmann: bool = False       # DiagGaussianActor + SquashedNormal, entropy-regularised actor loss (fb_ddpg.py:118-120,304-306,391-393,406)
    temp: float = 1.0
    add_trunk: bool = False       # Linear(2 feature -> hidden) + ReLU between the embeds and the heads (fb_modules.py:96-100,169-173)
    rand_weight: bool = False     # mixed z = random convex-like combinations of B rows (fb_ddpg.py:475-482)
    norm_z: bool = True           # sqrt(z_dim)-sphere projection of B's output and of z (fb_modules.py:227-229, fb_ddpg.py:228,483)
    debug: bool = False           # backward_net = IdentityMap (fb_ddpg.py:128-130); needs z_dim == goal_dim


class OracleAgent:
    """The reference agent's update path on CPU.  Uses the same torch ops in the same order and the
    same RNG streams (numpy global for replay indices and the mix mask, torch CPU generator for z,
    randperm and the two action-noise draws — Appendix B of SURVEY.md), so that from equal
    parameters and equal seeds it walks the reference's trajectory."""

    def __init__(self, cfg: OracleConfig, generator: tp.Optional[torch.Generator] = None) -> None:
        self.cfg = cfg
        d = cfg.dims
        self.actor = init_params(boltzmann_actor_spec(d) if cfg.boltzmann else actor_spec(d, cfg.add_trunk, cfg.preprocess), generator)
        self.forward_net = init_params(forward_map_spec(d, cfg.add_trunk, cfg.preprocess), generator)
        self.backward_net = init_params([] if cfg.debug else backward_map_spec(d), generator)
        self.forward_target_net = collections.OrderedDict((k, v.clone()) for k, v in self.forward_net.items())
        self.backward_target_net = collections.OrderedDict((k, v.clone()) for k, v in self.backward_net.items())
        self._build_optimizers()

    def _build_optimizers(self) -> None:
        for net in (self.actor, self.forward_net, self.backward_net):
            for v in net.values():
                v.requires_grad_(True)
        cfg = self.cfg
        self.actor_opt = torch.optim.Adam(list(self.actor.values()), lr=cfg.lr)
        self.fb_opt = torch.optim.Adam([{"params": list(self.forward_net.values())},
                                        {"params": list(self.backward_net.values()), "lr": cfg.lr_coef * cfg.lr}],
                                       lr=cfg.lr)

    def load_params(self, **nets: tp.Mapping[str, tp.Any]) -> None:
        for net_name, values in nets.items():
            net = getattr(self, net_name)
            for k in net:
                net[k].data.copy_(torch.as_tensor(np.asarray(values[k])))

    # -- fb_ddpg.py:291-387 --------------------------------------------------------------------
    def update_fb(self, obs: Tensor, action: Tensor, discount: Tensor, next_obs: Tensor, next_goal: Tensor,
                  z: Tensor, step: int) -> tp.Dict[str, float]:
        cfg, d = self.cfg, self.cfg.dims
        std = schedule(cfg.stddev_schedule, step)
        with torch.no_grad():
            if cfg.boltzmann:
                mu, std_b = diag_gaussian_actor(self.actor, next_obs, z)
                next_action = torch.tanh(mu + std_b * torch.randn(mu.shape, dtype=mu.dtype))
            else:
                mu = actor_mean(self.actor, next_obs, z)
                noise = torch.randn(mu.shape, dtype=mu.dtype)
                next_action = truncated_normal_sample(mu, noise, std, cfg.stddev_clip)
            tF1, tF2 = forward_map(self.forward_target_net, next_obs, z, next_action)
            tB = backward_map(self.backward_target_net, next_goal, d.z_dim, cfg.norm_z)
        F1, F2 = forward_map(self.forward_net, obs, z, action)
        Bm = backward_map(self.backward_net, next_goal, d.z_dim, cfg.norm_z)
        terms = fb_loss_terms(F1, F2, Bm, tF1, tF2, tB, discount, cfg.ortho_coef, z, cfg.q_loss_coef if cfg.q_loss else None)
        metrics: tp.Dict[str, float] = {}
        if cfg.metrics:
            metrics = fb_metrics({k: v.detach() for k, v in terms.items()}, F1.detach(), Bm.detach(), z)
            metrics["fb_opt_lr"] = self.fb_opt.param_groups[0]["lr"]
        self.fb_opt.zero_grad(set_to_none=True)
        terms["fb_loss"].backward()
        self.fb_opt.step()
        return metrics

    # -- fb_ddpg.py:389-421 --------------------------------------------------------------------
    def update_actor(self, obs: Tensor, z: Tensor, step: int) -> tp.Dict[str, float]:
        cfg = self.cfg
        std = schedule(cfg.stddev_schedule, step)
        if cfg.boltzmann:
            mu, std_b = diag_gaussian_actor(self.actor, obs, z)
            x = mu + std_b * torch.randn(mu.shape, dtype=mu.dtype)
            action = torch.tanh(x)
            log_prob = squashed_normal_log_prob(x, mu, std_b).sum(-1, keepdim=True)
        else:
            mu = actor_mean(self.actor, obs, z)
            noise = torch.randn(mu.shape, dtype=mu.dtype)
            action = truncated_normal_sample(mu, noise, std, cfg.stddev_clip)
            log_prob = normal_log_prob(action, mu, std).sum(-1, keepdim=True)
        F1, F2 = forward_map(self.forward_net, obs, z, action)
        Q1, Q2 = torch.einsum("sd, sd -> s", F1, z), torch.einsum("sd, sd -> s", F2, z)
        Q = torch.min(Q1, Q2)
        loss = (cfg.temp * log_prob - Q).mean() if cfg.boltzmann else -Q.mean()
        self.actor_opt.zero_grad(set_to_none=True)
        loss.backward()       # like the reference, this also fills forward_net grads (never used)
        self.actor_opt.step()
        if cfg.metrics:
            out = {"actor_loss": loss.item(), "q": Q.mean().item(), "actor_logprob": log_prob.mean().item()}
            if cfg.additional_metric:
                out["q1_success"] = (Q1 > Q2).float().mean().item()
            return out
        return {}

    # -- fb_ddpg.py:427-520 --------------------------------------------------------------------
    def update_from_batch(self, batch: tp.Mapping[str, tp.Any], step: int) -> tp.Dict[str, float]:
        cfg, d = self.cfg, self.cfg.dims
        t = {k: torch.as_tensor(v) for k, v in batch.items() if isinstance(v, np.ndarray)}
        obs, action, discount, next_obs = t["obs"], t["action"], t["discount"], t["next_obs"]
        next_goal = t["next_goal"] if cfg.use_goal else next_obs
        backward_input = t["goal"] if cfg.use_goal else obs
        z = sample_z(cfg.batch_size, d.z_dim, norm_z=cfg.norm_z)
        perm = torch.randperm(cfg.batch_size)
        backward_input = backward_input[perm]
        if cfg.mix_ratio > 0:
            mix_idxs = np.where(np.random.uniform(size=cfg.batch_size) < cfg.mix_ratio)[0]
            with torch.no_grad():
                if cfg.rand_weight:   # fb_ddpg.py:475-482: rows of U(0,1) weights, L2-normalised, scaled by one U(0,1) each
                    weight = F.normalize(torch.rand(size=(mix_idxs.shape[0], cfg.batch_size)), dim=1)
                    weight = torch.rand(mix_idxs.shape[0], 1) * weight
                    mix_z = torch.matmul(weight, backward_map(self.backward_net, backward_input, d.z_dim, cfg.norm_z))
                else:
                    mix_z = backward_map(self.backward_net, backward_input[mix_idxs], d.z_dim, cfg.norm_z)
            z[mix_idxs] = l2_project(mix_z, d.z_dim) if cfg.norm_z else mix_z
        if cfg.future_ratio > 0:   # hindsight replay (fb_ddpg.py:488-491)
            future_goal = t["future_goal"] if cfg.use_goal else t["future_obs"]
            future_idxs = np.where(np.random.uniform(size=cfg.batch_size) < cfg.future_ratio)[0]
            with torch.no_grad():
                z[future_idxs] = backward_map(self.backward_net, future_goal[future_idxs], d.z_dim, cfg.norm_z)
        metrics = self.update_fb(obs, action, discount, next_obs, next_goal, z, step)
        metrics.update(self.update_actor(obs, z, step))
        with torch.no_grad():
            soft_update(self.forward_net, self.forward_target_net, cfg.fb_target_tau)
            soft_update(self.backward_net, self.backward_target_net, cfg.fb_target_tau)
        return metrics

    def update(self, replay: OracleReplay, step: int) -> tp.Dict[str, float]:
        return self.update_from_batch(replay.sample(self.cfg.batch_size), step)


def synthetic_episode(rng: np.random.RandomState, length: int, d: Dims, with_goal: bool = False,
                      meta_z: bool = False) -> tp.Dict[str, np.ndarray]:
    """One synthetic episode of SURVEY.md section 8d: observation ~ N(0,1), action ~ U(-1,1),
    reward ~ U(0,1), discount = 1, `[length+1, dim]` rows including the dummy first transition."""
    n = length + 1
    ep = {"observation": rng.standard_normal((n, d.obs_dim)).astype(np.float32),
          "action": rng.uniform(-1, 1, (n, d.action_dim)).astype(np.float32),
          "reward": rng.uniform(0, 1, (n, 1)).astype(np.float32),
          "discount": np.ones((n, 1), np.float32)}
    if with_goal:
        ep["goal"] = rng.standard_normal((n, d.goal_dim)).astype(np.float32)
    if meta_z:
        ep["z"] = rng.standard_normal((n, d.z_dim)).astype(np.float32)
    return ep
