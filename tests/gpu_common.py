"""Helpers shared by the `-m gpu` parity tests: build an engine from a golden fixture / oracle parameter
set and compare against the oracle (CPU).  Test infrastructure only."""
import collections

import numpy as np
import torch

from conftest import subtree
from oracle import fb_oracle as O

REL_TOL = 1e-3   # BASELINE.json north_star: losses and grads within 1e-3 relative fp32


def rel(a, b):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def dims_from_params(fwd, bwd, actor):
    if len(bwd) == 0:   # cfg.debug: identity backward map, z_dim == goal_dim == obs_dim; the (unused) hidden width is make_golden's
        hidden, oa = fwd["obs_action_net.0.weight"].shape
        obs_dim = actor["obs_net.0.weight"].shape[1]
        return O.Dims(obs_dim=obs_dim, action_dim=oa - obs_dim, z_dim=fwd["F1.2.weight"].shape[0], goal_dim=obs_dim, hidden_dim=hidden,
                      feature_dim=fwd["obs_action_net.3.weight"].shape[0], backward_hidden_dim=30)
    if "policy.5.weight" in actor:   # cfg.boltzmann: the DiagGaussianActor has no embeds; the obs width comes from policy.0 ([obs | z] columns)
        hidden, oa = fwd["obs_action_net.0.weight"].shape
        z_dim = fwd["F1.2.weight"].shape[0]
        obs_dim = actor["policy.0.weight"].shape[1] - z_dim
        return O.Dims(obs_dim=obs_dim, action_dim=oa - obs_dim, z_dim=z_dim, goal_dim=bwd["B.0.weight"].shape[1], hidden_dim=hidden,
                      feature_dim=fwd["obs_action_net.3.weight"].shape[0], backward_hidden_dim=bwd["B.0.weight"].shape[0])
    if "obs_action_net.0.weight" not in fwd:   # preprocess = False: the feature width does not exist; taken from make_golden.CASES["small"]
        hidden, oza = fwd["trunk.0.weight"].shape
        oz = actor["trunk.0.weight"].shape[1]
        z_dim = fwd["F1.2.weight"].shape[0]
        return O.Dims(obs_dim=oz - z_dim, action_dim=oza - oz, z_dim=z_dim, goal_dim=bwd["B.0.weight"].shape[1], hidden_dim=hidden,
                      feature_dim=24, backward_hidden_dim=bwd["B.0.weight"].shape[0])
    hidden, oa = fwd["obs_action_net.0.weight"].shape
    obs_dim = actor["obs_net.0.weight"].shape[1]
    return O.Dims(obs_dim=obs_dim, action_dim=oa - obs_dim, z_dim=fwd["F1.2.weight"].shape[0],
                  goal_dim=bwd["B.0.weight"].shape[1], hidden_dim=hidden,
                  feature_dim=fwd["obs_action_net.3.weight"].shape[0], backward_hidden_dim=bwd["B.0.weight"].shape[0])


def golden_params(g, prefix):
    return collections.OrderedDict((k, torch.from_numpy(np.array(v))) for k, v in subtree(g, prefix).items())


def make_engine(d, batch, use_goal=False, rng_device=False, mix_ratio=0.5, ortho_coef=1.0, seed=0, global_batch=None,
                row_offset=0, contract_mode=0, mlp_mode=0, q_loss_coef=None, norm_z=True, add_trunk=False, fused=False, preprocess=True, boltzmann=False, temp=1.0, debug=False):
    from controllable_agent_b200.engine import EngineConfig, FBStepEngine
    cfg = EngineConfig(batch=batch, obs_dim=d.obs_dim, action_dim=d.action_dim, z_dim=d.z_dim, goal_dim=d.goal_dim,
                       hidden_dim=d.hidden_dim, feature_dim=d.feature_dim, backward_hidden_dim=d.backward_hidden_dim,
                       use_goal=use_goal, rng_device=rng_device, ortho_coef=ortho_coef, mix_ratio=mix_ratio, seed=seed,
                       global_batch=global_batch, row_offset=row_offset, contract_mode=contract_mode, mlp_mode=mlp_mode,
                       q_loss=q_loss_coef is not None, q_loss_coef=q_loss_coef if q_loss_coef is not None else 0.01, norm_z=norm_z, add_trunk=add_trunk, fused=fused, preprocess=preprocess,
                       boltzmann=boltzmann, temp=temp, debug=debug)
    return FBStepEngine(cfg, "cuda")


def load_params(eng, fwd=None, bwd=None, actor=None, fwd_tgt=None, bwd_tgt=None):
    from controllable_agent_b200 import _lib as L
    for net, which, src in ((L.NET_FORWARD, "param", fwd), (L.NET_BACKWARD, "param", bwd), (L.NET_ACTOR, "param", actor),
                            (L.NET_FORWARD, "target", fwd_tgt), (L.NET_BACKWARD, "target", bwd_tgt)):
        if src is None:
            continue
        views = eng.tensors(net, which)
        assert list(views.keys()) == list(src.keys()), (list(views.keys()), list(src.keys()))
        for k, v in views.items():
            v.copy_(torch.as_tensor(np.asarray(src[k])).to(v.device))


def read_tensors(eng, net, which):
    return collections.OrderedDict((k, v.detach().cpu().clone()) for k, v in eng.tensors(net, which).items())
