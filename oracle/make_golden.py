"""TEST INFRASTRUCTURE ONLY — generate `tests/golden/*.npz` from the UNMODIFIED reference.

Run in the build container (where `/root/reference` exists):

    python -m oracle.make_golden

Every array written here is an input to, or an output of, the reference's own code
(`FBDDPGAgent.update_fb / update_actor / update`, `ReplayBuffer.add / sample`) imported through
`oracle/ref_shim.py`.  The fixtures use reduced layer widths so that parameters, inputs and outputs
fit in a few hundred kB; the widths are deliberately awkward (not multiples of 4) to exercise the
padding paths of the CUDA kernels.  The GPU box has no `/root/reference`; tests there read only the
committed `.npz` files.
"""
from __future__ import annotations

import dataclasses
import os
import sys
import typing as tp

import numpy as np
import torch

from oracle import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES: tp.Dict[str, tp.Dict[str, tp.Any]] = {
    # name: agent dims + batch
    "small": dict(obs_dim=11, action_dim=3, z_dim=10, hidden_dim=48, feature_dim=24, backward_hidden_dim=30,
                  batch_size=32, goal_space=None, seed=11),
    "goal": dict(obs_dim=24, action_dim=6, z_dim=16, hidden_dim=64, feature_dim=32, backward_hidden_dim=46,
                 batch_size=64, goal_space="simplified_walker", seed=23),
    "wide": dict(obs_dim=24, action_dim=6, z_dim=50, hidden_dim=128, feature_dim=64, backward_hidden_dim=70,
                 batch_size=96, goal_space=None, seed=37),
}


def make_agent(R: tp.Any, case: tp.Mapping[str, tp.Any], use_tb: bool = True) -> tp.Any:
    cfg = R.FBDDPGAgentConfig(
        obs_type="states", obs_shape=(case["obs_dim"],), action_shape=(case["action_dim"],), device="cpu",
        use_tb=use_tb, use_wandb=False, use_hiplog=False, num_expl_steps=0, update_encoder=False,
        goal_space=case["goal_space"], hidden_dim=case["hidden_dim"], feature_dim=case["feature_dim"],
        backward_hidden_dim=case["backward_hidden_dim"], z_dim=case["z_dim"], batch_size=case["batch_size"],
        future_ratio=case.get("future_ratio", 0.0), q_loss=case.get("q_loss", False),
        q_loss_coef=case.get("q_loss_coef", 0.01), additional_metric=case.get("additional_metric", False),
        norm_z=case.get("norm_z", True), rand_weight=case.get("rand_weight", False),
        add_trunk=case.get("add_trunk", False), preprocess=case.get("preprocess", True), boltzmann=case.get("boltzmann", False),
        temp=case.get("temp", 1), debug=case.get("debug", False))
    return R.FBDDPGAgent(**dataclasses.asdict(cfg))


def _net(net: torch.nn.Module, prefix: str, out: tp.Dict[str, np.ndarray]) -> None:
    for name, p in net.named_parameters():
        out[f"{prefix}/{name}"] = p.detach().numpy().copy()


def _grads(net: torch.nn.Module, prefix: str, out: tp.Dict[str, np.ndarray]) -> None:
    for name, p in net.named_parameters():
        out[f"{prefix}/{name}"] = p.grad.detach().numpy().copy()


def gen_update_case(R: tp.Any, name: str, case: tp.Mapping[str, tp.Any]) -> tp.Dict[str, np.ndarray]:
    """One update_fb + update_actor + soft-update from a known snapshot with explicit noise."""
    torch.manual_seed(case["seed"])
    np.random.seed(case["seed"])
    agent = make_agent(R, case)
    B, O, A, Z = case["batch_size"], case["obs_dim"], case["action_dim"], case["z_dim"]
    G = ref_shim.GOAL_SPACE_DIMS[case["goal_space"]] if case["goal_space"] else O
    g = torch.Generator().manual_seed(case["seed"] + 1)
    # de-correlate targets from the online nets (they start as exact copies)
    with torch.no_grad():
        for net in (agent.forward_target_net, agent.backward_target_net):
            for p in net.parameters():
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        # non-trivial LayerNorm affine and biases everywhere
        for net in (agent.actor, agent.forward_net, agent.backward_net):
            for p in net.parameters():
                if p.dim() == 1:
                    p.add_(0.1 * torch.randn(p.shape, generator=g))
    obs = torch.randn(B, O, generator=g)
    next_obs = torch.randn(B, O, generator=g)
    action = torch.rand(B, A, generator=g) * 2 - 1
    discount = 0.98 * (torch.rand(B, 1, generator=g) > 0.05).float()   # a few terminal (0) discounts
    next_goal = torch.randn(B, G, generator=g) if case["goal_space"] else next_obs
    z = agent.sample_z(B)

    out: tp.Dict[str, np.ndarray] = {}
    for k, v in dict(obs=obs, next_obs=next_obs, action=action, discount=discount, next_goal=next_goal, z=z).items():
        out[f"in/{k}"] = v.numpy().copy()
    for net_name in ("actor", "forward_net", "backward_net", "forward_target_net", "backward_target_net"):
        _net(getattr(agent, net_name), f"param0/{net_name}", out)

    # ---- update_fb: record the N(0,1) draw the reference will make, then replay the seed ----------
    torch.manual_seed(case["seed"] + 2)
    out["in/noise_fb"] = torch.empty(B, A).normal_().numpy().copy()
    torch.manual_seed(case["seed"] + 2)
    # grads before the optimizer step: run backward through a spy on fb_opt.step
    grads_fb: tp.Dict[str, np.ndarray] = {}
    real_step = agent.fb_opt.step

    def spy_fb(*a: tp.Any, **k: tp.Any) -> tp.Any:
        _grads(agent.forward_net, "grad_fb/forward_net", grads_fb)
        _grads(agent.backward_net, "grad_fb/backward_net", grads_fb)
        return real_step(*a, **k)

    agent.fb_opt.step = spy_fb
    m_fb = agent.update_fb(obs=obs, action=action, discount=discount, next_obs=next_obs, next_goal=next_goal,
                           z=z, step=0)
    agent.fb_opt.step = real_step
    out.update(grads_fb)
    for k, v in m_fb.items():
        out[f"metric_fb/{k}"] = np.float64(v)
    _net(agent.forward_net, "param1/forward_net", out)
    _net(agent.backward_net, "param1/backward_net", out)

    # ---- update_actor (uses the just-updated forward_net, fb_ddpg.py:497) -------------------------
    torch.manual_seed(case["seed"] + 3)
    out["in/noise_actor"] = torch.empty(B, A).normal_().numpy().copy()
    torch.manual_seed(case["seed"] + 3)
    grads_actor: tp.Dict[str, np.ndarray] = {}
    real_astep = agent.actor_opt.step

    def spy_actor(*a: tp.Any, **k: tp.Any) -> tp.Any:
        _grads(agent.actor, "grad_actor/actor", grads_actor)
        return real_astep(*a, **k)

    agent.actor_opt.step = spy_actor
    m_actor = agent.update_actor(obs, z, 0)
    agent.actor_opt.step = real_astep
    out.update(grads_actor)
    for k, v in m_actor.items():
        out[f"metric_actor/{k}"] = np.float64(v)
    _net(agent.actor, "param1/actor", out)

    # ---- target tracking ----------------------------------------------------------------------
    R.utils.soft_update_params(agent.forward_net, agent.forward_target_net, agent.cfg.fb_target_tau)
    R.utils.soft_update_params(agent.backward_net, agent.backward_target_net, agent.cfg.fb_target_tau)
    _net(agent.forward_target_net, "param1/forward_target_net", out)
    _net(agent.backward_target_net, "param1/backward_target_net", out)

    # ---- a second fb+actor step so Adam's m/v and step>1 bias correction are pinned too ----------
    torch.manual_seed(case["seed"] + 4)
    out["in/noise_fb2"] = torch.empty(B, A).normal_().numpy().copy()
    torch.manual_seed(case["seed"] + 4)
    m_fb2 = agent.update_fb(obs=obs, action=action, discount=discount, next_obs=next_obs, next_goal=next_goal,
                            z=z, step=1)
    out["metric_fb2/fb_loss"] = np.float64(m_fb2["fb_loss"])
    _net(agent.forward_net, "param2/forward_net", out)
    _net(agent.backward_net, "param2/backward_net", out)
    out["cfg/stddev"] = np.float64(0.2)
    out["cfg/stddev_clip"] = np.float64(agent.cfg.stddev_clip)
    out["cfg/lr"] = np.float64(agent.cfg.lr)
    out["cfg/tau"] = np.float64(agent.cfg.fb_target_tau)
    out["cfg/ortho_coef"] = np.float64(agent.cfg.ortho_coef)
    if agent.cfg.q_loss:
        out["cfg/q_loss_coef"] = np.float64(agent.cfg.q_loss_coef)
    out["cfg/norm_z"] = np.int64(agent.cfg.norm_z)
    out["cfg/boltzmann"] = np.int64(agent.cfg.boltzmann)
    out["cfg/temp"] = np.float64(agent.cfg.temp)
    return out


def _episodes(rng: np.random.RandomState, lengths: tp.Sequence[int], obs_dim: int, act_dim: int, goal_dim: int,
              z_dim: int) -> tp.List[tp.Dict[str, np.ndarray]]:
    eps = []
    for n in lengths:
        ep = {"observation": rng.standard_normal((n + 1, obs_dim)).astype(np.float32),
              "action": rng.uniform(-1, 1, (n + 1, act_dim)).astype(np.float32),
              "reward": rng.uniform(0, 1, (n + 1,)).astype(np.float32),
              "discount": (rng.uniform(0, 1, (n + 1,)) > 0.1).astype(np.float32),
              "physics": rng.standard_normal((n + 1, 2)).astype(np.float32),
              "z": rng.standard_normal((n + 1, z_dim)).astype(np.float32)}
        if goal_dim:
            ep["goal"] = rng.standard_normal((n + 1, goal_dim)).astype(np.float32)
        eps.append(ep)
    return eps


def fill_reference_buffer(R: tp.Any, buf: tp.Any, episodes: tp.Sequence[tp.Mapping[str, np.ndarray]]) -> None:
    """Drive the reference's own `add()` one time step at a time (pretrain.py:573,607,649)."""
    for ep in episodes:
        n = len(ep["reward"])
        for t in range(n):
            st = R.StepType.FIRST if t == 0 else (R.StepType.LAST if t == n - 1 else R.StepType.MID)
            kw = dict(step_type=st, reward=float(ep["reward"][t]), discount=float(ep["discount"][t]),
                      observation=ep["observation"][t], action=ep["action"][t])
            if "goal" in ep:
                ts = R.dmc.ExtendedGoalTimeStep(goal=ep["goal"][t], **kw)
            else:
                ts = R.dmc.ExtendedTimeStep(**kw)
            ts.physics = ep["physics"][t]
            buf.add(ts, {"z": ep["z"][t]})


def gen_replay_case(R: tp.Any, name: str, lengths: tp.Sequence[int], max_episodes: int, goal_dim: int,
                    max_episode_length: tp.Optional[int], future: float, seed: int) -> tp.Dict[str, np.ndarray]:
    rng = np.random.RandomState(seed)
    episodes = _episodes(rng, lengths, obs_dim=5, act_dim=2, goal_dim=goal_dim, z_dim=3)
    buf = R.ReplayBuffer(max_episodes=max_episodes, discount=0.98, future=future, max_episode_length=max_episode_length)
    fill_reference_buffer(R, buf, episodes)
    out: tp.Dict[str, np.ndarray] = {"n_episodes": np.int64(len(episodes)), "max_episodes": np.int64(max_episodes),
                                     "future": np.float64(future), "seed": np.int64(seed),
                                     "max_episode_length": np.int64(-1 if max_episode_length is None else max_episode_length),
                                     "len": np.int64(len(buf)), "full": np.int64(buf._full),
                                     "fixed": np.int64(buf._is_fixed_episode_length),
                                     "avg_episode_length": np.int64(buf.avg_episode_length),
                                     "episodes_length": buf._episodes_length.copy()}
    for i, ep in enumerate(episodes):
        for k, v in ep.items():
            out[f"ep{i}/{k}"] = v
    for draw in range(3):
        np.random.seed(seed + 100 + draw)
        batch = buf.sample(16)
        for field in ("obs", "action", "reward", "discount", "next_obs", "goal", "next_goal", "future_obs", "future_goal"):
            v = getattr(batch, field)
            if v is not None:
                out[f"draw{draw}/{field}"] = np.asarray(v)
        for k, v in batch.meta.items():
            out[f"draw{draw}/meta/{k}"] = np.asarray(v)
    return out


def gen_trajectory_case(R: tp.Any, case: tp.Mapping[str, tp.Any], steps: int = 3) -> tp.Dict[str, np.ndarray]:
    """`agent.update(replay, step)` end to end (all six RNG draws live), metrics per step."""
    seed = case["seed"] + 50
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent = make_agent(R, case)
    goal_dim = ref_shim.GOAL_SPACE_DIMS[case["goal_space"]] if case["goal_space"] else 0
    rng = np.random.RandomState(seed)
    eps = []
    for _ in range(4):
        n = 40
        ep = {"observation": rng.standard_normal((n + 1, case["obs_dim"])).astype(np.float32),
              "action": rng.uniform(-1, 1, (n + 1, case["action_dim"])).astype(np.float32),
              "reward": rng.uniform(0, 1, (n + 1,)).astype(np.float32),
              "discount": np.ones((n + 1,), np.float32),
              "physics": np.zeros((n + 1, 2), np.float32),
              "z": rng.standard_normal((n + 1, case["z_dim"])).astype(np.float32)}
        if goal_dim:
            ep["goal"] = rng.standard_normal((n + 1, goal_dim)).astype(np.float32)
        eps.append(ep)
    buf = R.ReplayBuffer(max_episodes=4, discount=0.98, future=0.99)
    fill_reference_buffer(R, buf, eps)
    out: tp.Dict[str, np.ndarray] = {"seed": np.int64(seed), "steps": np.int64(steps)}
    for i, ep in enumerate(eps):
        for k, v in ep.items():
            out[f"ep{i}/{k}"] = v
    for net_name in ("actor", "forward_net", "backward_net"):
        _net(getattr(agent, net_name), f"param0/{net_name}", out)
    agent.cfg.update_every_steps = 1
    torch.manual_seed(seed + 1)
    np.random.seed(seed + 1)
    for step in range(steps):
        m = agent.update(buf, step)
        for k, v in m.items():
            out[f"step{step}/{k}"] = np.float64(v)
    for net_name in ("actor", "forward_net", "backward_net", "forward_target_net", "backward_target_net"):
        _net(getattr(agent, net_name), f"paramN/{net_name}", out)
    return out


def write_qloss_cases(R: tp.Any) -> None:
    """q_loss=True (fb_ddpg.py:330-341) and additional_metric=True (q1_success, fb_ddpg.py:403-404): added after the first
    fixtures, generated on their own.  q_loss_coef is raised from its 0.01 default so that the term carries weight in the
    gradients the fixtures pin."""
    for name, base in (("qloss", "wide"), ("qloss_goal", "goal")):
        case = dict(CASES[base], q_loss=True, q_loss_coef=0.5, additional_metric=True)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"update_{name}.npz"), **gen_update_case(R, name, case))
        print("wrote update_%s" % name)
    case = dict(CASES["small"], q_loss=True, q_loss_coef=0.5, additional_metric=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "trajectory_qloss.npz"), **gen_trajectory_case(R, case))
    print("wrote trajectory_qloss")


def write_randw_cases(R: tp.Any) -> None:
    """rand_weight=True (fb_ddpg.py:475-482): mixed z = random weighted sums of backward_net rows.  With norm_z=True the row scale
    cancels in the re-projection, so a norm_z=False trajectory pins the scale as well."""
    for name, extra in (("randw", dict()), ("randw_nonorm", dict(norm_z=False))):
        case = dict(CASES["small"], rand_weight=True, **extra)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"trajectory_{name}.npz"), **gen_trajectory_case(R, case))
        print("wrote trajectory_%s" % name)


def write_nonorm_cases(R: tp.Any) -> None:
    """norm_z=False (fb_modules.py:227-229, fb_ddpg.py:228-231,483-484): raw backward_net outputs, z = sqrt(Z) U(0,1) (x) direction,
    mixed z not re-projected; the diagonal orthonormality term then reaches the gradients."""
    for name, base in (("nonorm", "small"), ("nonorm_goal", "goal")):
        case = dict(CASES[base], norm_z=False)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"update_{name}.npz"), **gen_update_case(R, name, case))
        print("wrote update_%s" % name)
    case = dict(CASES["small"], norm_z=False, future_ratio=0.3)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "trajectory_nonorm.npz"), **gen_trajectory_case(R, case))
    print("wrote trajectory_nonorm")


def write_hindsight_cases(R: tp.Any) -> None:
    """hindsight trajectories (future_ratio > 0, fb_ddpg.py:488-491)."""
    for name, base in (("future", "small"), ("future_goal", "goal")):
        case = dict(CASES[base], future_ratio=0.4)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"trajectory_{name}.npz"), **gen_trajectory_case(R, case))
        print("wrote trajectory_%s" % name)


def write_trunk_cases(R: tp.Any) -> None:
    """add_trunk=True (fb_modules.py:96-100,169-173): a Linear + ReLU trunk between the embeds and the policy / F heads."""
    for name, base in (("trunk", "small"), ("trunk_goal", "goal")):
        case = dict(CASES[base], add_trunk=True)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"update_{name}.npz"), **gen_update_case(R, name, case))
        print("wrote update_%s" % name)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "trajectory_trunk.npz"), **gen_trajectory_case(R, dict(CASES["small"], add_trunk=True)))
    print("wrote trajectory_trunk")


def write_oracle_only_cases(R: tp.Any) -> None:
    """preprocess=False (one deep trunk, fb_modules.py:102-104,175-177) and boltzmann=True (DiagGaussianActor + SquashedNormal,
    fb_ddpg.py:118-120,304-306,391-393,406): these pin the ORACLE for the two branches the CUDA step does not implement yet
    (the agent raises NotImplementedError for them); the kernels of a later round are to be checked against these files."""
    for name, extra in (("nopre", dict(preprocess=False)), ("boltz", dict(boltzmann=True, temp=0.7))):
        case = dict(CASES["small"], **extra)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"update_{name}.npz"), **gen_update_case(R, name, case))
        print("wrote update_%s" % name)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"trajectory_{name}.npz"), **gen_trajectory_case(R, case))
        print("wrote trajectory_%s" % name)


def write_debug_cases(R: tp.Any) -> None:
    """debug=True (fb_ddpg.py:128-130, fb_modules.py:202-208): backward_net / backward_target_net are IdentityMap, so z lives in goal space
    (z_dim == goal_dim = obs_dim here) and backward_net has no parameters."""
    case = dict(CASES["small"], debug=True, z_dim=CASES["small"]["obs_dim"])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "update_debug.npz"), **gen_update_case(R, "debug", case))
    print("wrote update_debug")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "trajectory_debug.npz"), **gen_trajectory_case(R, case))
    print("wrote trajectory_debug")


# fixture families added after the first ones; `--<name>-only` regenerates one family without touching the others
LATER_FAMILIES = {"hindsight": write_hindsight_cases, "qloss": write_qloss_cases, "nonorm": write_nonorm_cases, "randw": write_randw_cases,
                  "trunk": write_trunk_cases, "oracle": write_oracle_only_cases, "debug": write_debug_cases}


def main() -> None:
    R = ref_shim.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)   # single-thread reductions: the most reproducible reference numbers
    only = [n for n in LATER_FAMILIES if f"--{n}-only" in sys.argv]
    for name, fn in LATER_FAMILIES.items():
        if not only or name in only:
            fn(R)
    if only:
        return
    for name, case in CASES.items():
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"update_{name}.npz"), **gen_update_case(R, name, case))
        print("wrote update_%s" % name)
    replay_cases = {
        "fixed": dict(lengths=[6] * 5, max_episodes=8, goal_dim=0, max_episode_length=None, future=0.99, seed=5),
        "fixed_goal_full": dict(lengths=[7] * 9, max_episodes=6, goal_dim=3, max_episode_length=None, future=0.9, seed=6),
        "ragged": dict(lengths=[5, 9, 3, 9, 7], max_episodes=6, goal_dim=2, max_episode_length=10, future=0.8, seed=7),
        "nofuture": dict(lengths=[4] * 3, max_episodes=3, goal_dim=0, max_episode_length=None, future=1.0, seed=8),
    }
    for name, kw in replay_cases.items():
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"replay_{name}.npz"), **gen_replay_case(R, name, **kw))
        print("wrote replay_%s" % name)
    for name in ("small", "goal"):
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"trajectory_{name}.npz"), **gen_trajectory_case(R, CASES[name]))
        print("wrote trajectory_%s" % name)


if __name__ == "__main__":
    main()
