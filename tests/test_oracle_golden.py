"""Pin the oracle restatement (oracle/fb_oracle.py) against fixtures produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only."""
import collections

import numpy as np
import pytest
import torch

from conftest import load_golden, subtree
from oracle import fb_oracle as O

UPDATE_CASES = ["small", "goal", "wide", "qloss", "qloss_goal", "nonorm", "nonorm_goal", "trunk", "trunk_goal"]   # trunk*: add_trunk=True; qloss*: q_loss=True (fb_ddpg.py:330-341); nonorm*: norm_z=False


def dims_from(g):
    f = subtree(g, "param0/forward_net")
    b = subtree(g, "param0/backward_net")
    a = subtree(g, "param0/actor")
    hidden, oa = f["obs_action_net.0.weight"].shape
    obs_dim = a["obs_net.0.weight"].shape[1]
    return O.Dims(obs_dim=obs_dim, action_dim=oa - obs_dim, z_dim=f["F1.2.weight"].shape[0],
                  goal_dim=b["B.0.weight"].shape[1], hidden_dim=hidden,
                  feature_dim=f["obs_action_net.3.weight"].shape[0], backward_hidden_dim=b["B.0.weight"].shape[0])


def params(g, prefix):
    return collections.OrderedDict((k, torch.from_numpy(v.copy())) for k, v in subtree(g, prefix).items())


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("case", UPDATE_CASES)
def test_spec_matches_reference_registration_order(case):
    g = load_golden(f"update_{case}")
    d = dims_from(g)
    trunk = case.startswith("trunk")
    for net, spec in (("forward_net", O.forward_map_spec(d, trunk)), ("backward_net", O.backward_map_spec(d)),
                      ("actor", O.actor_spec(d, trunk))):
        ref = subtree(g, f"param0/{net}")
        assert [n for n, _ in spec] == list(ref.keys())
        assert [tuple(s) for _, s in spec] == [v.shape for v in ref.values()]


@pytest.mark.parametrize("case", UPDATE_CASES)
def test_update_fb_losses_and_grads(case):
    g = load_golden(f"update_{case}")
    d = dims_from(g)
    t = {k: torch.from_numpy(v.copy()) for k, v in subtree(g, "in").items()}
    res = O.fb_loss_and_grads(
        params(g, "param0/forward_net"), params(g, "param0/backward_net"), params(g, "param0/forward_target_net"),
        params(g, "param0/backward_target_net"), params(g, "param0/actor"), t["obs"], t["action"], t["discount"],
        t["next_obs"], t["next_goal"], t["z"], t["noise_fb"], float(g["cfg/stddev"]), float(g["cfg/stddev_clip"]),
        float(g["cfg/ortho_coef"]), d.z_dim, float(g["cfg/q_loss_coef"]) if "cfg/q_loss_coef" in g else None,
        norm_z=bool(g["cfg/norm_z"]) if "cfg/norm_z" in g else True)
    assert (abs(res["metrics"]["orth_loss_diag"] + 2 * d.z_dim) > 1e-3) == case.startswith("nonorm")
    assert ("q_loss" in res["metrics"]) == case.startswith("qloss")
    for k, v in subtree(g, "metric_fb").items():
        if k == "fb_opt_lr":
            continue
        assert res["metrics"][k] == pytest.approx(float(v), rel=2e-5, abs=2e-6), k
    for net, key in (("forward_net", "grads_forward"), ("backward_net", "grads_backward")):
        for name, ref in subtree(g, f"grad_fb/{net}").items():
            assert rel(res[key][name].numpy(), ref) < 1e-5, (net, name)


@pytest.mark.parametrize("case", UPDATE_CASES)
def test_adam_actor_and_soft_update(case):
    g = load_golden(f"update_{case}")
    lr, tau = float(g["cfg/lr"]), float(g["cfg/tau"])
    # Adam step 1 on forward/backward nets from the reference's own grads
    for net in ("forward_net", "backward_net"):
        p0, p1, gr = params(g, f"param0/{net}"), subtree(g, f"param1/{net}"), subtree(g, f"grad_fb/{net}")
        for name in p0:
            p = p0[name].clone()
            O.adam_step(p, torch.from_numpy(gr[name]), torch.zeros_like(p), torch.zeros_like(p), 1, lr)
            assert np.abs(p.numpy() - p1[name]).max() < 2e-7, (net, name)
    # actor loss/grads use the just-updated forward_net
    t = {k: torch.from_numpy(v.copy()) for k, v in subtree(g, "in").items()}
    res = O.actor_loss_and_grads(params(g, "param0/actor"), params(g, "param1/forward_net"), t["obs"], t["z"],
                                 t["noise_actor"], float(g["cfg/stddev"]), float(g["cfg/stddev_clip"]))
    m = subtree(g, "metric_actor")
    assert float(res["actor_loss"]) == pytest.approx(float(m["actor_loss"]), rel=2e-5)
    assert float(res["q"]) == pytest.approx(float(m["q"]), rel=2e-5)
    assert float(res["actor_logprob"]) == pytest.approx(float(m["actor_logprob"]), rel=2e-5)
    if "q1_success" in m:   # additional_metric=True fixtures
        assert float(res["q1_success"]) == pytest.approx(float(m["q1_success"]), abs=1e-7)
    for name, ref in subtree(g, "grad_actor/actor").items():
        assert rel(res["grads_actor"][name].numpy(), ref) < 1e-5, name
    # soft update
    for net in ("forward", "backward"):
        tgt = params(g, f"param0/{net}_target_net")
        O.soft_update(params(g, f"param1/{net}_net"), tgt, tau)
        for name, ref in subtree(g, f"param1/{net}_target_net").items():
            assert np.abs(tgt[name].numpy() - ref).max() < 1e-7, name


def _rebuild_oracle_replay(g, discount=0.98):
    buf = O.OracleReplay(int(g["max_episodes"]), discount, float(g["future"]))
    mel = int(g["max_episode_length"])
    for i in range(int(g["n_episodes"])):
        ep = {k: (v if v.ndim > 1 else v[:, None]) for k, v in subtree(g, f"ep{i}").items()}
        if mel > 0 and not buf.storage:
            # max_episode_length pre-sizes the storage rows (in_memory_replay_buffer.py:123-125)
            for k, v in ep.items():
                buf.storage[k] = np.empty((buf.max_episodes, mel) + v.shape[1:], np.float32)
        buf.add_episode(ep)
    return buf


@pytest.mark.parametrize("case", ["fixed", "fixed_goal_full", "ragged", "nofuture"])
def test_replay_sample_bit_exact(case):
    g = load_golden(f"replay_{case}")
    buf = _rebuild_oracle_replay(g)
    assert len(buf) == int(g["len"]) and buf.full == bool(g["full"]) and buf.fixed_length == bool(g["fixed"])
    np.testing.assert_array_equal(buf.episodes_length, g["episodes_length"])
    for draw in range(3):
        np.random.seed(int(g["seed"]) + 100 + draw)
        batch = buf.sample(16)
        ref = subtree(g, f"draw{draw}")
        for field in ("obs", "action", "reward", "discount", "next_obs", "goal", "next_goal", "future_obs", "future_goal"):
            if field in ref:
                np.testing.assert_array_equal(batch[field], ref[field], err_msg=field)
            else:
                assert batch[field] is None, field
        np.testing.assert_array_equal(batch["meta"]["z"], ref["meta/z"])
        assert set(batch["meta"].keys()) == {k[5:] for k in ref if k.startswith("meta/")}


@pytest.mark.parametrize("case", ["small", "goal", "future", "future_goal", "qloss", "nonorm", "randw", "randw_nonorm", "trunk"])
def test_full_update_trajectory(case):
    """agent.update(replay, step) x3 with all RNG streams live: the oracle agent walks the
    reference's trajectory from the same parameters and seeds ("future*": hindsight z, future_ratio = 0.4)."""
    torch.set_num_threads(1)
    g = load_golden(f"trajectory_{case}")
    a = subtree(g, "param0/actor")
    f = subtree(g, "param0/forward_net")
    b = subtree(g, "param0/backward_net")
    hidden, oa = f["obs_action_net.0.weight"].shape
    obs_dim = a["obs_net.0.weight"].shape[1]
    d = O.Dims(obs_dim=obs_dim, action_dim=oa - obs_dim, z_dim=f["F1.2.weight"].shape[0], goal_dim=b["B.0.weight"].shape[1],
               hidden_dim=hidden, feature_dim=f["obs_action_net.3.weight"].shape[0], backward_hidden_dim=b["B.0.weight"].shape[0])
    use_goal = "ep0/goal" in g
    agent = O.OracleAgent(O.OracleConfig(dims=d, batch_size=64 if use_goal else 32, use_goal=use_goal,
                                         future_ratio=0.4 if case.startswith("future") else (0.3 if case == "nonorm" else 0.0),
                                         norm_z=not case.endswith("nonorm"), rand_weight=case.startswith("randw"), add_trunk=case == "trunk",
                                         q_loss=case == "qloss", q_loss_coef=0.5, additional_metric=case == "qloss"))
    agent.load_params(actor=a, forward_net=f, backward_net=b, forward_target_net=f, backward_target_net=b)
    buf = O.OracleReplay(4, 0.98, 0.99)
    for i in range(4):
        buf.add_episode({k: (v if v.ndim > 1 else v[:, None]) for k, v in subtree(g, f"ep{i}").items()})
    seed = int(g["seed"])
    torch.manual_seed(seed + 1)
    np.random.seed(seed + 1)
    for step in range(int(g["steps"])):
        m = agent.update(buf, step)
        ref = subtree(g, f"step{step}")
        assert set(ref) == set(m)
        for k, v in ref.items():
            assert m[k] == pytest.approx(float(v), rel=2e-4, abs=1e-5), (step, k)
    for net in ("actor", "forward_net", "backward_net", "forward_target_net", "backward_target_net"):
        for name, ref in subtree(g, f"paramN/{net}").items():
            assert np.abs(getattr(agent, net)[name].detach().numpy() - ref).max() < 5e-5, (net, name)


def test_schedule():
    assert O.schedule("0.2", 10) == 0.2
    assert O.schedule("linear(1,0.2,200)", 100) == pytest.approx(0.6)
    assert O.schedule("step_linear(1,0.5,100,0.1,100)", 150) == pytest.approx(0.3)
    with pytest.raises(NotImplementedError):
        O.schedule("cosine(1)", 0)


SMALL = O.Dims(obs_dim=11, action_dim=3, z_dim=10, goal_dim=11, hidden_dim=48, feature_dim=24, backward_hidden_dim=30)   # make_golden.CASES["small"]


def _branch_dims(case):
    """make_golden's "small" dims; debug=True puts z in goal space (z_dim = obs_dim = 11)."""
    import dataclasses
    return dataclasses.replace(SMALL, z_dim=11) if case == "debug" else SMALL


@pytest.mark.parametrize("case", ["nopre", "boltz", "debug"])
def test_oracle_only_branches_update(case):
    """preprocess=False, boltzmann=True and debug=True (identity backward map): the oracle's restatement of these fb_ddpg branches is
    pinned against the reference here (the CUDA step is checked against the same files in the -m gpu tests)."""
    g = load_golden(f"update_{case}")
    boltz = bool(g["cfg/boltzmann"])
    SMALL = _branch_dims(case)
    fwd_spec = O.forward_map_spec(SMALL, preprocess=case != "nopre")
    act_spec = O.boltzmann_actor_spec(SMALL) if boltz else O.actor_spec(SMALL, preprocess=case != "nopre")
    for net, spec in (("forward_net", fwd_spec), ("actor", act_spec), ("backward_net", [] if case == "debug" else O.backward_map_spec(SMALL))):
        ref = subtree(g, f"param0/{net}")
        assert [(n, tuple(s)) for n, s in spec] == [(k, v.shape) for k, v in ref.items()]
    t = {k: torch.from_numpy(v.copy()) for k, v in subtree(g, "in").items()}
    res = O.fb_loss_and_grads(
        params(g, "param0/forward_net"), params(g, "param0/backward_net"), params(g, "param0/forward_target_net"),
        params(g, "param0/backward_target_net"), params(g, "param0/actor"), t["obs"], t["action"], t["discount"],
        t["next_obs"], t["next_goal"], t["z"], t["noise_fb"], float(g["cfg/stddev"]), float(g["cfg/stddev_clip"]),
        float(g["cfg/ortho_coef"]), SMALL.z_dim, boltzmann=boltz)
    for k, v in subtree(g, "metric_fb").items():
        if k != "fb_opt_lr":
            assert res["metrics"][k] == pytest.approx(float(v), rel=2e-5, abs=2e-6), k
    for net, key in (("forward_net", "grads_forward"), ("backward_net", "grads_backward")):
        for name, ref in subtree(g, f"grad_fb/{net}").items():
            assert rel(res[key][name].numpy(), ref) < 1e-5, (net, name)
    res = O.actor_loss_and_grads(params(g, "param0/actor"), params(g, "param1/forward_net"), t["obs"], t["z"], t["noise_actor"],
                                 float(g["cfg/stddev"]), float(g["cfg/stddev_clip"]), boltzmann=boltz, temp=float(g["cfg/temp"]))
    m = subtree(g, "metric_actor")
    for k in ("actor_loss", "q", "actor_logprob"):
        assert float(res[k]) == pytest.approx(float(m[k]), rel=2e-5, abs=2e-6), k
    for name, ref in subtree(g, "grad_actor/actor").items():
        assert rel(res["grads_actor"][name].numpy(), ref) < 1e-5, name


@pytest.mark.parametrize("case", ["nopre", "boltz", "debug"])
def test_oracle_only_branches_trajectory(case):
    torch.set_num_threads(1)
    g = load_golden(f"trajectory_{case}")
    agent = O.OracleAgent(O.OracleConfig(dims=_branch_dims(case), batch_size=32, preprocess=case != "nopre", boltzmann=case == "boltz", temp=0.7,
                                         debug=case == "debug"))
    a, f, b = subtree(g, "param0/actor"), subtree(g, "param0/forward_net"), subtree(g, "param0/backward_net")
    agent.load_params(actor=a, forward_net=f, backward_net=b, forward_target_net=f, backward_target_net=b)
    buf = O.OracleReplay(4, 0.98, 0.99)
    for i in range(4):
        buf.add_episode({k: (v if v.ndim > 1 else v[:, None]) for k, v in subtree(g, f"ep{i}").items()})
    seed = int(g["seed"])
    torch.manual_seed(seed + 1)
    np.random.seed(seed + 1)
    for step in range(int(g["steps"])):
        m = agent.update(buf, step)
        ref = subtree(g, f"step{step}")
        assert set(ref) == set(m)
        for k, v in ref.items():
            assert m[k] == pytest.approx(float(v), rel=2e-4, abs=1e-5), (step, k)
    for net in ("actor", "forward_net", "backward_net", "forward_target_net", "backward_target_net"):
        for name, ref in subtree(g, f"paramN/{net}").items():
            assert np.abs(getattr(agent, net)[name].detach().numpy() - ref).max() < 5e-5, (net, name)
