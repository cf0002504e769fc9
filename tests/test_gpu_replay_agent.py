"""GPU tests of the drop-in boundary: ReplayBuffer (HBM storage + gather kernel) bit-exact against the reference's
sample() fixtures, and FBDDPGAgent.update() against the reference's golden trajectories."""
import dataclasses
import io

import numpy as np
import pytest
import torch

from conftest import load_golden, subtree
from gpu_common import rel

pytestmark = pytest.mark.gpu


@dataclasses.dataclass
class FakeTimeStep:
    """Field-compatible stand-in for url_benchmark.dmc.ExtendedTimeStep / ExtendedGoalTimeStep (dmc.py:35-73)."""
    step_type: int
    reward: float
    discount: float
    observation: np.ndarray
    action: np.ndarray
    physics: np.ndarray

    def __getitem__(self, k):
        return getattr(self, k)

    def last(self):
        return self.step_type == 2


@dataclasses.dataclass
class FakeGoalTimeStep(FakeTimeStep):
    goal: np.ndarray = None


def _fill(buf, g, via_add):
    mel = int(g["max_episode_length"])
    for i in range(int(g["n_episodes"])):
        ep = subtree(g, f"ep{i}")
        if via_add:
            n = len(ep["reward"])
            for t in range(n):
                st = 0 if t == 0 else (2 if t == n - 1 else 1)
                kw = dict(step_type=st, reward=float(ep["reward"][t]), discount=float(ep["discount"][t]), observation=ep["observation"][t],
                          action=ep["action"][t], physics=ep["physics"][t])
                ts = FakeGoalTimeStep(goal=ep["goal"][t], **kw) if "goal" in ep else FakeTimeStep(**kw)
                buf.add(ts, {"z": ep["z"][t]})
        else:
            buf.add_episode(ep)
    return mel


@pytest.mark.parametrize("case", ["fixed", "fixed_goal_full", "ragged", "nofuture"])
@pytest.mark.parametrize("via_add", [False, True])
def test_replay_sample_bit_exact_vs_reference(case, via_add):
    from controllable_agent_b200 import ReplayBuffer
    g = load_golden(f"replay_{case}")
    mel = int(g["max_episode_length"])
    buf = ReplayBuffer(int(g["max_episodes"]), 0.98, float(g["future"]), max_episode_length=mel if mel > 0 else None)
    _fill(buf, g, via_add)
    assert len(buf) == int(g["len"]) and buf._full == bool(g["full"]) and buf._is_fixed_episode_length == bool(g["fixed"])
    assert buf.avg_episode_length == int(g["avg_episode_length"])
    np.testing.assert_array_equal(buf._episodes_length, g["episodes_length"])
    for draw in range(3):
        np.random.seed(int(g["seed"]) + 100 + draw)
        batch = buf.sample(16)
        ref = subtree(g, f"draw{draw}")
        for field in ("obs", "action", "reward", "discount", "next_obs", "goal", "next_goal", "future_obs", "future_goal"):
            got = getattr(batch, field)
            if field in ref:
                assert got.is_cuda
                np.testing.assert_array_equal(got.cpu().numpy(), ref[field], err_msg=field)   # bit-exact
            else:
                assert got is None, field
        np.testing.assert_array_equal(batch.meta["z"].cpu().numpy(), ref["meta/z"])
        assert batch.to("cuda").obs.data_ptr() == batch.obs.data_ptr()   # already on the device: .to() is a no-op


def test_replay_pickle_roundtrip_and_storage_view():
    from controllable_agent_b200 import ReplayBuffer
    g = load_golden("replay_fixed_goal_full")
    buf = ReplayBuffer(int(g["max_episodes"]), 0.98, float(g["future"]))
    _fill(buf, g, False)
    st = buf._storage
    assert set(st) >= {"observation", "action", "reward", "discount", "goal", "z", "physics"}
    assert st["observation"].shape[0] == int(g["max_episodes"])
    f = io.BytesIO()
    torch.save(buf, f, pickle_protocol=4)
    f.seek(0)
    buf2 = torch.load(f, weights_only=False)
    assert len(buf2) == len(buf) and buf2._idx == buf._idx and buf2._full == buf._full
    for draw in range(2):
        np.random.seed(99 + draw)
        a = buf.sample(16)
        np.random.seed(99 + draw)
        b = buf2.sample(16)
        for field in ("obs", "action", "reward", "discount", "next_obs", "goal", "next_goal", "future_obs", "future_goal"):
            assert torch.equal(getattr(a, field), getattr(b, field)), field


def test_pack_episode_kernel_matches_host_packing():
    import ctypes as C
    from controllable_agent_b200 import ReplayBuffer, _lib as L
    g = load_golden("replay_fixed_goal_full")
    buf = ReplayBuffer(int(g["max_episodes"]), 0.98, float(g["future"]))
    _fill(buf, g, False)
    ep = {k: (v if v.ndim > 1 else v[:, None]) for k, v in subtree(g, "ep3").items()}
    view = buf.view()
    rows = torch.zeros_like(buf._rows)
    d = {k: torch.as_tensor(v).cuda().contiguous() for k, v in ep.items()}
    lib = L.load()
    L.check(lib.fb_replay_pack_episode(C.byref(view), rows.data_ptr(), 3, len(ep["reward"]), ep["observation"].shape[1], ep["action"].shape[1],
                                       d["observation"].data_ptr(), d["action"].data_ptr(), d["reward"].data_ptr(), d["discount"].data_ptr(),
                                       d["goal"].data_ptr(), d["z"].data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(rows[3], buf._rows[3])


SMALL_CASES = ("small", "future", "qloss", "nonorm", "randw", "randw_nonorm", "trunk", "nopre", "boltz", "debug")   # fixtures generated from make_golden.CASES["small"]


def _agent_for(g, case, **kw):
    from controllable_agent_b200 import FBDDPGAgent
    a, f, b = subtree(g, "param0/actor"), subtree(g, "param0/forward_net"), subtree(g, "param0/backward_net")
    from gpu_common import dims_from_params
    d = dims_from_params(f, b, a)
    hidden, obs_dim, oa = d.hidden_dim, d.obs_dim, d.obs_dim + d.action_dim
    agent = FBDDPGAgent(obs_type="states", obs_shape=(obs_dim,), action_shape=(oa - obs_dim,), device="cuda", num_expl_steps=0,
                        update_encoder=True, goal_space=None if case in SMALL_CASES else "simplified_walker", use_tb=True, use_wandb=False,
                        use_hiplog=False, hidden_dim=hidden, feature_dim=d.feature_dim,
                        backward_hidden_dim=d.backward_hidden_dim, z_dim=f["F1.2.weight"].shape[0],
                        batch_size=32 if case in SMALL_CASES else 64, update_every_steps=1,
                        future_ratio=0.4 if case.startswith("future") else (0.3 if case == "nonorm" else 0.0), **kw)
    for net, src in ((agent.actor, a), (agent.forward_net, f), (agent.backward_net, b), (agent.forward_target_net, f),
                     (agent.backward_target_net, b)):
        for (name, p) in net.named_parameters():
            p.data.copy_(torch.as_tensor(src[name]))
    return agent


@pytest.mark.parametrize("case", ["small", "goal", "future", "future_goal", "qloss", "nonorm", "randw", "randw_nonorm", "trunk", "nopre", "boltz", "debug"])
@pytest.mark.parametrize("foreign_replay", [False, True, "fused"])   # "fused": HBM replay, the MLP stacks as fused persistent kernels
def test_agent_update_walks_reference_trajectory(case, foreign_replay):
    fuse = foreign_replay == "fused"
    foreign_replay = foreign_replay is True
    """agent.update(replay, step) x3 with the reference's RNG streams (rng_mode=reference, torch draws on the CPU generator
    as in the CPU-generated fixture).  Step 0 is gated at 1e-3; later steps inherit Adam's sign(g) amplification of
    ulp-level gradient differences (SURVEY.md 7.3) and are gated loosely."""
    from controllable_agent_b200 import ReplayBuffer
    g = load_golden(f"trajectory_{case}")
    extra = dict(q_loss=True, q_loss_coef=0.5, additional_metric=True) if case == "qloss" else {}   # fb_ddpg.py:330-341,403-404
    if case == "nonorm":   # norm_z = False with hindsight rows (future_ratio = 0.3)
        extra = dict(norm_z=False)
    if case.startswith("randw"):   # rand_weight = True (fb_ddpg.py:475-482), with and without the re-projection
        extra = dict(rand_weight=True, norm_z=case == "randw")
    if case == "trunk":   # add_trunk = True (fb_modules.py:96-100,169-173)
        extra = dict(add_trunk=True)
    if case == "nopre":   # preprocess = False (fb_modules.py:102-104,175-177)
        extra = dict(preprocess=False)
    if case == "boltz":   # boltzmann = True (fb_modules.py:129-151), temp as make_golden generated it
        extra = dict(boltzmann=True, temp=float(g["cfg/temp"]) if "cfg/temp" in g else 0.7)
    if case == "debug":   # debug = True: identity backward map (fb_ddpg.py:128-130)
        extra = dict(debug=True)
    agent = _agent_for(g, case, rng_mode="reference", fuse_stacks=fuse, **extra)
    agent.draw_device = "cpu"
    eps = [subtree(g, f"ep{i}") for i in range(4)]
    if foreign_replay:   # a host-memory replay object with the reference's sample() contract: explicit-batch path
        from oracle import fb_oracle as O
        from controllable_agent_b200 import EpisodeBatch
        ob = O.OracleReplay(4, 0.98, 0.99)
        for ep in eps:
            ob.add_episode({k: (v if v.ndim > 1 else v[:, None]) for k, v in ep.items()})

        class Host:
            def sample(self, n):
                s = ob.sample(n)
                return EpisodeBatch(obs=s["obs"], action=s["action"], reward=s["reward"], discount=s["discount"], next_obs=s["next_obs"],
                                    goal=s["goal"], next_goal=s["next_goal"], future_obs=s["future_obs"], future_goal=s["future_goal"],
                                    meta=s["meta"])
        buf = Host()
    else:
        buf = ReplayBuffer(4, 0.98, 0.99)
        for ep in eps:
            buf.add_episode(ep)
    seed = int(g["seed"])
    torch.manual_seed(seed + 1)
    np.random.seed(seed + 1)
    for step in range(int(g["steps"])):
        m = agent.update(buf, step)
        ref = subtree(g, f"step{step}")
        assert set(ref) == set(m), (sorted(ref), sorted(m))
        tol = 1e-3 if step == 0 else 2e-2
        for k, v in ref.items():
            if k == "q1_success":   # a count of rows / batch: rows with Q1 ~ Q2 may fall on either side
                assert abs(m[k] - float(v)) <= 2.0 / agent.cfg.batch_size + 1e-6, (step, k, m[k], float(v))
                continue
            assert m[k] == pytest.approx(float(v), rel=tol, abs=2e-4), (step, k)
    for net in ("actor", "forward_net", "backward_net", "forward_target_net", "backward_target_net"):
        for (name, p) in getattr(agent, net).named_parameters():
            assert np.abs(p.detach().cpu().numpy() - g[f"paramN/{net}/{name}"]).max() < 6.5e-4, (net, name)   # <= 2 lr x 3 steps


def test_agent_device_rng_step_statistics_and_api():
    """rng_mode=device: the whole update is one graph launch; check the Philox draws and the public surface."""
    from controllable_agent_b200 import FBDDPGAgent, ReplayBuffer, _lib as L
    torch.manual_seed(3)
    np.random.seed(3)
    B, O_, A_, Z = 256, 24, 6, 50
    agent = FBDDPGAgent(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True,
                        goal_space=None, use_tb=True, use_wandb=True, use_hiplog=False, batch_size=B, update_every_steps=2)
    assert sum(p.numel() for p in agent.forward_net.parameters()) == 3363940
    assert sum(p.numel() for p in agent.backward_net.parameters()) == 317754
    assert sum(p.numel() for p in agent.actor.parameters()) == 2211846
    rs = np.random.RandomState(0)
    buf = ReplayBuffer(12, 0.98, 0.99)
    for _ in range(9):
        n = 30
        buf.add_episode({"observation": rs.standard_normal((n + 1, O_)), "action": rs.uniform(-1, 1, (n + 1, A_)),
                         "reward": rs.uniform(0, 1, (n + 1,)), "discount": np.ones(n + 1), "physics": np.zeros((n + 1, 2))})
    assert agent.update(buf, 1) == {}          # gated by update_every_steps (fb_ddpg.py:430-431)
    p0 = agent.forward_net.F1[0].weight.detach().clone()
    t0 = agent.forward_target_net.F1[0].weight.detach().clone()
    seen_z = []
    for step in range(0, 8, 2):
        m = agent.update(buf, step)
        assert set(m) == set(L.METRIC_KEYS) | {"fb_opt_lr"}
        assert all(np.isfinite(v) for v in m.values()), m
        assert m["z_norm"] == pytest.approx(np.sqrt(Z), rel=1e-4) and m["B_norm"] == pytest.approx(np.sqrt(Z), rel=1e-4)
        assert m["orth_loss_diag"] == pytest.approx(-2 * Z, rel=1e-4)
        e = agent.engine
        z = e.view("z").clone()
        seen_z.append(z)
        obs, nobs = e.view("actor_in_o")[B:], e.view("actor_in_o")[:B]
        # every gathered (obs, next_obs) pair must be consecutive rows of one stored episode
        rows = buf._rows[:9, :, :O_].reshape(-1, O_)
        idx = torch.cdist(obs.double(), rows.double()).argmin(1)
        assert torch.equal(rows[idx], obs)       # gathered rows are bit-exact copies of stored rows
        assert torch.equal(rows[idx + 1], nobs) and bool(((idx % 31) < 30).all())
    assert not torch.equal(seen_z[0], seen_z[1])       # fresh draws each step
    assert not torch.equal(agent.forward_net.F1[0].weight, p0) and not torch.equal(agent.forward_target_net.F1[0].weight, t0)
    assert agent.engine.get_adam_steps() == (4, 4)
    # public helpers
    meta = agent.init_meta()
    assert meta["z"].shape == (Z,)
    with torch.no_grad():
        act = agent.act(rs.standard_normal(O_).astype(np.float32), meta, 0, eval_mode=True)
    assert act.shape == (A_,) and np.all(np.abs(act) <= 1)
    gm = agent.get_goal_meta(rs.standard_normal(O_).astype(np.float32))
    assert gm["z"].shape == (Z,) and np.linalg.norm(gm["z"]) == pytest.approx(np.sqrt(Z), rel=1e-4)
    im = agent.infer_meta(buf)
    assert im["z"].shape == (Z,)
    # pickling (pretrain.py:437-449) and init_from (fb_ddpg.py:166-175)
    f = io.BytesIO()
    torch.save({"agent": agent}, f, pickle_protocol=4)
    f.seek(0)
    other = torch.load(f, weights_only=False)["agent"]
    assert torch.equal(other.engine.param_fb, agent.engine.param_fb) and torch.equal(other.engine.m_actor, agent.engine.m_actor)
    assert other.engine.get_adam_steps() == (4, 4)
    fresh = FBDDPGAgent(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True,
                        goal_space=None, use_tb=False, use_wandb=False, use_hiplog=False, batch_size=B, update_every_steps=2)
    fresh.init_from(agent)
    assert torch.equal(fresh.engine.param_actor, agent.engine.param_actor) and torch.equal(fresh.engine.v_fb, agent.engine.v_fb)
    assert fresh.engine.get_adam_steps() == (4, 4)
    assert fresh.update(buf, 0) == {}          # metrics off -> empty dict, step still runs
    assert fresh.engine.get_adam_steps() == (5, 5)


def test_hindsight_rows_with_device_rng():
    """future_ratio > 0 with rng_mode=device: the hindsight mask is drawn by the Philox kernel, masked rows take
    z = backward_net(future_obs) (rows [B, 2B) of the batched backward_net forward), the others the mixed / random z."""
    from controllable_agent_b200 import FBDDPGAgent, ReplayBuffer
    torch.manual_seed(2)
    B, O_, A_, Z = 256, 24, 6, 50
    agent = FBDDPGAgent(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True,
                        goal_space=None, use_tb=True, use_wandb=False, use_hiplog=False, batch_size=B, update_every_steps=1,
                        hidden_dim=256, feature_dim=128, backward_hidden_dim=134, future_ratio=0.3, mix_ratio=0.5)
    rs = np.random.RandomState(0)
    buf = ReplayBuffer(8, 0.98, 0.9)
    for _ in range(8):
        n = 40
        buf.add_episode({"observation": rs.standard_normal((n + 1, O_)), "action": rs.uniform(-1, 1, (n + 1, A_)),
                         "reward": rs.uniform(0, 1, (n + 1,)), "discount": np.ones(n + 1)})
    fracs = []
    for step in range(4):
        m = agent.update(buf, step)
        assert all(np.isfinite(v) for v in m.values()), m
        e = agent.engine
        z, bm, zr = e.view("z"), e.view("B_mix"), e.view("z_rand")
        hind = (z == bm[B:]).all(dim=1)                      # hindsight rows: copied verbatim
        mixed = ~hind & ~(z == zr).all(dim=1)                # mixing rows: renormalised backward_net(obs[perm])
        fracs.append((float(hind.float().mean()), float(mixed.float().mean())))
        assert torch.allclose(z.norm(dim=1), torch.full((B,), float(np.sqrt(Z)), device=z.device), rtol=1e-4)
        # the hindsight input is a stored observation of the same episode at or after the sampled step
        fin = e.view("mix_input")[B:]
        rows = buf._rows[:8, :, :O_].reshape(-1, O_)
        idx = torch.cdist(fin.double(), rows.double()).argmin(1)
        assert torch.equal(rows[idx], fin)                   # bit-exact copies of stored rows
    h, mx = np.mean([f[0] for f in fracs]), np.mean([f[1] for f in fracs])
    assert 0.2 < h < 0.4 and 0.25 < mx < 0.45, fracs    # 0.3 and 0.5 * 0.7


def test_agent_host_replay_with_device_rng():
    """A host-memory replay (the reference's ReplayBuffer contract: sample() -> numpy EpisodeBatch) feeding the device step
    with rng_mode=device: the batch crosses as one packed upload (fb_upload_batch), FB_PHASE_SAMPLE draws z / noise / perm /
    mix mask on the device without a bound replay (FB_RUN_HOST_BATCH), with and without a goal space."""
    from controllable_agent_b200 import EpisodeBatch, FBDDPGAgent, _lib as L
    for goal_space, G in ((None, 0), ("simplified_walker", 3)):
        torch.manual_seed(5)
        B, O_, A_, Z = 128, 24, 6, 50
        agent = FBDDPGAgent(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True,
                            goal_space=goal_space, use_tb=True, use_wandb=True, use_hiplog=False, batch_size=B, update_every_steps=1,
                            hidden_dim=256, feature_dim=128, backward_hidden_dim=134)
        rs = np.random.RandomState(1)

        class Host:
            samples = []

            def sample(self, n):
                f = lambda *shape: rs.standard_normal(shape).astype(np.float32)  # noqa: E731
                Host.samples.append(EpisodeBatch(obs=f(n, O_), action=rs.uniform(-1, 1, (n, A_)).astype(np.float32), reward=f(n, 1),
                                                 discount=np.full((n, 1), 0.98, np.float32), next_obs=f(n, O_),
                                                 goal=f(n, G) if G else None, next_goal=f(n, G) if G else None))
                return Host.samples[-1]

        zs, host = [], Host()
        for step in range(6):
            agent.cfg.prefetch_host_batch = step >= 3   # steps 3..5: the next batch is sampled + uploaded while the step runs
            m = agent.update(host, step)
            assert set(m) == set(L.METRIC_KEYS) | {"fb_opt_lr"} and all(np.isfinite(v) for v in m.values()), m
            assert m["z_norm"] == pytest.approx(np.sqrt(Z), rel=1e-4)
            assert len(Host.samples) == step + 1 + (1 if step >= 3 else 0)
            e, b = agent.engine, Host.samples[step]
            assert torch.equal(e.view("actor_in_o")[B:].cpu(), torch.from_numpy(b.obs))
            assert torch.equal(e.view("actor_in_o")[:B].cpu(), torch.from_numpy(b.next_obs))
            assert torch.equal(e.view("in_oa")[:, O_:].cpu(), torch.from_numpy(b.action))
            assert torch.equal(e.view("discount").cpu(), torch.from_numpy(b.discount))
            assert torch.equal(e.view("next_goal").cpu(), torch.from_numpy(b.next_goal if G else b.next_obs))
            zs.append(e.view("z").clone())
        assert not torch.equal(zs[0], zs[1]) and not torch.equal(zs[1], zs[2])
        assert agent.engine.get_adam_steps() == (6, 6)
        assert agent.last_update_launches == agent.engine.launch_count(L.PHASE_ALL | L.RUN_HOST_BATCH)
        assert agent.engine.launch_count(L.PHASE_ALL | L.RUN_HOST_BATCH) == agent.engine.launch_count(L.PHASE_ALL) - 1


@pytest.mark.parametrize("rng_mode", ["device", "reference"])
@pytest.mark.parametrize("goal_space,G,future", [(None, 0, 1.0), ("simplified_walker", 3, 0.9)])
def test_agent_samples_reference_layout_host_replay_natively(goal_space, G, future, rng_mode):
    """A host buffer with the reference ReplayBuffer's attribute layout (in_memory_replay_buffer.py:66-88) is sampled without its
    Python sample(): index draws in the reference's order on the numpy generator, row gathers by fb_host_gather_rows into the pinned
    block.  The rows that land on the device are exactly what its own sample() would have returned for the same generator state."""
    from controllable_agent_b200 import EpisodeBatch, FBDDPGAgent, _lib as L
    from test_cpu_boundary import _RefLayoutHostReplay
    B, O_, A_, Z = 64, 24, 6, 50
    rep = _RefLayoutHostReplay(np.random.default_rng(2), 9, 30, O_, A_, G, future)

    def ref_sample(n):   # what in_memory_replay_buffer.py:139-190 returns for these attributes
        from controllable_agent_b200.replay import draw_sample_indices
        # rng_mode="reference": the reference's own numpy calls (its generator stream); "device": the cheaper equivalent draws
        ep, st, fu = draw_sample_indices(rep, n, exact_stream=rng_mode == "reference")
        S = rep._storage
        return EpisodeBatch(obs=S["observation"][ep, st - 1], action=S["action"][ep, st], reward=S["reward"][ep, st],
                            discount=rep._discount * S["discount"][ep, st], next_obs=S["observation"][ep, st],
                            goal=S["goal"][ep, st - 1] if G else None, next_goal=S["goal"][ep, st] if G else None,
                            future_obs=S["observation"][ep, fu - 1] if fu is not None else None,
                            future_goal=S["goal"][ep, fu - 1] if (fu is not None and G) else None)
    rep.sample = ref_sample
    torch.manual_seed(5)
    agent = FBDDPGAgent(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True,
                        goal_space=goal_space, use_tb=True, use_wandb=True, use_hiplog=False, batch_size=B, update_every_steps=1,
                        hidden_dim=256, feature_dim=128, backward_hidden_dim=134, future_ratio=0.3 if future < 1 else 0.0, rng_mode=rng_mode)
    agent.draw_device = "cpu"
    for step in range(4):
        agent.native_host_sampling = step % 2 == 0   # alternate: library gather / the object's own sample()
        agent.cfg.prefetch_host_batch = False
        state = np.random.get_state()
        m = agent.update(rep, step)
        after = np.random.get_state()
        np.random.set_state(state)
        want = ref_sample(B)                          # the batch the reference would have drawn from the same generator state
        np.random.set_state(after)
        e = agent.engine
        assert torch.equal(e.view("actor_in_o")[B:].cpu(), torch.from_numpy(want.obs))
        assert torch.equal(e.view("actor_in_o")[:B].cpu(), torch.from_numpy(want.next_obs))
        assert torch.equal(e.view("in_oa")[:, O_:].cpu(), torch.from_numpy(want.action))
        assert torch.equal(e.view("discount").cpu(), torch.from_numpy(want.discount.astype(np.float32)))
        assert torch.equal(e.view("next_goal").cpu(), torch.from_numpy(want.next_goal if G else want.next_obs))
        assert all(np.isfinite(v) for v in m.values()) and m["z_norm"] == pytest.approx(np.sqrt(Z), rel=1e-4)
    assert agent.engine.get_adam_steps() == (4, 4)


@pytest.mark.parametrize("goal_space,G,add_trunk", [(None, 24, False), ("simplified_walker", 3, False), (None, 24, True)])
def test_inference_plans_match_the_module_forward(goal_space, G, add_trunk):
    """act / get_goal_meta / compute_z_correl / infer_meta_from_obs_and_rewards run through the library's inference plans
    (FB_PHASE_INFER_*); checked against the same networks evaluated by the parameter-view nn.Modules (modules.py mirrors
    fb_modules.py; pinned against the oracle in test_cpu_boundary.py), at the reference's default widths."""
    import math
    import torch.nn.functional as F
    from controllable_agent_b200 import FBDDPGAgent
    torch.manual_seed(9)
    O_, A_, Z = 24, 6, 50
    agent = FBDDPGAgent(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True,
                        goal_space=goal_space, use_tb=False, use_wandb=False, use_hiplog=False, batch_size=256, num_inference_steps=600,
                        add_trunk=add_trunk)
    for p in list(agent.actor.parameters()) + list(agent.backward_net.parameters()):   # biases / LN affine away from their init
        if p.dim() == 1:
            p.data.add_(0.1 * torch.randn_like(p))
    rs = np.random.RandomState(4)
    for trial in range(3):
        obs, goal = rs.standard_normal(O_).astype(np.float32), rs.standard_normal(G).astype(np.float32)
        meta = agent.init_meta()
        with torch.no_grad():
            o, z = torch.as_tensor(obs).cuda()[None], torch.as_tensor(meta["z"]).cuda()[None]
            mu_ref = agent.actor(o, z, 0.2).mean[0].cpu().numpy()
            b_ref = agent.backward_net(torch.as_tensor(goal).cuda()[None])
            zg_ref = (math.sqrt(Z) * F.normalize(b_ref, dim=1))[0].cpu().numpy()
            corr_ref = torch.matmul(F.normalize(b_ref, 1), F.normalize(z, 1).T).item()
        assert np.abs(agent.act(obs, meta, 0, eval_mode=True) - mu_ref).max() < 2e-5
        a = agent.act(obs, meta, 0, eval_mode=False)
        assert a.shape == (A_,) and np.all(np.abs(a) <= 1) and np.abs(a - mu_ref).max() < 0.2 * 6   # mu + N(0, 0.2) noise
        assert np.abs(agent.get_goal_meta(goal)["z"] - zg_ref).max() < 2e-4

        class TS:
            observation = goal if goal_space is None else obs
        TS.goal = goal
        assert agent.compute_z_correl(TS, meta) == pytest.approx(corr_ref, rel=1e-3, abs=1e-6)
    N = 600   # two full chunks of 256 rows and a padded one
    xs, rw = torch.randn(N, G, device="cuda"), torch.rand(N, 1, device="cuda")
    with torch.no_grad():
        zr = torch.matmul(rw.T, agent.backward_net(xs)) / N
        zr = (math.sqrt(Z) * F.normalize(zr, dim=1))[0].cpu().numpy()
    got = agent.infer_meta_from_obs_and_rewards(xs, rw)["z"]
    assert got.shape == (Z,) and np.abs(got - zr).max() < 2e-4


def test_device_rng_with_norm_z_off():
    """rng_mode=device, norm_z=False: z = sqrt(Z) * U[0,1) (x) direction (fb_ddpg.py:230-231) -> E|z|^2 = Z/3; backward_net outputs and
    the mixed rows stay un-projected."""
    from controllable_agent_b200 import FBDDPGAgent, ReplayBuffer
    torch.manual_seed(5)
    O_, A_, Z, Bsz = 12, 4, 24, 512
    kw = dict(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True, goal_space=None,
              use_tb=True, use_wandb=False, use_hiplog=False, hidden_dim=64, feature_dim=32, backward_hidden_dim=38, z_dim=Z, batch_size=Bsz,
              update_every_steps=1, norm_z=False, lr=1e-12)   # lr ~ 0: the modules after the step still equal the step's networks
    rs = np.random.RandomState(0)
    buf = ReplayBuffer(8, 0.98, 0.99)
    for _ in range(8):
        buf.add_episode({"observation": rs.standard_normal((101, O_)), "action": rs.uniform(-1, 1, (101, A_)), "reward": rs.uniform(0, 1, 101),
                         "discount": np.ones(101)})
    agent = FBDDPGAgent(**kw, mix_ratio=0.0)
    m = agent.update(buf, 0)
    z = agent.engine.view("z").cpu().double()
    assert float((z ** 2).sum(1).mean()) == pytest.approx(Z / 3.0, rel=0.06)
    assert m["z_norm"] == pytest.approx(float(z.norm(dim=1).mean()), rel=1e-4)
    assert float(z.abs().max()) < np.sqrt(Z)
    Bm = agent.engine.view("B").cpu()
    with torch.no_grad():
        raw = agent.backward_net(agent.engine.view("next_goal").clone()).cpu()   # module forward with norm_z = False: no projection
    assert float((Bm - raw).abs().max()) < 1e-4 * max(1.0, float(raw.abs().max()))
    assert abs(m["B_norm"] - np.sqrt(Z)) > 1e-2 and abs(m["orth_loss_diag"] + 2 * Z) > 1e-2
    mixed = FBDDPGAgent(**kw, mix_ratio=1.0)
    mixed.update(buf, 0)
    zb = mixed.engine.view("z").cpu()
    assert float((zb.norm(dim=1) - np.sqrt(Z)).abs().min()) > 1e-3      # mixed rows are raw backward_net outputs, not re-projected
    assert float((zb - mixed.engine.view("B_mix").cpu()[:Bsz]).abs().max()) == 0.0


def test_rand_weight_with_device_rng():
    """rng_mode=device, rand_weight=True, norm_z=False: every row of z is either its random draw or
    (u_s / |W_s|) sum_t W[s,t] B_mix[t] recomputed here in float64 from the library's own weight block (fb_ddpg.py:475-482)."""
    from controllable_agent_b200 import FBDDPGAgent, ReplayBuffer
    torch.manual_seed(6)
    O_, A_, Z, Bsz = 12, 4, 24, 200
    kw = dict(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True, goal_space=None,
              use_tb=True, use_wandb=False, use_hiplog=False, hidden_dim=64, feature_dim=32, backward_hidden_dim=38, z_dim=Z, batch_size=Bsz,
              update_every_steps=1, norm_z=False, rand_weight=True, mix_ratio=0.5)
    rs = np.random.RandomState(1)
    buf = ReplayBuffer(8, 0.98, 0.99)
    for _ in range(8):
        buf.add_episode({"observation": rs.standard_normal((101, O_)), "action": rs.uniform(-1, 1, (101, A_)), "reward": rs.uniform(0, 1, 101),
                         "discount": np.ones(101)})
    agent = FBDDPGAgent(**kw)
    seen = []
    for step in range(2):
        m = agent.update(buf, step)
        assert all(np.isfinite(v) for v in m.values()), m
        e = agent.engine
        W, u = e.view("mix_w").cpu().double(), e.view("mix_u").cpu().double()[:, 0]
        assert 0.0 <= float(W.min()) and float(W.max()) < 1.0 and float(W.mean()) == pytest.approx(0.5, abs=0.01)
        assert 0.0 <= float(u.min()) and float(u.max()) < 1.0
        cand = (u / W.norm(dim=1))[:, None] * (W @ e.view("B_mix").cpu().double()[:Bsz])
        z, zr = e.view("z").cpu().double(), e.view("z_rand").cpu().double()
        d_mix, d_rand = (z - cand).abs().max(dim=1).values, (z - zr).abs().max(dim=1).values
        assert float(torch.minimum(d_mix, d_rand).max()) < 1e-5 * max(1.0, float(cand.abs().max()))
        n_mixed = int((d_mix < d_rand).sum())
        assert 0.35 * Bsz < n_mixed < 0.65 * Bsz, n_mixed
        seen.append(W.clone())
    assert not torch.equal(seen[0], seen[1])   # fresh weights every step


def test_unsupported_branches_raise():
    from controllable_agent_b200 import FBDDPGAgent
    base = dict(obs_type="states", obs_shape=(24,), action_shape=(6,), device="cuda", num_expl_steps=0, update_encoder=True, goal_space=None,
                use_tb=False, use_wandb=False, use_hiplog=False)
    for kw in (dict(obs_type="pixels"),):
        with pytest.raises(NotImplementedError):
            FBDDPGAgent(**{**base, **kw})
    with pytest.raises(ValueError):   # debug=True needs z_dim == goal_dim
        FBDDPGAgent(**{**base, "debug": True})
    with pytest.raises(RuntimeError):
        FBDDPGAgent(**{**base, "device": "cpu"})
    with pytest.raises(ValueError):
        FBDDPGAgent(**{**base, "future_ratio": 1.5})


class _RefLayoutReplay:
    """An object with the attribute layout of url_benchmark.in_memory_replay_buffer.ReplayBuffer (:66-88), built from a
    reference-generated fixture: `_storage` name -> [max_episodes, T+1, dim] numpy, as its pickles hold it."""

    def __init__(self, g, load_filled):
        n, E = int(g["n_episodes"]), int(g["max_episodes"])
        eps = [subtree(g, f"ep{i}") for i in range(n)]
        self._max_episodes, self._discount, self._future = E, 0.98, float(g["future"])
        self._max_episode_length = None
        self._storage = {}
        for name, v in eps[0].items():
            v = v if v.ndim > 1 else v[:, None]
            self._storage[name] = np.zeros((E,) + v.shape, np.float32)
        for i in range(n):   # the ring: episode i lands in slot i % E, as add() / load() write it
            for name, v in eps[i].items():
                self._storage[name][i % E] = v if v.ndim > 1 else v[:, None]
        self._idx, self._full = int(g["idx"]) if "idx" in g else n % E, bool(g["full"])
        # load()-filled buffers leave the lengths at 0 (in_memory_replay_buffer.py:192-208)
        self._episodes_length = np.zeros(E, np.int32) if load_filled else np.asarray(g["episodes_length"], np.int32)

    def __len__(self):
        return self._max_episodes if self._full else self._idx


@pytest.mark.parametrize("load_filled", [False, True])
def test_from_reference_adopts_a_reference_layout_buffer(load_filled):
    """ReplayBuffer.from_reference on a reference-layout object (an ExORL-style pickle's content): storage lands in HBM, bookkeeping
    carries over, sample() returns the reference's own draws bit-exactly (no test)."""
    from controllable_agent_b200 import ReplayBuffer
    g = load_golden("replay_fixed_goal_full")
    other = _RefLayoutReplay(g, load_filled)
    buf = ReplayBuffer.from_reference(other, device="cuda")
    assert len(buf) == int(g["len"]) and buf._full == bool(g["full"]) and buf._idx == other._idx
    assert buf._discount == 0.98 and buf._future == float(g["future"]) and buf._is_fixed_episode_length
    np.testing.assert_array_equal(buf._episodes_length[:len(buf)], np.asarray(g["episodes_length"])[:len(buf)])
    for draw in range(3):
        np.random.seed(int(g["seed"]) + 100 + draw)
        batch = buf.sample(16)
        ref = subtree(g, f"draw{draw}")
        for field in ("obs", "action", "reward", "discount", "next_obs", "goal", "next_goal", "future_obs", "future_goal"):
            if field in ref:
                np.testing.assert_array_equal(getattr(batch, field).cpu().numpy(), ref[field], err_msg=field)
    st1 = buf._storage
    assert buf._storage is st1, "the host view is cached until the device rows change"
    np.testing.assert_array_equal(st1["observation"], other._storage["observation"])
    buf.add_episode(subtree(g, "ep0"))
    assert buf._storage is not st1


class _StubEnv:
    """physics.reset_context / set_state / task.get_reward: what relabel_episode (in_memory_replay_buffer.py:40-55) calls."""

    class _Physics:
        state = None

        def reset_context(self):
            import contextlib
            return contextlib.nullcontext()

        def set_state(self, s):
            self.state = np.asarray(s)

    class _Task:
        @staticmethod
        def get_reward(physics):
            return float(physics.state.sum()) * 0.5 + 1.0

    def __init__(self):
        self.physics, self.task = self._Physics(), self._Task()


def test_load_with_relabel_recomputes_rewards_and_goals(tmp_path):
    """ReplayBuffer.load(env, dir) at its DEFAULT relabel=True (the import it relied on does not exist in the reference)."""
    from controllable_agent_b200 import ReplayBuffer
    rs = np.random.RandomState(0)
    eps = []
    for i in range(3):
        ep = {"observation": rs.randn(6, 5).astype(np.float32), "action": rs.randn(6, 2).astype(np.float32),
              "reward": np.zeros((6, 1), np.float32), "discount": np.ones((6, 1), np.float32), "physics": rs.randn(6, 4).astype(np.float32)}
        np.savez(tmp_path / f"ep{i:02d}.npz", **ep)
        eps.append(ep)
    env = _StubEnv()
    buf = ReplayBuffer(3, 0.98, 1.0)
    buf.load(env, tmp_path, goal_func=lambda e: e.physics.state[:2] * 2.0)
    assert len(buf) == 3 and buf._full
    st = buf._storage
    for i, ep in enumerate(eps):
        np.testing.assert_allclose(st["reward"][i, :, 0], ep["physics"].sum(1) * 0.5 + 1.0, rtol=1e-6)
        np.testing.assert_allclose(st["goal"][i], ep["physics"][:, :2] * 2.0, rtol=1e-6)
    np.random.seed(5)
    batch = buf.sample(8)
    assert batch.goal is not None and batch.reward.shape == (8, 1)

    class _Reward:
        @staticmethod
        def from_physics(p):
            return float(p[0])
    buf.relabel(_Reward())
    np.testing.assert_allclose(buf._storage["reward"][:, :, 0], np.stack([ep["physics"][:, 0] for ep in eps]), rtol=1e-6)


def test_update_fb_then_update_actor_on_explicit_tensors_against_oracle():
    """The public update_fb / update_actor pair (fb_ddpg.py:291-421) on explicit tensors, with future_ratio > 0 configured and a
    stale hindsight mask left behind by update(): z must be used as given."""
    from controllable_agent_b200 import FBDDPGAgent, ReplayBuffer, _lib as L
    from gpu_common import read_tensors
    from oracle import fb_oracle as O
    d = O.Dims(obs_dim=11, action_dim=3, z_dim=10, goal_dim=11, hidden_dim=48, feature_dim=24, backward_hidden_dim=30)
    B = 32
    torch.manual_seed(3)
    np.random.seed(3)
    agent = FBDDPGAgent(obs_type="states", obs_shape=(d.obs_dim,), action_shape=(d.action_dim,), device="cuda", num_expl_steps=0,
                        update_encoder=True, goal_space=None, use_tb=True, use_wandb=False, use_hiplog=False, hidden_dim=d.hidden_dim,
                        feature_dim=d.feature_dim, backward_hidden_dim=d.backward_hidden_dim, z_dim=d.z_dim, batch_size=B,
                        update_every_steps=1, future_ratio=0.5, rng_mode="reference")
    rs = np.random.RandomState(1)
    replay = ReplayBuffer(4, 0.98, 0.9)
    for _ in range(4):
        replay.add_episode(O.synthetic_episode(rs, 20, d))
    agent.update(replay, 0)   # leaves a hindsight mask and future rows behind
    cpu = lambda net: {k: v.detach().cpu().clone() for k, v in net.named_parameters()}   # noqa: E731
    actor, fwd, bwd = cpu(agent.actor), cpu(agent.forward_net), cpu(agent.backward_net)
    fwd_t, bwd_t = cpu(agent.forward_target_net), cpu(agent.backward_target_net)
    g = torch.Generator().manual_seed(9)
    obs, nobs = torch.randn(B, d.obs_dim, generator=g), torch.randn(B, d.obs_dim, generator=g)
    act = torch.rand(B, d.action_dim, generator=g) * 2 - 1
    disc = torch.full((B, 1), 0.98)
    z = O.sample_z(B, d.z_dim, g)
    m = agent.update_fb(obs, act, disc, nobs, nobs, z, 1)
    e = agent.engine
    noise_fb, noise_actor = e.view("noise_fb").cpu(), e.view("noise_actor").cpu()
    assert rel(e.view("z"), z) < 1e-6, "z is used as given"
    ora = O.fb_loss_and_grads(fwd, bwd, fwd_t, bwd_t, actor, obs, act, disc, nobs, nobs, z, noise_fb, 0.2, 0.3, 1.0, d.z_dim)
    assert m["fb_loss"] == pytest.approx(ora["metrics"]["fb_loss"], rel=1e-3)
    fwd1 = read_tensors(e, L.NET_FORWARD, "param")
    ma = agent.update_actor(obs, z, 1)
    ora_a = O.actor_loss_and_grads(actor, fwd1, obs, z, noise_actor, 0.2, 0.3)
    assert ma["actor_loss"] == pytest.approx(float(ora_a["actor_loss"]), rel=1e-3)


def test_workspace_shaped_offline_loop(tmp_path):
    """What url_benchmark's offline workspace does with an agent and a replay buffer, in its order (train_offline.py:56-134,
    pretrain.py:147-206,374-435,437-494): ctor pokes (`agent.cfg.update_every_steps = 1`, `_future`, `_discount`, `_max_episodes`
    from `_storage`), the train loop (update -> metrics dict -> checkpoint of {agent, global_step, ...} excluding the replay),
    eval (zero-shot z from rewards through infer_meta_from_obs_and_rewards, act() per environment step under eval_mode, z_correl),
    then a fresh workspace that reloads the checkpoint through agent.init_from and keeps training on the same trajectory."""
    from controllable_agent_b200 import FBDDPGAgent, ReplayBuffer
    O_, A_, Z, B = 24, 6, 50, 128
    torch.manual_seed(1)
    np.random.seed(1)
    base = dict(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True, goal_space=None,
                use_tb=True, use_wandb=False, use_hiplog=True, batch_size=B, z_dim=Z, hidden_dim=256, feature_dim=128, backward_hidden_dim=134,
                num_inference_steps=300)

    class Workspace:
        _CHECKPOINTED_KEYS = ("agent", "global_step", "global_episode", "replay_loader")

        def __init__(self, replay=None):
            self.agent = FBDDPGAgent(**base)
            assert isinstance(self.agent, FBDDPGAgent)           # the isinstance gates of pretrain.py:151-155,405-408
            self.agent.cfg.update_every_steps = 1                # train_offline.py:59
            self.global_step, self.global_episode = 0, 0
            self.replay_loader = replay
            if replay is not None:
                replay._future, replay._discount = 0.99, 0.98    # train_offline.py:92-95
                replay._max_episodes = len(replay._storage["discount"])

        def save_checkpoint(self, fp, exclude=()):
            payload = {k: self.__dict__[k] for k in self._CHECKPOINTED_KEYS if k not in exclude}
            with open(fp, "wb") as f:
                torch.save(payload, f, pickle_protocol=4)

        def load_checkpoint(self, fp, only=None):
            with open(fp, "rb") as f:
                payload = torch.load(f, weights_only=False)
            for name, val in payload.items():
                if only is not None and name not in only:
                    continue
                if name == "agent":
                    self.agent.init_from(val)
                elif name == "replay_loader":
                    val._current_episode.clear()
                    val._max_episodes = len(val._storage["discount"])
                    self.replay_loader = val
                else:
                    setattr(self, name, val)

    rs = np.random.RandomState(0)
    replay = ReplayBuffer(max_episodes=6, discount=0.98, future=0.99)
    for _ in range(6):
        n = 41
        replay.add_episode({"observation": rs.randn(n, O_).astype(np.float32), "action": rs.uniform(-1, 1, (n, A_)).astype(np.float32),
                            "reward": rs.uniform(0, 1, (n, 1)).astype(np.float32), "discount": np.ones((n, 1), np.float32),
                            "physics": rs.randn(n, 4).astype(np.float32)})
    ws = Workspace(replay)
    logged = []
    ckpt = tmp_path / "latest.pt"
    for _ in range(6):
        metrics = ws.agent.update(ws.replay_loader, ws.global_step)
        assert {"fb_loss", "orth_loss", "fb_opt_lr"} <= set(metrics) and all(np.isfinite(v) for v in metrics.values())
        logged.append(metrics)
        ws.global_step += 1
        if ws.global_step % 3 == 0:
            ws.save_checkpoint(ckpt, exclude=["replay_loader"])
    assert logged[-1]["fb_loss"] != logged[0]["fb_loss"]

    # eval: zero-shot inference of z from (next_obs, reward) samples, then act() per step under eval_mode semantics
    class Reward:
        @staticmethod
        def from_physics(p):
            return float(p[0])
    obs_list, reward_list, n = [], [], 0
    while n < ws.agent.cfg.num_inference_steps:
        batch = ws.replay_loader.sample(B, custom_reward=Reward()).to("cuda")
        obs_list.append(batch.next_obs)
        reward_list.append(batch.reward)
        n += batch.next_obs.size(0)
    obs_t = torch.cat(obs_list, 0)[:ws.agent.cfg.num_inference_steps]
    rew_t = torch.cat(reward_list, 0)[:ws.agent.cfg.num_inference_steps]
    meta = ws.agent.infer_meta_from_obs_and_rewards(obs_t, rew_t)
    assert meta["z"].shape == (Z,) and abs(np.linalg.norm(meta["z"]) - np.sqrt(Z)) < 1e-3
    was_training = ws.agent.training
    ws.agent.train(False)
    action = ws.agent.act(rs.randn(O_).astype(np.float32), meta, ws.global_step, eval_mode=True)
    ws.agent.train(was_training)
    assert action.shape == (A_,) and np.all(np.abs(action) <= 1.0)

    class TS:
        observation = rs.randn(O_).astype(np.float32)
    assert -1.0 <= ws.agent.compute_z_correl(TS(), meta) <= 1.0
    meta2 = ws.agent.infer_meta(ws.replay_loader)
    assert meta2["z"].shape == (Z,)

    # a fresh workspace reloads the checkpoint (agent.init_from of the pickled agent) and continues: same parameters, same Adam
    # state, therefore the same next update as the original agent given the same RNG streams
    ws2 = Workspace(replay)
    ws2.load_checkpoint(ckpt)
    assert ws2.global_step == 6
    for name in ("actor", "forward_net", "backward_net", "forward_target_net", "backward_target_net"):
        for p, q in zip(getattr(ws.agent, name).parameters(), getattr(ws2.agent, name).parameters()):
            assert torch.equal(p.data, q.data), name
    assert ws2.agent.engine.get_adam_steps() == ws.agent.engine.get_adam_steps() == (6, 6)
    sd, sd2 = ws.agent.fb_opt.state_dict(), ws2.agent.fb_opt.state_dict()
    assert float(sd["state"][0]["step"]) == float(sd2["state"][0]["step"]) == 6.0
    assert torch.equal(sd["state"][0]["exp_avg"], sd2["state"][0]["exp_avg"])
    m1 = ws.agent.update(ws.replay_loader, 6)
    m2 = ws2.agent.update(ws2.replay_loader, 6)
    assert m1.keys() == m2.keys()   # (device RNG streams of the two agents differ: values are not compared)
