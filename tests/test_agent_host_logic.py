"""Host logic of FBDDPGAgent without a GPU: the engine (the ctypes owner of the CUDA library) is replaced by a recorder, so that
what the Python layer itself decides is checked on the CPU —

  * which config branches are refused (and that nothing but CUDA is accepted as a device),
  * the gate on update_every_steps and the keys of the metrics dict per config flag (fb_ddpg.py:356-377,413-418,430-431),
  * rng_mode="reference": the draws handed to the library (z, permutation, mix mask, rand_weight rows, hindsight mask, both action
    noises) are the ones the reference makes from the same seeds, in the reference's order — compared with
    the oracle agent, which is pinned against the reference's trajectories.

The arithmetic of the step is not exercised here (the recorder computes nothing): that is what the `-m gpu` tests are for."""
import collections

import numpy as np
import pytest
import torch

from oracle import fb_oracle as O

D = O.Dims(obs_dim=11, action_dim=3, z_dim=10, goal_dim=11, hidden_dim=48, feature_dim=24, backward_hidden_dim=30)
BATCH = 32


class RecorderEngine:
    """Stands in for controllable_agent_b200.engine.FBStepEngine: same attributes, CPU tensors, records every call."""

    def __init__(self, cfg, device):
        from controllable_agent_b200 import _lib as L
        self.cfg, self.device, self.has_nccl = cfg, torch.device("cpu"), False
        self.calls = []
        specs = {L.NET_FORWARD: O.forward_map_spec(D, cfg.add_trunk), L.NET_BACKWARD: O.backward_map_spec(D), L.NET_ACTOR: O.actor_spec(D, cfg.add_trunk)}
        self._views = {}
        for which in ("param", "grad", "m", "v", "target"):
            for net, spec in specs.items():
                if which == "target" and net == L.NET_ACTOR:
                    continue
                self._views[(net, which)] = collections.OrderedDict((n, torch.zeros(s)) for n, s in spec)
        flat = lambda *keys: torch.zeros(sum(v.numel() for k in keys for v in self._views[k].values()))  # noqa: E731
        self.param_fb, self.target_fb = flat((L.NET_FORWARD, "param"), (L.NET_BACKWARD, "param")), flat((L.NET_FORWARD, "param"), (L.NET_BACKWARD, "param"))
        self.param_actor = flat((L.NET_ACTOR, "param"))
        self.metrics = {k: float(i) for i, k in enumerate(L.METRIC_KEYS + L.OPTIONAL_METRIC_KEYS)}

    def tensors(self, net, which="param"):
        return self._views[(net, which)]

    def get_adam_steps(self):
        return (0, 0)

    def read_metrics(self):
        return dict(self.metrics)

    def launch_count(self, mask):
        return 0

    def __getattr__(self, name):   # set_scalars, set_indices, set_z, set_noise, run, bind_replay, upload_batch, ...
        def record(*args, **kw):
            self.calls.append((name, args, kw))
        return record

    def last(self, name):
        hits = [c for c in self.calls if c[0] == name]
        assert hits, (name, [c[0] for c in self.calls])
        return hits[-1]


@pytest.fixture
def make_agent(monkeypatch):
    import controllable_agent_b200.agent as A
    monkeypatch.setattr(A, "FBStepEngine", RecorderEngine)
    if not torch.cuda.is_available():   # the agent's pinned staging blocks / stream waits have nothing to talk to on a CPU-only box
        monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)

        class _NoStream:
            def synchronize(self):
                pass
        monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _NoStream())

    def make(**kw):
        base = dict(obs_type="states", obs_shape=(D.obs_dim,), action_shape=(D.action_dim,), device="cuda", num_expl_steps=0, update_encoder=True,
                    goal_space=None, use_tb=True, use_wandb=False, use_hiplog=False, hidden_dim=D.hidden_dim, feature_dim=D.feature_dim,
                    backward_hidden_dim=D.backward_hidden_dim, z_dim=D.z_dim, batch_size=BATCH, update_every_steps=1, rng_mode="reference")
        agent = A.FBDDPGAgent(**{**base, **kw})
        agent.draw_device = "cpu"
        return agent
    return make


class HostReplay:
    """A replay object with the reference's sample() contract over the oracle's host storage."""

    def __init__(self, seed, with_future=True):
        from controllable_agent_b200 import EpisodeBatch
        self._batch = EpisodeBatch
        rs = np.random.RandomState(seed)
        self.buf = O.OracleReplay(4, 0.98, 0.99 if with_future else 1.0)
        for _ in range(4):
            self.buf.add_episode(O.synthetic_episode(rs, 20, D))

    def sample(self, n):
        s = self.buf.sample(n)
        return self._batch(obs=s["obs"], action=s["action"], reward=s["reward"], discount=s["discount"], next_obs=s["next_obs"],
                           future_obs=s["future_obs"], meta=s["meta"])


def test_refused_configurations():
    from controllable_agent_b200 import FBDDPGAgent
    base = dict(obs_type="states", obs_shape=(24,), action_shape=(6,), device="cuda", num_expl_steps=0, update_encoder=True, goal_space=None,
                use_tb=False, use_wandb=False, use_hiplog=False)
    for kw in (dict(obs_type="pixels"),):
        with pytest.raises(NotImplementedError):
            FBDDPGAgent(**{**base, **kw})
    with pytest.raises(ValueError, match="identity"):      # debug=True: z must live in goal space (z_dim 50 != obs_dim 24)
        FBDDPGAgent(**{**base, "debug": True})
    with pytest.raises(ValueError):
        FBDDPGAgent(**{**base, "future_ratio": -0.1})
    with pytest.raises(RuntimeError, match="no CPU fallback|CPU"):
        FBDDPGAgent(**{**base, "device": "cpu"})
    if not torch.cuda.is_available():   # a supported configuration still needs the device: no silent CPU route
        with pytest.raises(RuntimeError, match="CUDA"):
            FBDDPGAgent(**base)


def test_update_gate_and_metric_keys(make_agent):
    from controllable_agent_b200 import _lib as L
    replay = HostReplay(0)
    agent = make_agent(update_every_steps=2)
    assert agent.update(replay, 1) == {}
    assert not [c for c in agent.engine.calls if c[0] == "run"]          # gated: nothing was enqueued (fb_ddpg.py:430-431)
    m = agent.update(replay, 2)
    assert set(m) == set(L.METRIC_KEYS) | {"fb_opt_lr"}
    assert agent.engine.last("run")[1][0] & L.PHASE_METRICS
    # metrics off: the step runs, the dict is empty
    quiet = make_agent(use_tb=False)
    assert quiet.update(replay, 0) == {}
    mask = quiet.engine.last("run")[1][0]
    assert mask & L.PHASE_FB_ADAM and mask & L.PHASE_ACTOR_ADAM and not mask & L.PHASE_METRICS
    # hiplog only: the update_fb block is reported, the actor block is not (fb_ddpg.py:356 vs :413)
    hip = make_agent(use_tb=False, use_hiplog=True)
    assert set(hip.update(replay, 0)) == set(L.METRIC_KEYS[:14]) | {"fb_opt_lr"}
    # optional keys follow their flags
    assert set(make_agent(q_loss=True).update(replay, 0)) == set(L.METRIC_KEYS) | {"fb_opt_lr", "q_loss"}
    assert set(make_agent(additional_metric=True).update(replay, 0)) == set(L.METRIC_KEYS) | {"fb_opt_lr", "q1_success"}
    assert set(make_agent(q_loss=True, use_tb=False, use_hiplog=True).update(replay, 0)) == set(L.METRIC_KEYS[:14]) | {"fb_opt_lr", "q_loss"}


@pytest.mark.parametrize("kw", [dict(), dict(mix_ratio=0.0), dict(future_ratio=0.4), dict(rand_weight=True), dict(norm_z=False, rand_weight=True, future_ratio=0.3)])
def test_reference_rng_mode_hands_over_the_reference_draws(make_agent, kw):
    """Same seeds -> the engine receives exactly the draws the oracle agent (== the reference, test_full_update_trajectory) makes."""
    seed = 123
    agent = make_agent(**kw)
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent.update(HostReplay(5), 0)
    e = agent.engine
    got_z = torch.as_tensor(e.last("set_z")[1][0])
    got_noise = [torch.as_tensor(x) for x in e.last("set_noise")[1]]
    idx = e.last("set_indices")[2]
    got_perm, got_mix = torch.as_tensor(idx["perm"]), np.asarray(idx["mix_mask"])

    # the oracle's draw sequence, replayed by hand (fb_oracle.OracleAgent.update_from_batch / update_fb / update_actor)
    torch.manual_seed(seed)
    np.random.seed(seed)
    HostReplay(5).sample(BATCH)                                  # numpy: episode / step / future indices
    norm_z = kw.get("norm_z", True)
    z = O.sample_z(BATCH, D.z_dim, norm_z=norm_z)                # torch: randn (+ rand when norm_z is off)
    perm = torch.randperm(BATCH)                                 # torch CPU
    mix_ratio = kw.get("mix_ratio", 0.5)
    mix = np.zeros(BATCH, bool)
    if mix_ratio > 0:
        mix = np.random.uniform(size=BATCH) < mix_ratio          # numpy
        if kw.get("rand_weight"):
            w = torch.rand(size=(int(mix.sum()), BATCH))         # torch CPU
            u = torch.rand(int(mix.sum()), 1)
    fut = None
    if kw.get("future_ratio", 0.0) > 0:
        fut = np.random.uniform(size=BATCH) < kw["future_ratio"]  # numpy, after the mix mask
    noise_fb, noise_actor = torch.randn(BATCH, D.action_dim), torch.randn(BATCH, D.action_dim)

    assert torch.equal(got_z, z) and torch.equal(got_perm, perm)
    np.testing.assert_array_equal(got_mix.astype(bool), mix)
    assert torch.equal(got_noise[0], noise_fb) and torch.equal(got_noise[1], noise_actor)
    if fut is not None:
        np.testing.assert_array_equal(np.asarray(e.last("set_future_mask")[1][0]).astype(bool), fut)
    else:
        assert not [c for c in e.calls if c[0] == "set_future_mask"]
    if kw.get("rand_weight") and mix_ratio > 0:
        W, U = (torch.as_tensor(x) for x in e.last("set_mix_weights")[1])
        rows = np.where(mix)[0]
        assert torch.equal(W[rows], w) and torch.equal(U[rows], u[:, 0])
    else:
        assert not [c for c in e.calls if c[0] == "set_mix_weights"]
    # the batch the library receives is the sampled one, field by field
    up = e.last("upload_batch")[1]
    torch.manual_seed(seed)
    np.random.seed(seed)
    ref = HostReplay(5).sample(BATCH)
    np.testing.assert_array_equal(np.asarray(up[0]), ref.obs)
    np.testing.assert_array_equal(np.asarray(up[1]), ref.action)
    np.testing.assert_array_equal(np.asarray(up[2]), ref.discount)
    np.testing.assert_array_equal(np.asarray(up[3]), ref.next_obs)


def test_hindsight_needs_a_replay_with_future_rows(make_agent):
    """fb_ddpg.py:463 asserts `future_goal is not None`; here: a clear error before anything is launched."""
    agent = make_agent(future_ratio=0.5)
    replay = HostReplay(3, with_future=False)
    replay._future = 1.0
    with pytest.raises(ValueError, match="future < 1"):
        agent.update(replay, 0)
    assert not [c for c in agent.engine.calls if c[0] == "run"]


def test_update_fb_on_explicit_tensors_clears_the_hindsight_mask(make_agent):
    """update_fb promises `z` is used as given: with future_ratio > 0 the hindsight mask of the previous update() is zeroed."""
    agent = make_agent(future_ratio=0.5)
    agent.update(HostReplay(4), 0)
    assert np.asarray(agent.engine.last("set_future_mask")[1][0]).sum() > 0   # update() drew hindsight rows
    g = torch.Generator().manual_seed(0)
    obs, nobs = torch.randn(BATCH, D.obs_dim, generator=g), torch.randn(BATCH, D.obs_dim, generator=g)
    agent.update_fb(obs, torch.zeros(BATCH, D.action_dim), torch.ones(BATCH, 1), nobs, nobs, torch.randn(BATCH, D.z_dim, generator=g), 0)
    assert np.asarray(agent.engine.last("set_future_mask")[1][0]).sum() == 0
    assert np.asarray(agent.engine.last("set_indices")[2]["mix_mask"]).sum() == 0


def test_optimizer_state_dict_reports_the_device_step_count(make_agent):
    """fb_opt / actor_opt are real torch optimizers whose `step` entries follow the device-side Adam counters whenever their
    state_dict() is taken (reference init_from, checkpoint code: fb_ddpg.py:173-175)."""
    agent = make_agent()
    assert isinstance(agent.fb_opt, torch.optim.Adam) and isinstance(agent.actor_opt, torch.optim.Optimizer)
    agent.engine.get_adam_steps = lambda: (7, 5)
    sd_fb, sd_actor = agent.fb_opt.state_dict(), agent.actor_opt.state_dict()
    assert {float(s["step"]) for s in sd_fb["state"].values()} == {7.0}
    assert {float(s["step"]) for s in sd_actor["state"].values()} == {5.0}
