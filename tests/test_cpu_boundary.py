"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/fb_b200.h declares, the
parameter layout matches the reference's registration order and sizes, the config surface matches the reference's
dataclass field for field, and the product refuses to run without CUDA (no CPU fallback)."""
import ctypes as C
import dataclasses
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, subtree
from oracle import fb_oracle as O


def _lib():
    from controllable_agent_b200 import _lib as L
    return L, L.load()


def test_library_exports_every_declared_symbol():
    L, lib = _lib()
    header = open(os.path.join(ROOT, "include", "fb_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 27
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/fb_b200.h but not exported"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)   # the ctypes table binds exactly the header
    assert lib.fb_abi_version() == L.FB_ABI_VERSION
    assert lib.fb_error_string(-1) == b"bad argument"


def _create(L, lib, d, batch=64, **kw):
    c = L.fb_config(abi_version=L.FB_ABI_VERSION, batch=batch, global_batch=kw.get("global_batch", batch), row_offset=kw.get("row_offset", 0),
                    obs_dim=d.obs_dim, action_dim=d.action_dim, z_dim=d.z_dim, goal_dim=d.goal_dim, hidden_dim=d.hidden_dim,
                    feature_dim=d.feature_dim, backward_hidden_dim=d.backward_hidden_dim, use_goal=kw.get("use_goal", 0), rng_device=0, contract_mode=kw.get("contract_mode", 0), mlp_mode=kw.get("mlp_mode", 0),
                    ortho_coef=1.0, mix_ratio=kw.get("mix_ratio", 0.5), beta1=0.9, beta2=0.999, adam_eps=1e-8, seed=0,
                    future_ratio=kw.get("future_ratio", 0.0), q_loss=kw.get("q_loss", 0), q_loss_coef=0.01, no_norm_z=kw.get("no_norm_z", 0),
                    rand_weight=kw.get("rand_weight", 0), add_trunk=kw.get("add_trunk", 0), no_preprocess=kw.get("no_preprocess", 0),
                    boltzmann=kw.get("boltzmann", 0), temp=0.7, log_std_min=kw.get("log_std_min", -5.0), log_std_max=2.0,
                    debug_identity_b=kw.get("debug", 0))
    h = C.c_void_p()
    return lib.fb_create(C.byref(c), C.byref(h)), h


def _table(lib, h, net):
    out = []
    for i in range(lib.fb_num_tensors(h, net)):
        off, rows, cols = C.c_size_t(), C.c_int(), C.c_int()
        name = C.create_string_buffer(64)
        assert lib.fb_tensor_info(h, net, i, C.byref(off), C.byref(rows), C.byref(cols), name, 64) == 0
        out.append((name.value.decode(), off.value, (rows.value, cols.value) if cols.value else (rows.value,)))
    return out


@pytest.mark.parametrize("d", [O.Dims(), O.Dims(obs_dim=78, action_dim=12, goal_dim=2), O.Dims(obs_dim=17, z_dim=100, goal_dim=17),
                               O.Dims(obs_dim=11, action_dim=3, z_dim=10, goal_dim=11, hidden_dim=48, feature_dim=24, backward_hidden_dim=30)])
def test_flat_layout_matches_reference_registration_order(d):
    L, lib = _lib()
    rc, h = _create(L, lib, d, use_goal=int(d.goal_dim != d.obs_dim))
    assert rc == 0
    for net, spec in ((L.NET_FORWARD, O.forward_map_spec(d)), (L.NET_BACKWARD, O.backward_map_spec(d)), (L.NET_ACTOR, O.actor_spec(d))):
        tab = _table(lib, h, net)
        assert [(n, s) for n, _, s in tab] == [(n, tuple(s)) for n, s in spec]
        offs = [o for _, o, _ in tab]
        assert offs == sorted(offs) and all(o % 32 == 0 for o in offs)           # 128-byte aligned, non-overlapping
        for (n, o, s), (_, o2, _) in zip(tab, tab[1:]):
            assert o + int(np.prod(s)) <= o2
    n_fb, n_actor = lib.fb_flat_size(h, 0), lib.fb_flat_size(h, 1)
    assert n_fb % 4 == 0 and n_actor % 4 == 0
    last = _table(lib, h, L.NET_BACKWARD)[-1]
    assert last[1] + int(np.prod(last[2])) <= n_fb
    assert lib.fb_workspace_bytes(h) > 0
    lib.fb_destroy(h)


@pytest.mark.parametrize("d", [O.Dims(), O.Dims(obs_dim=11, action_dim=3, z_dim=10, goal_dim=11, hidden_dim=48, feature_dim=24, backward_hidden_dim=30)])
def test_add_trunk_layout_and_optional_branch_plans(d):
    """cfg.add_trunk (fb_modules.py:96-100,169-173): trunk tensors sit between the embeds and the heads, whose first layer then
    reads hidden_dim columns.  fb_create runs the whole plan builder (sizing pass), so every combination of the optional branches
    (q_loss, norm_z off, rand_weight, add_trunk, hindsight) must at least plan on both GEMM paths."""
    L, lib = _lib()
    rc, h = _create(L, lib, d, add_trunk=1)
    assert rc == 0
    for net, spec in ((L.NET_FORWARD, O.forward_map_spec(d, True)), (L.NET_BACKWARD, O.backward_map_spec(d)), (L.NET_ACTOR, O.actor_spec(d, True))):
        assert [(n, s) for n, _, s in _table(lib, h, net)] == [(n, tuple(s)) for n, s in spec]
    base_ws = lib.fb_workspace_bytes(h)
    lib.fb_destroy(h)
    rc, h = _create(L, lib, d)
    assert rc == 0 and lib.fb_workspace_bytes(h) < base_ws   # the trunk activations cost workspace
    lib.fb_destroy(h)
    rc, h = _create(L, lib, d, no_preprocess=1)   # preprocess = False: trunk.{0,1,3,5} then the heads (fb_modules.py:102-104,175-177)
    assert rc == 0
    for net, spec in ((L.NET_FORWARD, O.forward_map_spec(d, preprocess=False)), (L.NET_ACTOR, O.actor_spec(d, preprocess=False))):
        assert [(n, s) for n, _, s in _table(lib, h, net)] == [(n, tuple(s)) for n, s in spec]
    lib.fb_destroy(h)
    rc, h = _create(L, lib, d, boltzmann=1)   # boltzmann = True: the DiagGaussianActor's one stack (fb_modules.py:129-139); forward_net unchanged
    assert rc == 0
    for net, spec in ((L.NET_FORWARD, O.forward_map_spec(d)), (L.NET_ACTOR, O.boltzmann_actor_spec(d))):
        assert [(n, s) for n, _, s in _table(lib, h, net)] == [(n, tuple(s)) for n, s in spec]
    lib.fb_destroy(h)
    assert _create(L, lib, d, boltzmann=1, log_std_min=3.0)[0] == -1   # empty log-std interval
    dz = dataclasses.replace(d, z_dim=d.goal_dim)   # debug = True: identity backward map, z in goal space; no "B.*" tensors
    rc, h = _create(L, lib, dz, debug=1)
    assert rc == 0 and _table(lib, h, L.NET_BACKWARD) == []
    assert [(n, s) for n, _, s in _table(lib, h, L.NET_FORWARD)] == [(n, tuple(s)) for n, s in O.forward_map_spec(dz)]
    lib.fb_destroy(h)
    if d.z_dim != d.goal_dim:
        assert _create(L, lib, d, debug=1)[0] == -1
    for bits in range(64):   # ... and it plans with every other branch
        kw = dict(q_loss=bits & 1, no_norm_z=(bits >> 1) & 1, rand_weight=(bits >> 2) & 1, add_trunk=(bits >> 3) & 1,
                  future_ratio=0.3 if bits & 16 else 0.0, boltzmann=(bits >> 5) & 1, debug=1)
        for mlp_mode in (0, 1):
            rc, h = _create(L, lib, dz, mlp_mode=mlp_mode, contract_mode=mlp_mode, **kw)
            assert rc == 0, (kw, mlp_mode)
            lib.fb_destroy(h)
    for bits in range(128):
        kw = dict(q_loss=bits & 1, no_norm_z=(bits >> 1) & 1, rand_weight=(bits >> 2) & 1, add_trunk=(bits >> 3) & 1,
                  future_ratio=0.3 if bits & 16 else 0.0, no_preprocess=(bits >> 5) & 1, boltzmann=(bits >> 6) & 1)
        for mlp_mode in (0, 1):
            rc, h = _create(L, lib, d, mlp_mode=mlp_mode, contract_mode=mlp_mode, **kw)
            assert rc == 0, (kw, mlp_mode)
            lib.fb_destroy(h)
    assert _create(L, lib, dataclasses.replace(d, z_dim=119, obs_dim=d.obs_dim), q_loss=1)[0] == -3      # the fp64 inverse lives in shared memory
    assert _create(L, lib, dataclasses.replace(d, z_dim=129), rand_weight=1)[0] == -3


def test_default_layout_parameter_counts():
    """forward_net 3 363 940, backward_net 317 754, actor 2 211 846 parameters."""
    L, lib = _lib()
    rc, h = _create(L, lib, O.Dims(), batch=1024)
    assert rc == 0
    counts = [sum(int(np.prod(s)) for _, _, s in _table(lib, h, net)) for net in (L.NET_FORWARD, L.NET_BACKWARD, L.NET_ACTOR)]
    assert counts == [3363940, 317754, 2211846]
    lib.fb_destroy(h)


def test_create_rejects_bad_configs():
    L, lib = _lib()
    d = O.Dims()
    assert _create(L, lib, d, batch=1)[0] == -1                                      # batch < 2: no off-diagonal entries
    assert _create(L, lib, d, batch=64, global_batch=32)[0] == -1                    # local rows exceed the global batch
    assert _create(L, lib, d, batch=64, global_batch=128, row_offset=96)[0] == -1    # row block sticks out
    assert _create(L, lib, dataclasses.replace(d, goal_dim=3))[0] == -1              # goal_dim != obs_dim without a goal space
    assert _create(L, lib, dataclasses.replace(d, hidden_dim=4096))[0] == -3         # LayerNorm width beyond the kernels' limit
    rc, h = _create(L, lib, d, batch=64, global_batch=128, row_offset=64)
    assert rc == 0
    lib.fb_destroy(h)
    # calls on an unbound handle fail with FB_E_STATE instead of touching the device
    rc, h = _create(L, lib, d)
    assert lib.fb_run(h, L.PHASE_ALL, 0, None) == -2
    assert lib.fb_set_z(h, None, None) == -2
    lib.fb_destroy(h)


def test_batch_row_layout_is_16_byte_aligned():
    L, lib = _lib()
    offs, pitch = (C.c_int32 * 9)(), C.c_int32()
    assert lib.fb_batch_row_layout(24, 6, 3, 50, 1, offs, C.byref(pitch)) == 0
    o = list(offs)
    assert all(x % 4 == 0 for x in o) and pitch.value % 4 == 0
    assert o == sorted(o) and o[0] == 0 and o[1] == 24 and o[2] == 32 and o[3] == 36 and o[4] == 60 and o[5] == 64 and o[6] == 68
    assert pitch.value == 68 + 52 + 24 + 4


def test_config_surface_matches_reference_dataclass():
    """Every field of the reference's FBDDPGAgentConfig (fb_ddpg.py:37-82), same order and same default."""
    from controllable_agent_b200.agent import FBDDPGAgentConfig
    ref_fields = [
        ("_target_", None), ("name", "fb_ddpg"), ("obs_type", "???"), ("obs_shape", "???"), ("action_shape", "???"), ("device", "${device}"),
        ("lr", 1e-4), ("lr_coef", 1), ("fb_target_tau", 0.01), ("update_every_steps", 2), ("use_tb", "${use_tb}"), ("use_wandb", "${use_wandb}"),
        ("use_hiplog", "${use_hiplog}"), ("num_expl_steps", "???"), ("num_inference_steps", 5120), ("hidden_dim", 1024),
        ("backward_hidden_dim", 526), ("feature_dim", 512), ("z_dim", 50), ("stddev_schedule", "0.2"), ("stddev_clip", 0.3),
        ("update_z_every_step", 300), ("update_z_proba", 1.0), ("nstep", 1), ("batch_size", 1024), ("init_fb", True),
        ("update_encoder", "${update_encoder}"), ("goal_space", "${goal_space}"), ("ortho_coef", 1.0), ("log_std_bounds", (-5, 2)),
        ("temp", 1), ("boltzmann", False), ("debug", False), ("future_ratio", 0.0), ("mix_ratio", 0.5), ("rand_weight", False),
        ("preprocess", True), ("norm_z", True), ("q_loss", False), ("q_loss_coef", 0.01), ("additional_metric", False), ("add_trunk", False)]
    mine = dataclasses.fields(FBDDPGAgentConfig)
    assert [f.name for f in mine[:len(ref_fields)]] == [n for n, _ in ref_fields]
    cfg = FBDDPGAgentConfig()
    for name, default in ref_fields[1:]:
        assert getattr(cfg, name) == default, name
    assert cfg._target_.endswith(".FBDDPGAgent")


def test_no_cpu_fallback():
    from controllable_agent_b200 import FBDDPGAgent, FBStepEngine, EngineConfig
    with pytest.raises(RuntimeError, match="no CPU"):
        FBStepEngine(EngineConfig(batch=8, obs_dim=4, action_dim=2, z_dim=4, goal_dim=4), "cpu")
    with pytest.raises(RuntimeError, match="CPU fallback"):
        FBDDPGAgent(obs_type="states", obs_shape=(24,), action_shape=(6,), device="cpu", num_expl_steps=0, update_encoder=True,
                    goal_space=None, use_tb=False, use_wandb=False, use_hiplog=False)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            FBStepEngine(EngineConfig(batch=8, obs_dim=4, action_dim=2, z_dim=4, goal_dim=4), "cuda")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "controllable_agent_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.lower() or fn == "README.md", f"{fn} mentions the oracle"


def test_module_mirrors_match_oracle_forward():
    """The host-side nn.Module mirrors (act / infer_meta plumbing) compute what the oracle restatement computes."""
    from controllable_agent_b200 import modules as M
    torch.manual_seed(0)
    d = O.Dims(obs_dim=11, action_dim=3, z_dim=10, goal_dim=11, hidden_dim=48, feature_dim=24, backward_hidden_dim=30)
    actor, fwd, bwd = M.Actor(11, 10, 3, 24, 48), M.ForwardMap(11, 10, 3, 24, 48), M.BackwardMap(11, 10, 30)
    assert [n for n, _ in actor.named_parameters()] == [n for n, _ in O.actor_spec(d)]
    assert [n for n, _ in fwd.named_parameters()] == [n for n, _ in O.forward_map_spec(d)]
    assert [n for n, _ in bwd.named_parameters()] == [n for n, _ in O.backward_map_spec(d)]
    obs, z, act = torch.randn(5, 11), torch.randn(5, 10), torch.rand(5, 3) * 2 - 1
    with torch.no_grad():
        pa, pf, pb = (dict(m.named_parameters()) for m in (actor, fwd, bwd))
        assert torch.allclose(actor(obs, z, 0.2).mean, O.actor_mean(pa, obs, z), atol=1e-6)
        f1, f2 = fwd(obs, z, act)
        o1, o2 = O.forward_map(pf, obs, z, act)
        assert torch.allclose(f1, o1, atol=1e-6) and torch.allclose(f2, o2, atol=1e-6)
        assert torch.allclose(bwd(obs), O.backward_map(pb, obs, 10), atol=1e-6)
        # cfg.boltzmann: DiagGaussianActor + SquashedNormal (fb_modules.py:129-151, utils.py:188-233)
        ga = M.DiagGaussianActor(11, 10, 3, 48, (-5.0, 2.0))
        assert [n for n, _ in ga.named_parameters()] == [n for n, _ in O.boltzmann_actor_spec(d)]
        dist = ga(obs, z)
        mu, std = O.diag_gaussian_actor(dict(ga.named_parameters()), obs, z)
        assert torch.allclose(dist.loc, mu, atol=1e-6) and torch.allclose(dist.scale, std, atol=1e-6) and torch.allclose(dist.mean, torch.tanh(mu))
        x = mu + std * torch.randn(5, 3)
        assert torch.allclose(dist.log_prob(torch.tanh(x)), O.squashed_normal_log_prob(x, mu, std), atol=2e-4)
        torch.manual_seed(3); a1 = dist.sample()
        torch.manual_seed(3); a2 = dist.rsample()
        assert torch.allclose(a1, a2, atol=1e-6) and a1.abs().max() <= 1.0   # torch.normal and loc + scale * randn consume the CPU generator alike
    M.hard_update_params(torch.nn.Identity(), torch.nn.Identity())      # the states-only agent's encoder: nothing to copy, no error
    M.soft_update_params(torch.nn.Identity(), torch.nn.Identity(), 0.01)
    tgt = M.BackwardMap(11, 10, 30)
    M.soft_update_params(bwd, tgt, 0.25)
    M.hard_update_params(bwd, tgt)
    assert all(torch.equal(a, b) for a, b in zip(bwd.parameters(), tgt.parameters()))
    assert M.schedule("linear(1,0.2,200)", 100) == pytest.approx(0.6)
    assert M.schedule("step_linear(1,0.5,100,0.1,100)", 150) == pytest.approx(0.3)
    # same construction order + same init calls => same parameters as the reference for the same torch seed
    g = load_golden("update_small")
    ref = subtree(g, "param0/actor")
    assert tuple(ref["obs_net.0.weight"].shape) == (48, 11)


@pytest.mark.parametrize("case,dims,seed", [("small", (11, 10, 3, 24, 48, 30, 11), 11), ("wide", (24, 50, 6, 64, 128, 70, 24), 37)])
def test_seeded_init_reproduces_reference_parameters(case, dims, seed):
    """Building the mirrors in the reference's order (Actor, ForwardMap, BackwardMap, ... fb_ddpg.py:117-139) after the same
    torch.manual_seed consumes the CPU generator identically: the 2-D weights equal the reference agent's (to the ulp-level
    differences LAPACK's QR shows between thread counts; a different random stream would differ by O(0.1))."""
    from controllable_agent_b200 import modules as M
    O_, Z, A, Fd, H, Hb, G = dims
    g = load_golden(f"update_{case}")
    torch.manual_seed(seed)
    actor = M.Actor(O_, Z, A, Fd, H)
    fwd = M.ForwardMap(O_, Z, A, Fd, H)
    bwd = M.BackwardMap(G, Z, Hb)
    for net, key in ((actor, "actor"), (fwd, "forward_net"), (bwd, "backward_net")):
        for name, p in net.named_parameters():
            if p.dim() == 2:   # the fixture perturbs only 1-D tensors of the online nets
                np.testing.assert_allclose(p.detach().numpy(), g[f"param0/{key}/{name}"], rtol=0, atol=2e-5, err_msg=f"{key}/{name}")


def test_replay_load_reads_episode_files_in_order_until_full(tmp_path):
    """ReplayBuffer.load (in_memory_replay_buffer.py:192-208): sorted `.npz` episodes, stop when the buffer is full.  The device
    commit (add_episode) is GPU-tested; here it is replaced by a recorder that keeps the reference's ring bookkeeping."""
    from controllable_agent_b200.replay import ReplayBuffer, load_episode

    class Recorder(ReplayBuffer):
        def add_episode(self, episode):
            self.seen = getattr(self, "seen", []) + [episode]
            self._idx = (self._idx + 1) % self._max_episodes
            self._full = self._full or self._idx == 0

    rs = np.random.RandomState(0)
    for i in (3, 1, 2, 0, 4):
        np.savez(tmp_path / f"episode_{i:03d}.npz", observation=rs.standard_normal((6, 5)).astype(np.float32) + i,
                 action=np.full((6, 2), i, np.float32), reward=np.zeros((6, 1), np.float32), discount=np.ones((6, 1), np.float32))
    ep = load_episode(tmp_path / "episode_002.npz")
    assert set(ep) == {"observation", "action", "reward", "discount"} and ep["action"][0, 0] == 2
    buf = Recorder(max_episodes=4, discount=0.98, future=0.99)
    buf.load(None, tmp_path, relabel=False)
    assert [int(e["action"][0, 0]) for e in buf.seen] == [0, 1, 2, 3] and buf._full      # the fifth file is not read: buffer full
    big = Recorder(max_episodes=8, discount=0.98, future=0.99)
    big.load(None, str(tmp_path), relabel=False)
    assert len(big.seen) == 5 and not big._full and len(big) == 5


def test_episode_batch_collate_to_and_reward_masking():
    """Reads like url_benchmark/test_replay_buffer.py:10-32, on this package's EpisodeBatch (same container contract)."""
    from controllable_agent_b200 import EpisodeBatch
    shapes = dict(obs=(4, 12), action=(5, 11), next_obs=(6, 10))
    meta = dict(a=np.random.rand(16), b=np.random.rand(17))
    batch = EpisodeBatch(reward=np.array([1.0]), discount=np.array([0.5]), meta=meta, **{x: np.random.rand(*y) for x, y in shapes.items()})
    batches = EpisodeBatch.collate_fn([batch, batch])
    assert batches.obs.shape == (2, 4, 12)
    assert isinstance(batches.meta, dict) and len(batches.meta) == 2 and batches.meta["a"].shape == (2, 16)
    cpu = batch.to("cpu")
    assert cpu.reward.shape == (1,) and cpu.goal is None
    batches = EpisodeBatch.collate_fn([cpu, cpu])
    assert batches.reward.shape == (2, 1)
    no_reward = batches.with_no_reward()
    assert not no_reward.reward.abs().sum(), "reward should be masked"
    assert batches.reward.abs().sum(), "reward should not be masked"
    assert no_reward.obs is batches.obs, "observations must not be copied"
    assert [x.shape for x in batches.unpack()] == [(2, 4, 12), (2, 5, 11), (2, 1), (2, 1), (2, 6, 10)]
    mixed = EpisodeBatch(reward=np.array([1.0]), discount=np.array([0.5]), goal=np.zeros(3), **{x: np.random.rand(*y) for x, y in shapes.items()})
    with pytest.raises(RuntimeError, match="mixed with Nones"):
        EpisodeBatch.collate_fn([batch, mixed])


class _RefLayoutHostReplay:
    """Attribute layout of url_benchmark.in_memory_replay_buffer.ReplayBuffer (:66-88) with a filled `_storage`."""

    def __init__(self, rs, E, T, O, A, G, future, ragged=False):
        self._max_episodes, self._discount, self._future = E, 0.98, future
        self._storage = {"observation": rs.standard_normal((E, T + 1, O)).astype(np.float32), "action": rs.random((E, T + 1, A)).astype(np.float32),
                         "reward": rs.random((E, T + 1, 1)).astype(np.float32), "discount": (rs.random((E, T + 1, 1)) > 0.1).astype(np.float32),
                         "physics": np.zeros((E, T + 1, 3), np.float32)}
        if G:
            self._storage["goal"] = rs.standard_normal((E, T + 1, G)).astype(np.float32)
        self._episodes_length = np.full(E, T, np.int32)
        if ragged:
            self._episodes_length = rs.integers(2, T + 1, size=E).astype(np.int32)
        self._is_fixed_episode_length, self._episodes_selection_probability = not ragged, None
        self._idx, self._full = 0, True

    def __len__(self):
        return self._max_episodes


@pytest.mark.parametrize("G,future,ragged", [(0, 1.0, False), (3, 1.0, False), (0, 0.9, True), (2, 0.8, False)])
def test_host_gather_rows_matches_reference_sample_gathers(G, future, ragged):
    """fb_host_gather_rows = the fancy-index gathers of in_memory_replay_buffer.py:162-183 on host storage, written in the packed
    batch-row layout; HostStorageView + draw_sample_indices reproduce the reference's index draws (no GPU involved)."""
    from controllable_agent_b200.replay import HostStorageView, draw_sample_indices
    L, lib = _lib()
    rs = np.random.default_rng(5)
    E, T, O, A, B = 7, 19, 11, 3, 64
    rep = _RefLayoutHostReplay(rs, E, T, O, A, G, future, ragged)
    view = HostStorageView.adopt(rep)
    assert view is not None and view.still_valid()
    np.random.seed(3)
    ep, st, fu = draw_sample_indices(rep, B)
    assert (fu is None) == (future >= 1.0) and st.min() >= 1 and (st <= rep._episodes_length[ep]).all()
    offs, pitch = (C.c_int32 * 9)(), C.c_int32()
    assert lib.fb_batch_row_layout(O, A, G, 0, int(fu is not None), offs, C.byref(pitch)) == 0
    rows = np.full((B, pitch.value), np.nan, np.float32)
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)  # noqa: E731
    ep32, st32, fu32 = i32(ep), i32(st), (i32(fu) if fu is not None else None)
    rc = lib.fb_host_gather_rows(C.byref(view.c), ep32.ctypes.data, st32.ctypes.data, fu32.ctypes.data if fu32 is not None else None, B,
                                 0.98, rows.ctypes.data, pitch.value)
    assert rc == 0
    S = rep._storage
    np.testing.assert_array_equal(rows[:, offs[0]:offs[0] + O], S["observation"][ep, st - 1])
    np.testing.assert_array_equal(rows[:, offs[1]:offs[1] + A], S["action"][ep, st])
    np.testing.assert_array_equal(rows[:, offs[2]], S["reward"][ep, st][:, 0])
    np.testing.assert_array_equal(rows[:, offs[2] + 1], (np.float32(0.98) * S["discount"][ep, st])[:, 0])
    np.testing.assert_array_equal(rows[:, offs[3]:offs[3] + O], S["observation"][ep, st])
    if G:
        np.testing.assert_array_equal(rows[:, offs[4]:offs[4] + G], S["goal"][ep, st - 1])
        np.testing.assert_array_equal(rows[:, offs[5]:offs[5] + G], S["goal"][ep, st])
    if fu is not None:
        np.testing.assert_array_equal(rows[:, offs[7]:offs[7] + O], S["observation"][ep, fu - 1])
        if G:
            np.testing.assert_array_equal(rows[:, offs[8]:offs[8] + G], S["goal"][ep, fu - 1])
    # out-of-range steps are refused, not read
    bad = st32.copy(); bad[0] = 0
    assert lib.fb_host_gather_rows(C.byref(view.c), ep32.ctypes.data, bad.ctypes.data, None, B, 0.98, rows.ctypes.data, pitch.value) == -1
    for bad_ep in (-1, view.c.max_episodes):   # and so are episode indices outside the storage
        bad = ep32.copy(); bad[-1] = bad_ep
        assert lib.fb_host_gather_rows(C.byref(view.c), bad.ctypes.data, st32.ctypes.data, None, B, 0.98, rows.ctypes.data, pitch.value) == -1
    assert view.c.max_episodes == S["observation"].shape[0]
    # a storage the library cannot read in place (float64 field) is not adopted: the caller falls back to the object's sample()
    rep._storage["action"] = rep._storage["action"].astype(np.float64)
    assert HostStorageView.adopt(rep) is None and not view.still_valid()


def test_draw_sample_indices_cheap_stream_has_the_same_support():
    """rng_mode="device" draws the replay indices of a host buffer through cheaper numpy calls (exact_stream=False): same ranges and the
    same handling of `future` as the reference's calls; ragged buffers fall back to the reference's calls themselves."""
    from controllable_agent_b200.replay import draw_sample_indices
    rs = np.random.default_rng(0)
    rep = _RefLayoutHostReplay(rs, 6, 25, 4, 2, 0, 0.8)
    np.random.seed(1)
    ep, st, fu = draw_sample_indices(rep, 4096, exact_stream=False)
    assert ep.min() == 0 and ep.max() == 5 and st.min() == 1 and st.max() == 25
    assert fu is not None and (fu >= st).all() and fu.max() == 25 and (fu > st).any()
    rep._future = 1.0
    assert draw_sample_indices(rep, 64, exact_stream=False)[2] is None
    # an appended episode (the length vector changes) is seen on the next draw
    rep._episodes_length = np.concatenate([rep._episodes_length, [25]]).astype(np.int32)
    rep._max_episodes = 7
    ep, st, _ = draw_sample_indices(rep, 4096, exact_stream=False)
    assert ep.max() == 6
    ragged = _RefLayoutHostReplay(rs, 6, 25, 4, 2, 0, 1.0, ragged=True)
    np.random.seed(2)
    a = draw_sample_indices(ragged, 256, exact_stream=False)
    np.random.seed(2)
    b = draw_sample_indices(ragged, 256, exact_stream=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and (a[1] <= ragged._episodes_length[a[0]]).all()
