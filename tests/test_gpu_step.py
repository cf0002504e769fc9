"""GPU parity of the FB-DDPG gradient step (through the C ABI) against
  (a) the golden fixtures produced by the UNMODIFIED reference (tests/golden/update_*.npz) and
  (b) the oracle restatement on fresh seeded inputs, at the reference's default widths.
Tolerance: 1e-3 relative (BASELINE.json north_star); typical observed error is ~1e-6."""
import numpy as np
import pytest
import torch

from conftest import load_golden, subtree
from gpu_common import REL_TOL, dims_from_params, golden_params, load_params, make_engine, read_tensors, rel
from oracle import fb_oracle as O

pytestmark = pytest.mark.gpu


def _L():
    from controllable_agent_b200 import _lib as L
    return L


@pytest.mark.parametrize("tile_cfg", [-1, 0, 1, 2])
@pytest.mark.parametrize("M,N,K,ak,bk,relu,splitk", [
    (128, 128, 64, 1, 1, 0, 1), (200, 150, 70, 1, 1, 1, 1), (1024, 512, 74, 1, 1, 0, 1), (96, 50, 128, 1, 1, 0, 1),
    (300, 526, 526, 1, 0, 0, 1), (526, 24, 1000, 0, 0, 0, 4), (50, 1024, 333, 0, 0, 0, 3), (130, 6, 48, 1, 0, 0, 1),
    (64, 64, 16, 0, 1, 0, 1), (190, 70, 90, 0, 1, 1, 1), (257, 129, 1025, 1, 1, 0, 2), (1, 7, 5, 1, 1, 0, 1), (33, 1, 3, 1, 0, 0, 1), (5, 3, 1, 0, 0, 0, 1)])
def test_sgemm_against_float64(M, N, K, ak, bk, relu, splitk, tile_cfg):
    L = _L()
    lib = L.load()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((M, K) if ak else (K, M), generator=g)
    Bm = torch.randn((N, K) if bk else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = (A.double() if ak else A.double().T) @ (Bm.double().T if bk else Bm.double()) + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    dA, dB, dbias = A.cuda(), Bm.cuda(), bias.cuda()
    dC = torch.zeros(M, N, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    L.check(lib.fb_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), dbias.data_ptr(), M, N, K, A.shape[1], Bm.shape[1], N,
                         ak, bk, relu, splitk, tile_cfg, s))
    torch.cuda.synchronize()
    assert rel(dC, ref) < 1e-5


@pytest.mark.parametrize("M,N,K,bk,relu", [
    (128, 128, 32, 1, 0), (128, 128, 64, 1, 0), (256, 128, 128, 1, 1), (1024, 512, 1024, 1, 1), (200, 150, 72, 1, 1), (96, 50, 128, 1, 0),
    (1024, 6, 1024, 1, 0), (300, 528, 528, 0, 0), (130, 8, 48, 0, 0), (257, 129, 1028, 1, 0), (64, 64, 8, 1, 0), (2048, 1024, 1024, 0, 0)])
@pytest.mark.parametrize("pairs", [False, True])   # True: CTA pairs (tcgen05 cta_group::2, 256-row work items)
def test_sgemm_tcgen05_3xtf32_against_float64(M, N, K, bk, relu, pairs):
    """The tensor-core GEMM (gemm_tc.cuh) holds fp32-grade accuracy: rel err vs float64 at the level of the SIMT fp32 kernel."""
    L = _L()
    lib = L.load()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((M, K), generator=g)
    Bm = torch.randn((N, K) if bk else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ (Bm.double().T if bk else Bm.double()) + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    dA, dB, dbias = A.cuda(), Bm.cuda(), bias.cuda()
    ldc = (N + 3) // 4 * 4
    dC = torch.zeros(M, ldc, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    L.check(lib.fb_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), dbias.data_ptr(), M, N, K, A.shape[1], Bm.shape[1], ldc,
                         1, bk, relu, 1, 4 if pairs else 3, s))
    torch.cuda.synchronize()
    err = rel(dC[:, :N], ref)
    print(f"tcgen05 3xTF32 M={M} N={N} K={K}: rel err vs float64 {err:.2e}")
    assert err < 1e-5, err
    assert float(dC[:, N:].abs().max()) == 0.0 if ldc > N else True


@pytest.mark.parametrize("M,N,K,bk,splitk", [(1024, 50, 1024, 1, 4), (2048, 6, 1024, 0, 8), (50, 1024, 1024, 1, 3), (257, 129, 1028, 1, 5),
                                              (128, 128, 64, 1, 2), (1024, 512, 2048, 0, 2)])
@pytest.mark.parametrize("pairs", [False, True])
def test_sgemm_tcgen05_split_k(M, N, K, bk, splitk, pairs):
    """split-K of the tensor-core GEMM: every k-range adds its partial tile into a zeroed C (the bias rides on the first)."""
    L = _L()
    lib = L.load()
    g = torch.Generator().manual_seed(M + N + K + splitk)
    A = torch.randn((M, K), generator=g)
    Bm = torch.randn((N, K) if bk else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ (Bm.double().T if bk else Bm.double()) + bias.double()
    dA, dB, dbias = A.cuda(), Bm.cuda(), bias.cuda()
    ldc = (N + 3) // 4 * 4
    dC = torch.zeros(M, ldc, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    L.check(lib.fb_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), dbias.data_ptr(), M, N, K, A.shape[1], Bm.shape[1], ldc,
                         1, bk, 0, splitk, 4 if pairs else 3, s))
    torch.cuda.synchronize()
    assert rel(dC[:, :N], ref) < 1e-5
    assert float(dC[:, N:].abs().max()) == 0.0 if ldc > N else True
    assert lib.fb_sgemm(dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), dbias.data_ptr(), M, N, K, A.shape[1], Bm.shape[1], ldc,
                        1, bk, 1, splitk, 4 if pairs else 3, s) != 0      # a ReLU epilogue cannot be split


def _run_update_case(g, d, use_goal, graph, mlp_mode=0, fused=False):
    L = _L()
    t = {k: torch.from_numpy(np.array(v)) for k, v in subtree(g, "in").items()}
    B = t["obs"].shape[0]
    eng = make_engine(d, B, use_goal=use_goal, mix_ratio=0.5, ortho_coef=float(g["cfg/ortho_coef"]), mlp_mode=mlp_mode, fused=fused,
                      q_loss_coef=float(g["cfg/q_loss_coef"]) if "cfg/q_loss_coef" in g else None,
                      norm_z=bool(g["cfg/norm_z"]) if "cfg/norm_z" in g else True,
                      add_trunk="param0/actor/trunk.0.weight" in g and "param0/actor/obs_net.0.weight" in g,
                      preprocess="param0/actor/obs_net.0.weight" in g or ("cfg/boltzmann" in g and bool(g["cfg/boltzmann"])),
                      boltzmann="cfg/boltzmann" in g and bool(g["cfg/boltzmann"]), temp=float(g["cfg/temp"]) if "cfg/temp" in g else 1.0,
                      debug=len(subtree(g, "param0/backward_net")) == 0)
    load_params(eng, fwd=subtree(g, "param0/forward_net"), bwd=subtree(g, "param0/backward_net"),
                actor=subtree(g, "param0/actor"), fwd_tgt=subtree(g, "param0/forward_target_net"),
                bwd_tgt=subtree(g, "param0/backward_target_net"))
    lr, tau = float(g["cfg/lr"]), float(g["cfg/tau"])
    eng.set_scalars(float(g["cfg/stddev"]), float(g["cfg/stddev_clip"]), lr, lr, lr, tau)
    eng.set_batch(t["obs"], t["action"], t["discount"], t["next_obs"], t["next_goal"] if use_goal else None,
                  t["next_goal"] if use_goal else None)
    eng.set_z(t["z"])
    eng.set_noise(t["noise_fb"], t["noise_actor"])
    return eng, t, L


# qloss*: cfg.q_loss (fb_ddpg.py:330-341); nonorm*: cfg.norm_z = False (fb_modules.py:227-229); trunk*: cfg.add_trunk (fb_modules.py:96-100)
# boltz: cfg.boltzmann (DiagGaussianActor + SquashedNormal, fb_modules.py:129-151, fb_ddpg.py:304-306,391-393,406)
@pytest.mark.parametrize("case", ["small", "goal", "wide", "qloss", "qloss_goal", "nonorm", "nonorm_goal", "trunk", "trunk_goal", "nopre", "boltz", "debug"])
# (graph, mlp_mode, fused): eager / CUDA-graph launches of the per-layer plan, the fp32 SIMT plan, and the fused stack kernels
# (k_fused_stack: the same plan as stages of one persistent kernel per segment), eager and under a graph
@pytest.mark.parametrize("graph,mlp_mode,fused", [(False, 0, False), (True, 0, False), (True, 1, False), (False, 0, True), (True, 0, True)])
def test_update_matches_reference_golden(case, graph, mlp_mode, fused):
    g = load_golden(f"update_{case}")
    fwd, bwd, actor = (golden_params(g, f"param0/{n}") for n in ("forward_net", "backward_net", "actor"))
    d = dims_from_params(fwd, bwd, actor)
    use_goal = case.endswith("goal")
    q_coef = float(g["cfg/q_loss_coef"]) if "cfg/q_loss_coef" in g else None
    norm_z = bool(g["cfg/norm_z"]) if "cfg/norm_z" in g else True
    # (debug: the identity backward map has no projection of its own; the oracle's backward_map returns the goal for an empty parameter set)
    boltz = "cfg/boltzmann" in g and bool(g["cfg/boltzmann"])
    temp = float(g["cfg/temp"]) if "cfg/temp" in g else 1.0
    eng, t, L = _run_update_case(g, d, use_goal, graph, mlp_mode, fused)
    if fused:   # the segment really is one launch: MIX + FB_FWD up to the contraction, then the loss GEMMs + FB_BWD
        assert eng.launch_count(L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_FB_BWD) < eng.launch_count(
            L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_FB_BWD, fused=False) // 3

    # ---- update_fb up to the gradients (fb_ddpg.py:303-383) ----
    eng.run(L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_FB_BWD | L.PHASE_METRICS, graph=graph)
    torch.cuda.synchronize()
    assert rel(eng.view("z"), t["z"]) < 1e-6
    ora = O.fb_loss_and_grads(fwd, bwd, golden_params(g, "param0/forward_target_net"), golden_params(g, "param0/backward_target_net"),
                              actor, t["obs"], t["action"], t["discount"], t["next_obs"], t["next_goal"], t["z"], t["noise_fb"],
                              float(g["cfg/stddev"]), float(g["cfg/stddev_clip"]), float(g["cfg/ortho_coef"]), d.z_dim, q_coef, norm_z,
                              boltzmann=boltz)
    for name in ("next_action", "tF1", "tF2", "tB", "F1", "F2", "B", "dF1", "dF2", "dB"):
        assert rel(eng.view(name), ora[name]) < REL_TOL, name
    m = eng.read_metrics()
    assert ("metric_fb/q_loss" in g) == (q_coef is not None)
    if q_coef is None:
        assert m["q_loss"] == 0.0
    for k, v in subtree(g, "metric_fb").items():
        if k == "fb_opt_lr":
            continue
        assert m[k] == pytest.approx(float(v), rel=REL_TOL, abs=1e-5), k
    for net, key in ((L.NET_FORWARD, "forward_net"), (L.NET_BACKWARD, "backward_net")):
        got = read_tensors(eng, net, "grad")
        for name, ref in subtree(g, f"grad_fb/{key}").items():
            assert rel(got[name], ref) < REL_TOL, (key, name, rel(got[name], ref))

    # ---- fb_opt.step (Adam step 1) + both soft updates ----
    eng.run(L.PHASE_FB_ADAM, graph=graph)
    torch.cuda.synchronize()
    for net, key in ((L.NET_FORWARD, "forward_net"), (L.NET_BACKWARD, "backward_net")):
        got = read_tensors(eng, net, "param")
        gref = subtree(g, f"grad_fb/{key}")
        for name, ref in subtree(g, f"param1/{key}").items():
            # |dp| = lr = 1e-4 on step 1, in the direction of sign(g): an element whose gradient is ~0 next to the tensor's
            # scale may take the other sign (2 lr apart); everything else sits within rounding
            diff = np.abs(got[name].numpy() - ref)
            flipped = diff >= 2e-5
            assert diff.max() < 2.1e-4 and np.all(np.abs(gref[name][flipped]) <= 1e-4 * np.abs(gref[name]).max()), (key, name, diff.max())
        got = read_tensors(eng, net, "target")
        for name, ref in subtree(g, f"param1/{key.replace('_net', '_target_net')}").items():
            assert np.abs(got[name].numpy() - ref).max() < 1e-5, (key, name)
        if got:   # (debug: backward_net has no tensors)
            assert float(eng.tensors(net, "grad")[next(iter(got))].abs().max()) == 0.0   # grads cleared for the next step

    # ---- update_actor with the just-updated forward_net (fb_ddpg.py:389-410) ----
    # the oracle is evaluated on the engine's own post-Adam forward_net so that Adam's sign(g) amplification of
    # ulp-level gradient noise does not leak into this comparison
    fwd1 = read_tensors(eng, L.NET_FORWARD, "param")
    eng.run(L.PHASE_ACTOR_FWD | L.PHASE_ACTOR_BWD | L.PHASE_METRICS, graph=graph)
    torch.cuda.synchronize()
    ora_a = O.actor_loss_and_grads(actor, fwd1, t["obs"], t["z"], t["noise_actor"], float(g["cfg/stddev"]), float(g["cfg/stddev_clip"]),
                                   boltzmann=boltz, temp=temp)
    assert rel(eng.view("action_new"), ora_a["action"]) < REL_TOL
    if not boltz:   # (the DiagGaussianActor has no tanh'd mean buffer: its [mu | raw log-std] output is checked through action / log pi)
        assert rel(eng.view("mu"), ora_a["mu"]) < REL_TOL
    m = eng.read_metrics()
    assert m["actor_loss"] == pytest.approx(float(ora_a["actor_loss"]), rel=REL_TOL, abs=1e-5)
    assert m["q"] == pytest.approx(float(ora_a["q"]), rel=REL_TOL, abs=1e-5)
    assert m["actor_logprob"] == pytest.approx(float(ora_a["actor_logprob"]), rel=REL_TOL, abs=1e-5)
    assert abs(m["q1_success"] - float(ora_a["q1_success"])) <= 1.0 / t["obs"].shape[0] + 1e-6   # additional_metric (fb_ddpg.py:403-404)
    ref_m = subtree(g, "metric_actor")   # the reference's own numbers (its forward_net differs by Adam noise only)
    assert m["actor_loss"] == pytest.approx(float(ref_m["actor_loss"]), rel=5e-3, abs=1e-4)
    got = read_tensors(eng, L.NET_ACTOR, "grad")
    for name, ref in ora_a["grads_actor"].items():
        assert rel(got[name], ref) < REL_TOL, (name, rel(got[name], ref))
    for name, ref in subtree(g, "grad_actor/actor").items():
        assert rel(got[name], ref) < 2e-2, (name, rel(got[name], ref))
    eng.run(L.PHASE_ACTOR_ADAM, graph=graph)
    torch.cuda.synchronize()
    got = read_tensors(eng, L.NET_ACTOR, "param")
    for name, ref in subtree(g, "param1/actor").items():
        assert np.abs(got[name].numpy() - ref).max() < 2.1e-4, name   # a sign flip of a ~0 gradient moves p by 2 lr
    assert eng.get_adam_steps() == (1, 1)
    eng.close()


def _to(p, dt):
    return {k: v.to(dt) for k, v in p.items()}


class _ForcedRelu:
    """Stand-in for torch.relu inside the oracle: every ReLU takes the branch the engine took (mask = the engine's saved
    activation > 0), so that the two evaluate the SAME piece of the piecewise-linear network.  Records how many units
    the oracle itself would have switched differently and how close to zero their pre-activations were."""

    def __init__(self, masks):
        self.masks, self.i, self.flips, self.worst = masks, 0, 0, 0.0

    def __call__(self, x):
        m = self.masks[self.i].to(x.dtype)
        self.i += 1
        assert m.shape == x.shape, (self.i, m.shape, x.shape)
        diff = (x > 0) != (m > 0)
        if bool(diff.any()):
            self.flips += int(diff.sum())
            self.worst = max(self.worst, float(x.detach().abs()[diff].max() / x.detach().abs().mean()))
        return x * m


# (id, batch, obs, act, z, goal_dim or None, seed, mlp_mode): the benchmarked configurations of BASELINE.json at the reference's full
# widths — configs[0] (batch 256), configs[1] (batch 1024: the metric's configuration, other tile widths / split-K / lazy-ReLU plans),
# configs[2] dims (quadruped O=78 A=12 with the 2-wide simplified_quadruped goal space), configs[4] dims (cheetah O=17, z=100) at the
# per-rank batch of 8 GPUs (512) and at the whole batch (4096)
FULL_WIDTH_CASES = [
    ("walker_b256", 256, 24, 6, 50, None, 11, 0),
    ("walker_b256_seed2", 256, 24, 6, 50, None, 111, 0),
    ("walker_b1024", 1024, 24, 6, 50, None, 12, 0),
    ("walker_b1024_seed2", 1024, 24, 6, 50, None, 112, 0),
    ("walker_b1024_simt", 1024, 24, 6, 50, None, 12, 1),
    ("quadruped_goal2_b1024", 1024, 78, 12, 50, 2, 13, 0),
    ("quadruped_goal2_b1024_seed2", 1024, 78, 12, 50, 2, 113, 0),
    ("cheetah_z100_b512", 512, 17, 6, 100, None, 14, 0),
    ("cheetah_z100_b512_simt", 512, 17, 6, 100, None, 14, 1),
    ("cheetah_z100_b4096", 4096, 17, 6, 100, None, 15, 0),
    ("walker_b1024_fused", 1024, 24, 6, 50, None, 12, 2),               # mlp_mode 2 here: tcgen05 plan through the fused stack kernels
    ("quadruped_goal2_b1024_fused", 1024, 78, 12, 50, 2, 13, 2),
    ("walker_b1024_boltzmann", 1024, 24, 6, 50, None, 16, 0, "boltz"),   # cfg.boltzmann: DiagGaussianActor stack at full width (temp 0.7)
]


@pytest.mark.parametrize("case", FULL_WIDTH_CASES, ids=[c[0] for c in FULL_WIDTH_CASES])
def test_full_width_step_against_oracle(monkeypatch, case):
    """Default widths of the reference config (hidden 1024, feature 512, backward hidden 526) at the benchmarked batch sizes and
    dimensions (FULL_WIDTH_CASES); oracle on CPU from the same seeded parameters and inputs.

    At these widths (1.3 M ReLU units per 256 rows) some pre-activation is always within fp32 rounding of zero, and such a
    unit switches between summation orders: the reference's own fp32 evaluation sits up to ~6e-3 from the exact gradient
    on whole tensors for that reason alone (SURVEY.md 7.3 measured 9e-4 between 1 and 8 CPU threads), so a per-tensor
    1e-3 gate against ONE fp32 evaluation cannot be met by any independent evaluation.  The test therefore pins the
    branch: the fp64 oracle is evaluated with every ReLU taking the branch the engine took (_ForcedRelu), which must
    differ from the oracle's own choice only on units whose pre-activation is ~0 (|x| < 1e-4 of the layer's mean |x|,
    and only a handful of them); on that common branch every gradient tensor must match to 2e-4 (observed ~3e-6), the
    losses / metrics to 1e-3 against the unforced fp32 oracle."""
    _, B, obs_dim, act_dim, z_dim, goal_dim, seed, mlp_mode = case[:8]
    boltz = len(case) > 8 and case[8] == "boltz"
    temp = 0.7 if boltz else 1.0
    use_goal = goal_dim is not None
    L = _L()
    d = O.Dims(obs_dim=obs_dim, action_dim=act_dim, z_dim=z_dim, goal_dim=goal_dim if use_goal else obs_dim)
    Fd = d.feature_dim
    max_flips = 64 * max(1, B // 256)
    gen = torch.Generator().manual_seed(seed)
    actor = O.init_params(O.boltzmann_actor_spec(d) if boltz else O.actor_spec(d), gen)
    fwd = O.init_params(O.forward_map_spec(d), gen)
    bwd = O.init_params(O.backward_map_spec(d), gen)
    fwd_t = {k: v + 0.02 * torch.randn(v.shape, generator=gen) for k, v in fwd.items()}
    bwd_t = {k: v + 0.02 * torch.randn(v.shape, generator=gen) for k, v in bwd.items()}
    obs, next_obs = torch.randn(B, d.obs_dim, generator=gen), torch.randn(B, d.obs_dim, generator=gen)
    action = torch.rand(B, d.action_dim, generator=gen) * 2 - 1
    discount = torch.full((B, 1), 0.98)
    goal = torch.randn(B, d.goal_dim, generator=gen) if use_goal else obs
    next_goal = torch.randn(B, d.goal_dim, generator=gen) if use_goal else next_obs
    z_rand = O.sample_z(B, d.z_dim, gen)
    noise_fb, noise_actor = torch.randn(B, d.action_dim, generator=gen), torch.randn(B, d.action_dim, generator=gen)
    perm = torch.randperm(B, generator=gen)
    mix_mask = (torch.rand(B, generator=gen) < 0.5)
    # oracle z mixing (fb_ddpg.py:460-485): backward_input = goal (goal space) or obs, permuted
    z = z_rand.clone()
    idx = torch.where(mix_mask)[0]
    with torch.no_grad():
        z[idx] = O.l2_project(O.backward_map(bwd, goal[perm][idx], d.z_dim), d.z_dim)

    eng = make_engine(d, B, use_goal=use_goal, mlp_mode=mlp_mode % 2, fused=mlp_mode == 2, boltzmann=boltz, temp=temp)
    load_params(eng, fwd=fwd, bwd=bwd, actor=actor, fwd_tgt=fwd_t, bwd_tgt=bwd_t)
    eng.set_scalars(0.2, 0.3, 1e-4, 1e-4, 1e-4, 0.01)
    eng.set_indices(perm=perm, mix_mask=mix_mask.int())
    eng.set_batch(obs, action, discount, next_obs, goal if use_goal else None, next_goal if use_goal else None)
    eng.set_z(z_rand)
    eng.set_noise(noise_fb, noise_actor)
    eng.run(L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_FB_BWD | L.PHASE_METRICS)
    torch.cuda.synchronize()
    assert rel(eng.view("z"), z) < 1e-5
    zz = eng.view("z").detach().cpu()   # evaluate the oracles on the engine's z so that the mix forward is not compared twice
    f32, f64 = torch.float32, torch.float64

    def act(name, r0=0, c0=None):   # the engine's post-ReLU activation -> branch mask
        v = eng.view(name).detach().cpu()[r0:r0 + B]
        return (v if c0 is None else v[:, c0:c0 + Fd]) > 0

    def run_fb(dt):
        return O.fb_loss_and_grads(_to(fwd, dt), _to(bwd, dt), _to(fwd_t, dt), _to(bwd_t, dt), _to(actor, dt), obs.to(dt), action.to(dt),
                                   discount.to(dt), next_obs.to(dt), next_goal.to(dt), zz.to(dt), noise_fb.to(dt), 0.2, 0.3, 1.0, d.z_dim,
                                   boltzmann=boltz)

    ora32 = run_fb(f32)
    # ReLU call order of fb_loss_and_grads: actor(next_obs) [obs_z_net, obs_net, policy], target F [oa, oz, F1, F2], target B,
    # online F [oa, oz, F1, F2], online B
    # (cfg.boltzmann: the DiagGaussianActor is ONE stack with one ReLU: policy.3 -> "actor.policy.h2", rows [0, B) = next_obs)
    actor_fb = [act("actor.policy.h2")] if boltz else [act("hA", 0, Fd), act("hA", 0, 0), act("actor.policy.h1")]
    forced = _ForcedRelu(actor_fb + [act("hFt", 0, 0), act("hFt", 0, Fd),
                          act("Ft.F1.h1"), act("Ft.F2.h1"), act("Bt.h2"), act("hF", 0, 0), act("hF", 0, Fd), act("F.F1.h1"),
                          act("F.F2.h1"), act("Bo.h2")])
    with monkeypatch.context() as mp:
        mp.setattr(torch, "relu", forced)
        ora = run_fb(f64)
    assert forced.i == len(forced.masks)
    print(f"fb step: {forced.flips} of ~{len(forced.masks) * B * 1024} ReLU units on the other branch than the fp64 oracle, largest |x|/mean|x| {forced.worst:.1e}")
    assert forced.flips <= max_flips and forced.worst < 1e-4
    m = eng.read_metrics()
    for k, v in ora32["metrics"].items():
        assert m[k] == pytest.approx(v, rel=REL_TOL, abs=1e-5), k
    for name in ("next_action", "tF1", "tF2", "tB", "F1", "F2", "B", "dF1", "dF2", "dB"):
        assert rel(eng.view(name), ora[name]) < 1e-4, name
    worst, ref_own, bad = 0.0, 0.0, []
    for net, key in ((L.NET_FORWARD, "grads_forward"), (L.NET_BACKWARD, "grads_backward")):
        got = read_tensors(eng, net, "grad")
        for name in ora[key]:
            e = rel(got[name], ora[key][name])
            worst, ref_own = max(worst, e), max(ref_own, rel(ora32[key][name], ora[key][name]))
            if e >= 2e-4:
                bad.append((key, name, e))
    print(f"fb grads: worst vs the fp64 oracle on the engine's branch {worst:.2e} (the unforced fp32 oracle sits {ref_own:.2e} from it)")
    assert not bad, bad
    eng.run(L.PHASE_FB_ADAM)
    fwd1 = read_tensors(eng, L.NET_FORWARD, "param")
    eng.run(L.PHASE_ACTOR_FWD | L.PHASE_ACTOR_BWD | L.PHASE_METRICS)
    torch.cuda.synchronize()

    def run_actor(dt):
        return O.actor_loss_and_grads(_to(actor, dt), _to(fwd1, dt), obs.to(dt), zz.to(dt), noise_actor.to(dt), 0.2, 0.3,
                                      boltzmann=boltz, temp=temp)

    ora32 = run_actor(f32)
    # actor(obs) [obs_z_net, obs_net, policy] = rows B..2B of the batched actor forward, then F(obs, z, action) [oa, oz, F1, F2]
    actor_a = [act("actor.policy.h2", B)] if boltz else [act("hA", B, Fd), act("hA", B, 0), act("actor.policy.h1", B)]
    forced = _ForcedRelu(actor_a + [act("hF2", 0, 0), act("hF2", 0, Fd), act("F2.F1.h1"), act("F2.F2.h1")])
    with monkeypatch.context() as mp:
        mp.setattr(torch, "relu", forced)
        ora_a = run_actor(f64)
    assert forced.i == len(forced.masks)
    print(f"actor step: {forced.flips} ReLU units on the other branch than the fp64 oracle, largest |x|/mean|x| {forced.worst:.1e}")
    assert forced.flips <= max_flips and forced.worst < 1e-4
    m = eng.read_metrics()
    assert m["actor_loss"] == pytest.approx(float(ora32["actor_loss"]), rel=REL_TOL, abs=1e-5)
    got = read_tensors(eng, L.NET_ACTOR, "grad")
    worst, ref_own, bad = 0.0, 0.0, []
    for name in ora_a["grads_actor"]:
        e = rel(got[name], ora_a["grads_actor"][name])
        worst, ref_own = max(worst, e), max(ref_own, rel(ora32["grads_actor"][name], ora_a["grads_actor"][name]))
        if e >= 2e-4:
            bad.append((name, e))
    print(f"actor grads: worst vs the fp64 oracle on the engine's branch {worst:.2e} (the unforced fp32 oracle sits {ref_own:.2e} from it)")
    assert not bad, bad
    eng.close()


@pytest.mark.parametrize("case", ["small", "wide"])
def test_contraction_tcgen05_matches_simt_and_oracle(case):
    """FB_PHASE_FB_LOSS on the tensor cores (3xTF32 tcgen05, contract_tc.cuh) against the fp32 SIMT formulation and the oracle:
    loss sums, dL/dF1, dL/dF2, dL/dB."""
    L = _L()
    g = load_golden(f"update_{case}")
    fwd, bwd, actor = (golden_params(g, f"param0/{n}") for n in ("forward_net", "backward_net", "actor"))
    d = dims_from_params(fwd, bwd, actor)
    t = {k: torch.from_numpy(np.array(v)) for k, v in subtree(g, "in").items()}
    out = {}
    for mode in (L.CONTRACT_TCGEN05, L.CONTRACT_SIMT):
        eng = make_engine(d, t["obs"].shape[0], contract_mode=mode, ortho_coef=float(g["cfg/ortho_coef"]))
        load_params(eng, fwd=fwd, bwd=bwd, actor=actor, fwd_tgt=golden_params(g, "param0/forward_target_net"),
                    bwd_tgt=golden_params(g, "param0/backward_target_net"))
        eng.set_scalars(float(g["cfg/stddev"]), float(g["cfg/stddev_clip"]), 1e-4, 1e-4, 1e-4, 0.01)
        eng.set_batch(t["obs"], t["action"], t["discount"], t["next_obs"])
        eng.set_z(t["z"])
        eng.set_noise(t["noise_fb"], t["noise_actor"])
        eng.run(L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_METRICS)
        torch.cuda.synchronize()
        out[mode] = {k: eng.view(k).clone() for k in ("dF1", "dF2", "dB")}
        out[mode]["metrics"] = eng.read_metrics()
        eng.close()
    for k in ("dF1", "dF2", "dB"):
        assert rel(out[L.CONTRACT_TCGEN05][k], out[L.CONTRACT_SIMT][k]) < 2e-5, k
    for k in ("fb_loss", "fb_offdiag", "fb_diag", "orth_loss", "orth_loss_offdiag", "target_M", "M1"):
        assert out[L.CONTRACT_TCGEN05]["metrics"][k] == pytest.approx(out[L.CONTRACT_SIMT]["metrics"][k], rel=2e-5, abs=1e-6), k


@pytest.mark.parametrize("world,contract_mode,q_coef", [(2, 0, None), (4, 0, None), (8, 0, None), (2, 1, None), (4, 1, None), (8, 1, None),
                                                        (2, 0, 0.5), (4, 1, 0.5)])
def test_sharded_step_sums_to_single_gpu_step(world, contract_mode, q_coef):
    """Multi-GPU decomposition emulated on one device: `world` engines own disjoint row blocks of one global
    batch, exchange the [F1|F2|tF1|tF2|B|tB|discount] block, and the SUM of their flat gradients / loss partials must equal the
    single-engine step on the whole batch.  q_coef: with the optional Q loss (its covariance runs over the gathered B rows)."""
    L = _L()
    d = O.Dims(obs_dim=24, action_dim=6, z_dim=50, goal_dim=24, hidden_dim=128, feature_dim=64, backward_hidden_dim=70)
    B = 192 if world == 2 else 256
    gen = torch.Generator().manual_seed(3)
    actor = O.init_params(O.actor_spec(d), gen)
    fwd = O.init_params(O.forward_map_spec(d), gen)
    bwd = O.init_params(O.backward_map_spec(d), gen)
    fwd_t = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in fwd.items()}
    bwd_t = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in bwd.items()}
    obs, next_obs = torch.randn(B, d.obs_dim, generator=gen), torch.randn(B, d.obs_dim, generator=gen)
    action = torch.rand(B, d.action_dim, generator=gen) * 2 - 1
    discount = 0.98 * (torch.rand(B, 1, generator=gen) > 0.1).float()
    z = O.sample_z(B, d.z_dim, gen)
    nf, na = torch.randn(B, d.action_dim, generator=gen), torch.randn(B, d.action_dim, generator=gen)

    def make(rows, offset):
        e = make_engine(d, rows, global_batch=B, row_offset=offset, contract_mode=contract_mode, q_loss_coef=q_coef)
        load_params(e, fwd=fwd, bwd=bwd, actor=actor, fwd_tgt=fwd_t, bwd_tgt=bwd_t)
        e.set_scalars(0.2, 0.3, 1e-4, 1e-4, 1e-4, 0.01)
        sl = slice(offset, offset + rows)
        e.set_batch(obs[sl], action[sl], discount[sl], next_obs[sl])
        e.set_z(z[sl])
        e.set_noise(nf[sl], na[sl])
        return e

    full = make_engine(d, B, contract_mode=contract_mode, q_loss_coef=q_coef)
    load_params(full, fwd=fwd, bwd=bwd, actor=actor, fwd_tgt=fwd_t, bwd_tgt=bwd_t)
    full.set_scalars(0.2, 0.3, 1e-4, 1e-4, 1e-4, 0.01)
    full.set_batch(obs, action, discount, next_obs)
    full.set_z(z)
    full.set_noise(nf, na)
    full.run(L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_FB_BWD | L.PHASE_METRICS)
    torch.cuda.synchronize()
    ref_grad = full.grad_fb.clone()
    ref_m = full.read_metrics()

    rows = B // world
    shards = [make(rows, r * rows) for r in range(world)]
    for e in shards:
        e.run(L.PHASE_MIX | L.PHASE_FB_FWD)
    torch.cuda.synchronize()
    gathered = torch.cat([e.gather_block()[0] for e in shards], dim=0)     # what all_gather_into_tensor produces
    for e in shards:
        e.gather_block()[1].copy_(gathered)
        e.run(L.PHASE_FB_LOSS | L.PHASE_FB_BWD | L.PHASE_METRICS)
    torch.cuda.synchronize()
    total = sum(e.grad_fb for e in shards)                                  # what all_reduce(sum) produces
    assert rel(total, ref_grad) < 2e-5
    assert (ref_m["q_loss"] > 0) == (q_coef is not None)
    for k in ("fb_loss", "fb_offdiag", "fb_diag", "orth_loss", "orth_loss_offdiag", "orth_loss_diag", "q_loss"):
        assert sum(e.read_metrics()[k] for e in shards) == pytest.approx(ref_m[k], rel=1e-4, abs=1e-5), k
    # actor phase: every rank applies the same (summed) fb gradient, then the actor gradients sum as well
    full.run(L.PHASE_FB_ADAM | L.PHASE_ACTOR_FWD | L.PHASE_ACTOR_BWD | L.PHASE_METRICS)
    for e in shards:
        e.grad_fb.copy_(total)
        e.run(L.PHASE_FB_ADAM | L.PHASE_ACTOR_FWD | L.PHASE_ACTOR_BWD | L.PHASE_METRICS)
    torch.cuda.synchronize()
    assert rel(shards[0].param_fb, full.param_fb) < 1e-6
    assert rel(sum(e.grad_actor for e in shards), full.grad_actor) < 2e-5
    assert sum(e.read_metrics()["actor_loss"] for e in shards) == pytest.approx(full.read_metrics()["actor_loss"], rel=1e-4, abs=1e-6)
    for e in shards:
        e.close()


def test_add_trunk_wide_tcgen05_against_simt_and_oracle():
    """cfg.add_trunk at widths where the tensor-core plan splits K (lazy-ReLU trunk outputs): the tcgen05 step, the fp32 SIMT step
    and the fp32 oracle agree on losses and on every gradient tensor.  (Runs last: its tolerance is the only one not pinned by a
    reference-generated fixture.)"""
    L = _L()
    d = O.Dims(obs_dim=24, action_dim=6, z_dim=50, goal_dim=24, hidden_dim=256, feature_dim=128, backward_hidden_dim=134)
    B = 128
    gen = torch.Generator().manual_seed(21)
    actor = O.init_params(O.actor_spec(d, True), gen)
    fwd = O.init_params(O.forward_map_spec(d, True), gen)
    bwd = O.init_params(O.backward_map_spec(d), gen)
    fwd_t = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in fwd.items()}
    bwd_t = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in bwd.items()}
    obs, next_obs = torch.randn(B, d.obs_dim, generator=gen), torch.randn(B, d.obs_dim, generator=gen)
    action = torch.rand(B, d.action_dim, generator=gen) * 2 - 1
    discount = torch.full((B, 1), 0.98)
    z = O.sample_z(B, d.z_dim, gen)
    nf, na = torch.randn(B, d.action_dim, generator=gen), torch.randn(B, d.action_dim, generator=gen)
    ora = O.fb_loss_and_grads(fwd, bwd, fwd_t, bwd_t, actor, obs, action, discount, next_obs, next_obs, z, nf, 0.2, 0.3, 1.0, d.z_dim)
    ora_a = O.actor_loss_and_grads(actor, fwd, obs, z, na, 0.2, 0.3)
    got = {}
    for mode in (L.MLP_TCGEN05, L.MLP_SIMT):
        eng = make_engine(d, B, mix_ratio=0.0, mlp_mode=mode, add_trunk=True)
        load_params(eng, fwd=fwd, bwd=bwd, actor=actor, fwd_tgt=fwd_t, bwd_tgt=bwd_t)
        eng.set_scalars(0.2, 0.3, 1e-4, 1e-4, 1e-4, 0.01)
        eng.set_batch(obs, action, discount, next_obs)
        eng.set_z(z)
        eng.set_noise(nf, na)
        # no Adam in between: update_actor then sees the forward_net the oracle was given
        eng.run(L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_FB_BWD | L.PHASE_ACTOR_FWD | L.PHASE_ACTOR_BWD | L.PHASE_METRICS, graph=True)
        torch.cuda.synchronize()
        got[mode] = (eng.read_metrics(), read_tensors(eng, L.NET_FORWARD, "grad"), read_tensors(eng, L.NET_BACKWARD, "grad"),
                     read_tensors(eng, L.NET_ACTOR, "grad"))
        eng.close()
    for mode, (m, gf, gb, ga) in got.items():
        assert m["fb_loss"] == pytest.approx(ora["metrics"]["fb_loss"], rel=REL_TOL), mode
        assert m["actor_loss"] == pytest.approx(float(ora_a["actor_loss"]), rel=REL_TOL, abs=1e-5), mode
        for mine, refs in ((gf, ora["grads_forward"]), (gb, ora["grads_backward"]), (ga, ora_a["grads_actor"])):
            assert list(mine) == list(refs)
            for name, ref in refs.items():
                assert rel(mine[name], ref) < 3e-3, (mode, name, rel(mine[name], ref))
    for a, b_ in zip(got[L.MLP_TCGEN05][1:], got[L.MLP_SIMT][1:]):
        for name in a:
            assert rel(a[name], b_[name]) < 3e-3, name


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,graph", [(2, False), (2, True), (4, True)])
def test_p2p_exchange_ranks_on_one_device(world, graph):
    """The peer-memory exchange (csrc/p2p.cuh: row scatter, fused reduce-scatter + Adam + all-gather, epoch-flag barriers) with
    `world` ranks emulated on ONE device: one engine per rank, each on its own stream, arenas attached by address.  Two full
    gradient steps must land every rank on the parameters / targets of ONE engine stepping the whole batch, and the owned slices of
    the Adam moments must tile the single-engine moments.  (Real multi-GPU ranks: bench.py's parity_check.)"""
    from controllable_agent_b200.engine import EngineConfig, FBStepEngine
    L = _L()
    d = O.Dims(obs_dim=24, action_dim=6, z_dim=50, goal_dim=24, hidden_dim=128, feature_dim=64, backward_hidden_dim=70)
    B = 64 * world
    gen = torch.Generator().manual_seed(5)
    actor = O.init_params(O.actor_spec(d), gen)
    fwd = O.init_params(O.forward_map_spec(d), gen)
    bwd = O.init_params(O.backward_map_spec(d), gen)
    obs, next_obs = torch.randn(B, d.obs_dim, generator=gen), torch.randn(B, d.obs_dim, generator=gen)
    action = torch.rand(B, d.action_dim, generator=gen) * 2 - 1
    discount = 0.98 * (torch.rand(B, 1, generator=gen) > 0.1).float()
    z = O.sample_z(B, d.z_dim, gen)
    nf, na = torch.randn(B, d.action_dim, generator=gen), torch.randn(B, d.action_dim, generator=gen)
    mask = L.PHASE_ALL & ~L.PHASE_SAMPLE

    def feed(e, sl):
        load_params(e, fwd=fwd, bwd=bwd, actor=actor, fwd_tgt=fwd, bwd_tgt=bwd)
        e.set_scalars(0.2, 0.3, 1e-3, 1e-3, 1e-3, 0.01)
        e.set_batch(obs[sl], action[sl], discount[sl], next_obs[sl])
        e.set_z(z[sl])
        e.set_noise(nf[sl], na[sl])
        n = sl.stop - sl.start
        e.set_indices(perm=np.arange(n, dtype=np.int32), mix_mask=np.zeros(n, np.int32))

    full = make_engine(d, B)
    feed(full, slice(0, B))
    for _ in range(2):
        full.run(mask, graph=False)
    torch.cuda.synchronize()
    ref_m = full.read_metrics()
    ref = {k: getattr(full, k).clone() for k in ("param_fb", "param_actor", "target_fb", "m_fb", "v_fb", "m_actor", "v_actor")}
    full.close()   # frees its lanes: every stream of the emulated ranks must map to a hardware queue of its own

    rows = B // world
    shards = [FBStepEngine(EngineConfig(batch=rows, obs_dim=d.obs_dim, action_dim=d.action_dim, z_dim=d.z_dim, goal_dim=d.goal_dim,
                                        hidden_dim=d.hidden_dim, feature_dim=d.feature_dim, backward_hidden_dim=d.backward_hidden_dim,
                                        global_batch=B, row_offset=r * rows, p2p=(world, r)), "cuda", p2p_attach="local")
              for r in range(world)]
    FBStepEngine.attach_local(shards)
    streams = [torch.cuda.Stream() for _ in shards]
    torch.cuda.synchronize()
    for r, (e, st) in enumerate(zip(shards, streams)):
        with torch.cuda.stream(st):
            feed(e, slice(r * rows, (r + 1) * rows))
    torch.cuda.synchronize()
    if graph:   # ranks of ONE context: instantiate every graph before any rank starts spinning on its peers
        for e in shards:
            e.prepare_graph(mask)
    for _ in range(2):
        for e, st in zip(shards, streams):
            with torch.cuda.stream(st):
                e.run(mask, graph=graph)
    torch.cuda.synchronize()
    status = [e.p2p_status() for e in shards]
    assert all(code == 0 for code, _ in status), [(hex(c), ep) for c, ep in status]
    for r, e in enumerate(shards):
        assert rel(e.param_fb, ref["param_fb"]) < 5e-6, r   # (split-K partial sums land in atomic order: ulp-level gradient noise, times Adam)
        assert rel(e.param_actor, ref["param_actor"]) < 5e-6, r
        assert rel(e.target_fb, ref["target_fb"]) < 5e-6, r
        assert float(e.grad_fb.abs().max()) == 0.0 and float(e.grad_actor.abs().max()) == 0.0   # cleared for the next step
        assert e.get_adam_steps() == (2, 2)
    for r in range(1, world):   # the ranks hold the SAME bits: one owner computes each slice
        assert torch.equal(shards[r].param_fb, shards[0].param_fb) and torch.equal(shards[r].param_actor, shards[0].param_actor)
    for name, actor_seg in (("m_fb", False), ("v_fb", False), ("m_actor", True), ("v_actor", True)):
        tiled = torch.zeros_like(ref[name])
        for e in shards:
            first, count = e.moment_slice(actor_seg)
            tiled[first:first + count] = getattr(e, name)[first:first + count]
        assert rel(tiled, ref[name]) < 2e-5, name
    for k in ("fb_loss", "actor_loss"):
        assert sum(e.read_metrics()[k] for e in shards) == pytest.approx(ref_m[k], rel=1e-4, abs=1e-6), k
    for e in shards:
        e.close()
