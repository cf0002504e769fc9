"""world_size-2 tests of the N > 1 path on CPU (gloo): the host-side sharding helpers and the exchange scheme of DESIGN.md
"Multi-GPU" — all-gather of the per-rank [F | B | targets | discount] row blocks, global batch x batch loss evaluated on every
rank with only its own rows live, all-reduce(sum) of the flat gradients — checked against the single-process full-batch
gradient with the oracle's torch-CPU forward (test infrastructure).  The CUDA step implements exactly this decomposition
(tests/test_gpu_step.py::test_sharded_step_sums_to_single_gpu_step checks the kernels against it on one GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fb_oracle as O


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(B, d, dt):
    gen = torch.Generator().manual_seed(7)
    fwd, bwd = O.init_params(O.forward_map_spec(d), gen), O.init_params(O.backward_map_spec(d), gen)
    fwd_t = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in fwd.items()}
    bwd_t = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in bwd.items()}
    c = lambda p: {k: v.to(dt) for k, v in p.items()}  # noqa: E731
    t = dict(obs=torch.randn(B, d.obs_dim, generator=gen), next_obs=torch.randn(B, d.obs_dim, generator=gen),
             action=torch.rand(B, d.action_dim, generator=gen) * 2 - 1, next_action=torch.rand(B, d.action_dim, generator=gen) * 2 - 1,
             discount=0.98 * (torch.rand(B, 1, generator=gen) > 0.1).float(), z=O.sample_z(B, d.z_dim, gen))
    return c(fwd), c(bwd), c(fwd_t), c(bwd_t), {k: v.to(dt) for k, v in t.items()}


def _flat_grads(params):
    return torch.cat([p.grad.reshape(-1) for p in params.values()])


def _worker(rank: int, world: int, port: int, out) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from controllable_agent_b200.dist_utils import reduce_metrics, shard_episodes, shard_layout
        torch.set_num_threads(1)
        dt = torch.float64
        d = O.Dims(obs_dim=9, action_dim=3, z_dim=10, goal_dim=9, hidden_dim=32, feature_dim=16, backward_hidden_dim=22)
        B = 48
        fwd, bwd, fwd_t, bwd_t, t = _inputs(B, d, dt)
        rows, off = shard_layout(B, world, rank)
        sl = slice(off, off + rows)
        f, b = O._with_grad(fwd), O._with_grad(bwd)
        # local rows through the networks (what FB_PHASE_FB_FWD leaves in the rank's exchange block)
        F1, F2 = O.forward_map(f, t["obs"][sl], t["z"][sl], t["action"][sl])
        Bm = O.backward_map(b, t["next_obs"][sl], d.z_dim)
        with torch.no_grad():
            tF1, tF2 = O.forward_map(fwd_t, t["next_obs"][sl], t["z"][sl], t["next_action"][sl])
            tB = O.backward_map(bwd_t, t["next_obs"][sl], d.z_dim)
        block = torch.cat([F1, F2, tF1, tF2, Bm, tB, t["discount"][sl]], dim=1).detach().contiguous()
        glob = torch.empty(B, block.shape[1], dtype=dt)
        dist.all_gather_into_tensor(glob, block)                       # exchange 1: row blocks in rank order
        Z = d.z_dim
        parts = [glob[:, i * Z:(i + 1) * Z].clone() for i in range(6)]
        disc = glob[:, 6 * Z:6 * Z + 1]
        for full, live in zip((parts[0], parts[1], parts[4]), (F1, F2, Bm)):
            full[sl] = live                                              # only this rank's rows carry gradient
        terms = O.fb_loss_terms(parts[0], parts[1], parts[4], parts[2], parts[3], parts[5], disc, 1.0)
        terms["fb_loss"].backward()
        g = torch.cat([_flat_grads(f), _flat_grads(b)])
        dist.all_reduce(g)                                             # exchange 2: gradients only
        m = reduce_metrics({"fb_loss": float(terms["fb_loss"].detach()) / world, "B_norm": float(Bm.detach().norm(dim=-1).mean())}, ["fb_loss", "B_norm"], world, "cpu")
        if rank == 0:
            out.put({"grad": g, "metrics": m, "layout": [shard_layout(B, world, r) for r in range(world)],
                     "episodes": [shard_episodes(11, world, r) for r in range(world)]})
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_reproduces_full_batch_gradient():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # single-process reference on the whole batch
    dt = torch.float64
    d = O.Dims(obs_dim=9, action_dim=3, z_dim=10, goal_dim=9, hidden_dim=32, feature_dim=16, backward_hidden_dim=22)
    fwd, bwd, fwd_t, bwd_t, t = _inputs(48, d, dt)
    f, b = O._with_grad(fwd), O._with_grad(bwd)
    F1, F2 = O.forward_map(f, t["obs"], t["z"], t["action"])
    Bm = O.backward_map(b, t["next_obs"], d.z_dim)
    with torch.no_grad():
        tF1, tF2 = O.forward_map(fwd_t, t["next_obs"], t["z"], t["next_action"])
        tB = O.backward_map(bwd_t, t["next_obs"], d.z_dim)
    terms = O.fb_loss_terms(F1, F2, Bm, tF1, tF2, tB, t["discount"], 1.0)
    terms["fb_loss"].backward()
    ref = torch.cat([_flat_grads(f), _flat_grads(b)])
    assert float((res["grad"] - ref).norm() / ref.norm()) < 1e-10
    assert res["metrics"]["fb_loss"] == pytest.approx(float(terms["fb_loss"].detach()), rel=1e-10)       # partial sums add up
    assert res["metrics"]["B_norm"] == pytest.approx(float(Bm.detach().norm(dim=-1).mean()), rel=1e-10)   # per-rank means average
    assert res["layout"] == [(24, 0), (24, 24)] and res["episodes"] == [(0, 5), (5, 10)]


def test_shard_layout_errors():
    from controllable_agent_b200.dist_utils import shard_layout
    with pytest.raises(ValueError):
        shard_layout(1024, 3, 0)
    with pytest.raises(ValueError):
        shard_layout(1024, 2, 2)
    assert [shard_layout(1024, 8, r) for r in (0, 7)] == [(128, 0), (128, 896)]
