#!/usr/bin/env python
"""bench.py — FB-DDPG gradient-steps/sec (batch=1024, z_dim=50) on N B200s (BASELINE.json metric).

One "step" = one `agent.update(replay, step)` with update_every_steps=1 (train_offline.py:59,118): replay sample,
z draw + mixing, update_fb (+Adam), update_actor (+Adam), both target soft updates.

  python bench.py [--gpus N] [--steps K] [--warmup W]                  our arm (CUDA step, device-resident replay)
  python bench.py --impl reference [--steps K] [--warmup W]            the reference on the host CPU cores, all host threads,
                                                                       rank 0 only: the UNMODIFIED reference files staged in
                                                                       baseline/_ref (kind "reference"), else the oracle port
N > 1 is launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`; the GLOBAL batch stays
1024 (strong scaling, BASELINE.json north_star), each rank steps 1024/N rows against its own replay shard.

Other BASELINE.json configs:  cfg 3: --obs-dim 78 --action-dim 12 --goal-space simplified_quadruped
                              cfg 5: --obs-dim 17 --z-dim 100 --batch 4096 [--gpus 8]
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "FB-DDPG gradient-steps/sec (batch=1024, z_dim=50)"
UNIT = "gradient-steps/s"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
GOAL_DIMS = {"simplified_walker": 3, "walker_pos_speed": 4, "walker_pos_speed_z": 6, "simplified_quadruped": 2, "quad_pos_speed": 7,
             "simplified_jaco": 3, "simplified_point_mass_maze": 2}
PARITY_TOL = 1e-3
# the actor gradient is a sum of per-sample terms that largely cancel (|sum| << sum of norms): ReLU units whose pre-activation is within
# fp32 rounding of zero switch with the summation order (tile shapes differ between a 1024-row and a 1024/N-row plan), and that alone moves
# the whole-tensor actor gradient by up to ~1e-2 relative — the reference itself differs by 9e-4 between 1 and 8 CPU threads;
# the parity tests hold actor gradients against the reference-generated fixtures at 2e-2 for the same reason (tests/test_gpu_step.py)
PARITY_TOL_ACTOR_GRAD = 2e-2


def parse() -> argparse.Namespace:
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=1024)
    p.add_argument("--z-dim", type=int, default=50)
    p.add_argument("--obs-dim", type=int, default=24)
    p.add_argument("--action-dim", type=int, default=6)
    p.add_argument("--goal-space", default=None, help="agent.goal_space (e.g. simplified_quadruped: goal_dim 2); the replay then stores goals")
    p.add_argument("--episodes", type=int, default=5000, help="episodes of the synthetic replay (whole job)")
    p.add_argument("--episode-len", type=int, default=1000)
    p.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer end-to-end leg (0: min(steps, 100))")
    p.add_argument("--cpu-steps", type=int, default=12, help="timed steps of the cpu_baseline leg (rank 0, N=1)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-cuda-eager", action="store_true", help="skip the reference-on-cuda (PyTorch eager) leg")
    p.add_argument("--no-parity-check", action="store_true")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--timeline", default=None, help="file for rank 0's per-launch table of one step (phase, kind, us; collectives marked)")
    p.add_argument("--collectives", default="auto", choices=["auto", "p2p", "graph", "torch"],
                   help="multi-GPU exchange: p2p = the library's own kernels over NVLink peer memory, graph = NCCL calls captured in the step "
                        "graph on the library's communicator, torch = torch.distributed calls between graph segments; auto = the agent's default")
    p.add_argument("--mlp-mode", default="tcgen05", choices=["tcgen05", "simt"], help="wide Linear products: tensor cores (3xTF32) or fp32 CUDA cores")
    return p.parse_args()


def workload(a: argparse.Namespace) -> dict:
    cfgname = "configs[1]" if (a.obs_dim, a.action_dim, a.z_dim, a.batch, a.goal_space) == (24, 6, 50, 1024, None) else "variant"
    goal = f" goal_space={a.goal_space} (goal_dim {GOAL_DIMS.get(a.goal_space, '?')})" if a.goal_space else ""
    return {"workload": f"fb_ddpg offline update (BASELINE.json {cfgname}): obs={a.obs_dim} act={a.action_dim} z={a.z_dim}{goal} "
                        f"batch={a.batch} hidden=1024 feature=512 backward_hidden=526, {a.episodes}x{a.episode_len}-step synthetic "
                        "replay, update_every_steps=1",
            "global_batch": a.batch, "episodes": a.episodes, "episode_len": a.episode_len}


# ---------------------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.rows: list = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ---------------------------------------------------------------------------------------------------------------------
# synthetic replay in host memory, reference `_storage` layout: name -> [E, T+1, dim] fp32 (in_memory_replay_buffer.py:126)
# ---------------------------------------------------------------------------------------------------------------------
def host_storage(a: argparse.Namespace, episodes: int, seed: int, physics_dim: int = 0) -> dict:
    import numpy as np
    rs = np.random.default_rng(seed)
    E, R = episodes, a.episode_len + 1
    st = {"observation": rs.standard_normal((E, R, a.obs_dim), dtype=np.float32),
          "action": rs.random((E, R, a.action_dim), dtype=np.float32) * 2 - 1,
          "reward": rs.random((E, R, 1), dtype=np.float32),
          "discount": np.ones((E, R, 1), np.float32)}
    if a.goal_space:
        st["goal"] = rs.standard_normal((E, R, GOAL_DIMS[a.goal_space]), dtype=np.float32)
    if physics_dim:
        st["physics"] = np.zeros((E, R, physics_dim), np.float32)   # the reference's sample() always reads it (:166)
    return st


# ---------------------------------------------------------------------------------------------------------------------
# the reference on the host CPU / on cuda through PyTorch eager.  oracle/ is test infrastructure: only these legs run it.
# ---------------------------------------------------------------------------------------------------------------------
def reference_staged() -> bool:
    if os.environ.get("FB_BENCH_FORCE_PORT"):   # calibration runs: time the oracle port although the reference is staged
        return False
    return os.path.isfile(os.path.join(REF_DIR, "url_benchmark", "agent", "fb_ddpg.py"))


def _all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is one process on the host cores, so it takes all of them."""
    import torch
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def make_reference_agent(a: argparse.Namespace, device: str, storage: dict):
    """(update callable, kind, description): the UNMODIFIED reference (baseline/_ref, imported under oracle/ref_shim.py's stubs for
    hydra / omegaconf / dm_env / dmc / goals) when staged, else the oracle port (torch-CPU restatement of the same op sequence)."""
    import numpy as np
    import torch
    E = storage["discount"].shape[0]
    if reference_staged():
        os.environ["FB_REFERENCE_ROOT"] = REF_DIR
        from oracle import ref_shim
        R = ref_shim.load()
        torch.manual_seed(1)
        np.random.seed(1)
        cfg = R.FBDDPGAgentConfig(obs_type="states", obs_shape=(a.obs_dim,), action_shape=(a.action_dim,), device=device, use_tb=False,
                                  use_wandb=False, use_hiplog=False, num_expl_steps=0, update_encoder=False, goal_space=a.goal_space,
                                  z_dim=a.z_dim, batch_size=a.batch, update_every_steps=1)
        agent = R.FBDDPGAgent(**dataclasses.asdict(cfg))
        rb = R.ReplayBuffer(max_episodes=E, discount=0.98, future=0.99)
        rb._storage = dict(storage)           # filled directly, as ReplayBuffer.load() fills it (in_memory_replay_buffer.py:192-208)
        rb._episodes_length[:] = a.episode_len
        rb._idx, rb._full = 0, True
        return (lambda i: agent.update(rb, i)), "reference", (
            f"unmodified url_benchmark FBDDPGAgent.update + ReplayBuffer.sample (baseline/_ref under oracle/ref_shim.py), device={device}")
    if device != "cpu":
        raise RuntimeError("baseline/_ref is not staged (run __graft_entry__.build() in the container): no reference for device=cuda")
    from oracle import fb_oracle as O
    torch.manual_seed(1)
    np.random.seed(1)
    gd = GOAL_DIMS[a.goal_space] if a.goal_space else a.obs_dim
    d = O.Dims(obs_dim=a.obs_dim, action_dim=a.action_dim, z_dim=a.z_dim, goal_dim=gd)
    agent = O.OracleAgent(O.OracleConfig(dims=d, batch_size=a.batch, metrics=False, use_goal=bool(a.goal_space)))
    replay = O.OracleReplay(E, 0.98, 0.99)
    for e in range(E):
        replay.add_episode({k: v[e] for k, v in storage.items() if k != "physics"})
    return (lambda i: agent.update(replay, i)), "port", "oracle/fb_oracle.py OracleAgent (the reference's torch-CPU op sequence restated)"


def time_reference(a: argparse.Namespace, device: str, steps: int, warmup: int, budget_s: float, episodes: int) -> dict:
    import torch
    threads = _all_host_threads()
    storage = host_storage(a, episodes, seed=0, physics_dim=18)
    update, kind, what = make_reference_agent(a, device, storage)
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    t0 = time.perf_counter()
    for i in range(max(warmup, 1)):
        update(i)
    sync()
    per = (time.perf_counter() - t0) / max(warmup, 1)
    done = steps
    if per * steps > budget_s:
        done = max(3, int(budget_s / per))
    t0 = time.perf_counter()
    for i in range(done):
        update(i)
    sync()
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": threads, "host_cpus": os.cpu_count(), "kind": kind, "steps": done,
            "ms_per_step": 1e3 * dt / done,
            "sample": f"{done} full agent.update() steps at batch={a.batch} of {what}, {threads} torch threads, on a "
                      f"{episodes}x{a.episode_len}-step synthetic replay in host memory"}


def run_reference(a: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_reference(a, "cpu", a.steps, a.warmup, budget_s=150.0, episodes=a.episodes)
    cfg = workload(a)
    cfg["parallelism"] = f"host CPU, {r['cores']} torch threads (torch.set_num_threads(os.cpu_count()), whatever the launcher exported)"
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": r["steps"],
            "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "host_cpus", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class HostReplay:
    """A replay buffer in HOST memory with the attribute layout and the sample() contract of the reference's
    in_memory_replay_buffer.ReplayBuffer (:66-88, :139-190: numpy index draws and fancy-index gathers returning numpy arrays), filled
    the way ReplayBuffer.load() fills `_storage`.  For the end-to-end leg: every step's batch is gathered on the host and crosses PCIe
    inside the timed region — by the library's own host gather when the agent recognises the layout (the default), or by this
    object's Python sample() (agent.native_host_sampling = False)."""

    def __init__(self, storage: dict, gamma: float, episode_len: int) -> None:
        import numpy as np
        self._storage = storage
        E = storage["observation"].shape[0]
        self._max_episodes, self._discount, self._future = E, gamma, 1.0
        self._episodes_length = np.full(E, episode_len, np.int32)
        self._is_fixed_episode_length, self._episodes_selection_probability = True, None
        self._idx, self._full = 0, True

    def __len__(self) -> int:
        return self._max_episodes

    def sample(self, batch_size: int):
        import numpy as np
        from controllable_agent_b200 import EpisodeBatch
        s = self._storage
        ep = np.random.randint(0, len(self), size=batch_size)
        t = np.random.randint(0, self._episodes_length[ep]) + 1
        goal = s["goal"][ep, t - 1] if "goal" in s else None
        next_goal = s["goal"][ep, t] if "goal" in s else None
        return EpisodeBatch(obs=s["observation"][ep, t - 1], action=s["action"][ep, t], reward=s["reward"][ep, t],
                            discount=s["discount"][ep, t] * self._discount, next_obs=s["observation"][ep, t], goal=goal, next_goal=next_goal)


def parity_check(agent, a: argparse.Namespace, world: int, rank: int, dev) -> dict:
    """Two gradient steps from the current parameter snapshot on uploaded inputs (rows, z, action noise; mixing off), run (1) through
    the N-rank path exactly as the timed region runs it (graph, collectives) and (2) on ONE engine holding the whole global batch;
    compares the Adam first moments (linear in the two steps' gradients), the parameters, the targets and the losses.  At N = 1 the
    second run is the eager (no-graph) launch sequence of the same engine.  Restores the snapshot afterwards."""
    import numpy as np
    import torch
    from controllable_agent_b200 import _lib as L
    from controllable_agent_b200.engine import FBStepEngine
    e = agent.engine
    n, Bl = a.batch, a.batch // world
    flats = ("param_fb", "m_fb", "v_fb", "target_fb", "param_actor", "m_actor", "v_actor")
    snap = {k: getattr(e, k).clone() for k in flats}   # (p2p: dead moment ranges are zeros on both sides)
    steps0 = e.get_adam_steps()
    g = torch.Generator().manual_seed(4242)   # identical on every rank
    G = agent.goal_dim if a.goal_space else 0
    obs, nobs = torch.randn(n, a.obs_dim, generator=g), torch.randn(n, a.obs_dim, generator=g)
    act = torch.rand(n, a.action_dim, generator=g) * 2 - 1
    disc = 0.98 * (torch.rand(n, 1, generator=g) > 0.02).float()
    goal, ngoal = (torch.randn(n, G, generator=g), torch.randn(n, G, generator=g)) if G else (None, None)
    z = torch.randn(n, a.z_dim, generator=g)
    z = (a.z_dim ** 0.5) * z / z.norm(dim=1, keepdim=True)
    nf, na = torch.randn(n, a.action_dim, generator=g), torch.randn(n, a.action_dim, generator=g)
    mask = (L.PHASE_ALL & ~L.PHASE_SAMPLE)

    def feed(eng, sl: slice, rows: int) -> None:
        eng.set_scalars(0.2, 0.3, 1e-4, 1e-4, 1e-4, 0.01)
        eng.set_batch(obs[sl], act[sl], disc[sl], nobs[sl], goal[sl] if G else None, ngoal[sl] if G else None)
        eng.set_z(z[sl])
        eng.set_noise(nf[sl], na[sl])
        eng.set_indices(perm=np.arange(rows, dtype=np.int32), mix_mask=np.zeros(rows, np.int32))

    def run_two(eng, runner) -> dict:
        ms = []
        for _ in range(2):
            runner()
            torch.cuda.synchronize(dev)
            ms.append(eng.read_metrics())
        out = {k: getattr(eng, k).clone() for k in ("param_fb", "param_actor", "target_fb")}
        fm = eng.full_moments()   # p2p exchange: the moment slices live on their owner ranks (all-gathered here: collective)
        out["m_fb"], out["m_actor"] = fm["m_fb"], fm["m_actor"]
        out["metrics"] = ms
        return out

    def restore(eng) -> None:
        for k, v in snap.items():
            getattr(eng, k).copy_(v)
        eng.set_adam_steps(*steps0)

    feed(e, slice(rank * Bl, (rank + 1) * Bl), Bl)
    got = run_two(e, lambda: agent._run(mask))
    if world > 1:
        got["metrics"] = [agent._reduce_metrics(m) for m in got["metrics"]]
    restore(e)
    res: dict = {"world": world, "steps": 2, "tol": PARITY_TOL}
    if rank == 0:
        if world > 1:
            ref_eng = FBStepEngine(dataclasses.replace(e.cfg, batch=n, global_batch=None, row_offset=0, nccl=None, p2p=None), dev)
            restore(ref_eng)
            feed(ref_eng, slice(0, n), n)
            ref = run_two(ref_eng, lambda: ref_eng.run(mask, graph=False))
            res["against"] = "one engine, whole global batch, eager launches"
        else:
            feed(e, slice(0, n), n)
            ref = run_two(e, lambda: e.run(mask, graph=False))
            restore(e)
            res["against"] = "the same engine, eager launches instead of the CUDA graph"

        def rel(x, y) -> float:
            return float((x.double() - y.double()).norm() / y.double().norm().clamp_min(1e-30))
        errs = {k: rel(got[k], ref[k]) for k in ("m_fb", "m_actor", "param_fb", "param_actor", "target_fb")}
        for s_ in range(2):
            for k in ("fb_loss", "actor_loss"):
                r_ = ref["metrics"][s_][k]
                errs[f"{k}[{s_}]"] = abs(got["metrics"][s_][k] - r_) / max(abs(r_), 1e-30)
        res["rel_err"] = {k: float(f"{v:.3e}") for k, v in errs.items()}
        res["max_rel"] = max(v for k, v in errs.items() if k != "m_actor")
        res["max_rel_actor_grad"] = errs["m_actor"]
        res["tol_actor_grad"] = PARITY_TOL_ACTOR_GRAD
        res["ok"] = bool(res["max_rel"] <= PARITY_TOL and errs["m_actor"] <= PARITY_TOL_ACTOR_GRAD) and all(np.isfinite(v) for v in errs.values())
        if world > 1:
            ref_eng.close()
    return res


def profile_traffic(kernel: str, csv_path: str = "profiles/r2_ncu_key_metrics.csv") -> dict | None:
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of one step (dram__bytes_read.sum +
    dram__bytes_write.sum averaged over its launches); bench.py itself cannot read DRAM counters.  None when the file is absent."""
    import csv
    path = os.path.join(ROOT, csv_path)
    try:
        with open(path, newline="") as f:
            rows = list(csv.reader(f))
        col = {name: i for i, name in enumerate(rows[0])}
        rd, wr = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        picked = [r for r in rows[2:] if kernel in r[0]]
        if not picked:
            return None
        total = sum(float(r[rd]) * scale[rows[1][rd]] + float(r[wr]) * scale[rows[1][wr]] for r in picked)
        return {"bytes_per_launch": total / len(picked), "launches": len(picked),
                "source": f"{csv_path} (ncu --set full of one step at an earlier commit of this round, not this run: dram__bytes_read.sum + "
                          f"dram__bytes_write.sum averaged over the step's {len(picked)} {kernel} launches)"}
    except (OSError, KeyError, ValueError, IndexError):
        return None


def run_ours(a: argparse.Namespace) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run --nproc-per-node N (see module docstring)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from controllable_agent_b200 import FBDDPGAgent, ReplayBuffer, _lib as L

    torch.manual_seed(1 + rank)
    np.random.seed(1 + rank)
    from controllable_agent_b200.dist_utils import shard_episodes
    e0_, e1_ = shard_episodes(a.episodes, world, rank)
    E = e1_ - e0_
    R = a.episode_len + 1
    replay = ReplayBuffer(max_episodes=E, discount=0.98, future=0.99, device=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    storage = {"observation": torch.randn((E, R, a.obs_dim), device=dev, generator=g),
               "action": torch.rand((E, R, a.action_dim), device=dev, generator=g) * 2 - 1,
               "reward": torch.rand((E, R, 1), device=dev, generator=g),
               "discount": torch.ones((E, R, 1), device=dev)}
    if a.goal_space:
        storage["goal"] = torch.randn((E, R, GOAL_DIMS[a.goal_space]), device=dev, generator=g)
    replay.load_storage(storage)
    del storage
    common = dict(obs_type="states", obs_shape=(a.obs_dim,), action_shape=(a.action_dim,), device=str(dev), num_expl_steps=0,
                  update_encoder=True, goal_space=a.goal_space, update_every_steps=1, batch_size=a.batch, z_dim=a.z_dim,
                  use_cuda_graph=not a.no_graph)
    if a.collectives != "auto":
        common["collectives"] = a.collectives
    agent = FBDDPGAgent(use_tb=False, use_wandb=False, use_hiplog=False, rng_mode="device", mlp_mode=a.mlp_mode, **common)
    eng = agent.engine

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- parity self-check: the N-rank path against one engine on the whole batch (the only place real NCCL / P2P ranks exist) ----
    parity = None
    if not a.no_parity_check:
        parity = parity_check(agent, a, world, rank, dev)
        ok = torch.tensor([1 if (rank != 0 or parity.get("ok", False)) else 0], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "parity_check failed", "parity_check": parity}), flush=True)
            if world > 1:
                dist.destroy_process_group()
            raise SystemExit(3)
        barrier()

    # ---- device-resident leg: `value` --------------------------------------------------------------------------------
    for i in range(max(a.warmup, 3)):
        agent.update(replay, i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(a.steps):
        agent.update(replay, i)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler is not None else None
    ms_total = float(ms.item())
    launches_per_step = agent.last_update_launches

    # ---- end-to-end leg: host buffers, H2D of the step's inputs and D2H of its metrics inside the timed region -------
    e2e_steps = a.e2e_steps or min(a.steps, 100)
    hstore = host_storage(a, E, seed=7 + rank, physics_dim=18 if (rank == 0 and world == 1 and not a.no_cpu_baseline) else 0)
    host = HostReplay(hstore, 0.98, a.episode_len)
    agent.cfg.use_tb = True          # metrics on: one D2H read of the step's losses per step

    def e2e_leg(prefetch: bool, native: bool = True) -> float:
        agent.cfg.prefetch_host_batch = prefetch
        agent.native_host_sampling = native
        agent._prefetched = None
        for i in range(3):
            agent.update(host, i)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        m_ = None
        for i in range(e2e_steps):
            m_ = agent.update(host, i)
        ev1.record()
        barrier()
        t = torch.tensor([max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_leg.metrics = m_
        return float(t.item())

    e2e_ms = e2e_leg(False)            # default flags: sample -> upload -> step -> read, strictly in sequence (the reference's own order)
    e2e_ms_prefetch = e2e_leg(True)    # opt-in cfg.prefetch_host_batch: the next batch is sampled and uploaded while the step runs
    e2e_ms_python = e2e_leg(False, native=False)   # the replay object's own Python sample() (what any foreign replay object gets)
    agent.native_host_sampling = True
    m = e2e_leg.metrics
    agent.cfg.prefetch_host_batch = False
    Bl = a.batch // world
    h2d = 4 * Bl * eng._row_pitch   # one copy of the packed batch rows [obs | action | reward, discount | next_obs (| goals)], 16-byte aligned fields
    d2h = 4 * L.METRIC_COUNT
    agent.cfg.use_tb = False

    # ---- per-kernel timings (CUDA events between launches, eager) -> roofline of the dominant kernel -----------------
    line: dict = {}
    # (N > 1: every rank walks the same eager launch sequence in lock step — the exchange kernels / NCCL calls of the plan keep them
    # aligned — so a collective's time is what it costs this rank including the wait for its peers)
    ops = []
    if world == 1 or agent.collectives_mode in ("p2p", "graph"):
        barrier()
        ops = eng.profile_ops(L.PHASE_ALL & ~L.PHASE_METRICS, reps=5)
        barrier()
    if rank == 0 and a.timeline:
        names = ["SAMPLE", "MIX", "FB_FWD", "FB_LOSS", "FB_BWD", "FB_ADAM", "ACTOR_FWD", "ACTOR_BWD", "ACTOR_ADAM", "METRICS"]
        bounds, acc_ = [], 0
        for ph in range(10):
            acc_ += eng.launch_count(1 << ph, fused=False)
            bounds.append(acc_)
        with open(a.timeline, "w") as f:
            tot = sum(o["ms"] for o in ops)
            coll = sum(o["ms"] for o in ops if o["kind"] == "collective")
            f.write(f"# rank 0 of {world}: per-launch CUDA-event durations of one step, eager and serialised (every entry carries ~4 us of event "
                    f"overhead); batch {a.batch} global / {a.batch // world} per rank, z {a.z_dim}, collectives = {agent.collectives_mode}\n")
            f.write(f"# {len(ops)} launches, {tot:.3f} ms summed: collective {coll * 1e3:.1f} us, everything else {1e3 * (tot - coll):.1f} us\n")
            ph = 0
            for i, o in enumerate(ops):
                while ph < 9 and i >= bounds[ph]:
                    ph += 1
                f.write(f"{i:3d} {names[ph]:10s} {o['kind']:11s} {o['ms'] * 1e3:8.1f} us\n")
    if rank == 0:
        by_kind: dict = {}
        for o in ops:
            k = by_kind.setdefault(o["kind"], {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
            k["ms"] += o["ms"]; k["launches"] += 1; k["flops"] += o["flops"]; k["bytes"] += o["bytes"]
        roofline = None
        total_ms = sum(v["ms"] for v in by_kind.values()) or 1.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        if "gemm_tc" in by_kind:
            # dominant kernel: the tcgen05 3xTF32 grouped GEMM.  Tensor roofline for fp32-grade products on this kernel:
            # measured dense bf16 rate / 2 (kind::tf32 runs at half the bf16 rate) / 3 (three MMA chains per product).  The timed
            # region is tens of milliseconds at full clocks, far from the power-capped steady state: the BURST figure applies.
            burst = float(peaks.get("bf16_tflops", 1640.0))
            sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
            src = "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1.64 PFLOP/s burst (B200_PROFILING.md)"
            gk = by_kind["gemm_tc"]
            achieved = gk["flops"] / (gk["ms"] * 1e-3) / 1e12
            peak = burst / 6.0
            tp_ = profile_traffic("k_gemm_tc")
            roofline = {"kernel": "k_gemm_tc (tcgen05 kind::tf32, 3xTF32 split, TMA + TMEM): all wide nn.Linear forward / dX / dW products",
                        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                        "peak_sustained": sustained / 6.0, "frac_of_sustained": achieved / (sustained / 6.0),
                        # not measured in this run: bench.py cannot read DRAM counters; read from the committed ncu --set full capture
                        "traffic": tp_["bytes_per_launch"] if tp_ else None, "traffic_from_profile": tp_,
                        "algorithmic_bytes_per_launch": gk["bytes"] / gk["launches"],
                        "peak_source": f"{src} = {burst:.1f} TFLOP/s dense bf16; /2 for tf32, /3 for the three chains of an fp32-grade product "
                                       "(achieved counts each algorithmic fp32 FLOP once); timing = CUDA events around each eager launch "
                                       "(each carries ~4 us of event overhead)",
                        "achieved_tensor_tflops_tf32": 3.0 * achieved, "frac_of_tf32_peak": 3.0 * achieved / (burst / 2.0),
                        "launches_per_step": gk["launches"], "avg_launch_us": 1e3 * gk["ms"] / gk["launches"],
                        "algorithmic_gflop_per_step": gk["flops"] / 1e9, "share_of_step": gk["ms"] / total_ms,
                        "step_level": {"achieved": gk["flops"] / (ms_total / a.steps * 1e-3) / 1e12, "frac": gk["flops"] / (ms_total / a.steps * 1e-3) / 1e12 / peak,
                                       "note": "all GEMM FLOPs of a step / the graph-replayed step time (no event overhead)"}}
        elif "gemm" in by_kind:
            import ctypes as C
            peak = C.c_double()
            L.check(L.load().fb_fp32_peak_tflops(C.byref(peak), torch.cuda.current_stream(dev).cuda_stream))
            gk = by_kind["gemm"]
            achieved = gk["flops"] / (gk["ms"] * 1e-3) / 1e12
            roofline = {"kernel": "k_gemm_grouped (fp32 SIMT grouped SGEMM: all MLP forward/backward layers)", "bound": "fp32_fma",
                        "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value, "traffic": None,
                        "peak_source": "FMA-chain microbenchmark (fb_fp32_peak_tflops) measured in this run; MEASURED_PEAKS.json has "
                                       "no fp32 CUDA-core figure (nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4)",
                        "launches_per_step": gk["launches"], "avg_launch_us": 1e3 * gk["ms"] / gk["launches"],
                        "algorithmic_gflop_per_step": gk["flops"] / 1e9, "share_of_step": gk["ms"] / total_ms}
        if roofline is not None and "adam" in by_kind:
            ak = by_kind["adam"]
            hbm = float(peaks.get("hbm_gbs", 6650.0))
            roofline["secondary"] = {"kernel": "k_adam (Adam + target soft update + gradient clear)", "bound": "hbm",
                                     "achieved": ak["bytes"] / (ak["ms"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                     "frac": ak["bytes"] / (ak["ms"] * 1e-3) / 1e9 / hbm}
        cpu = cuda_eager = None
        if world == 1 and not a.no_cpu_baseline:
            del host, hstore
            r_ = time_reference(a, "cpu", a.cpu_steps, 2, budget_s=30.0, episodes=a.episodes)
            cpu = {k: r_[k] for k in ("value", "unit", "cores", "host_cpus", "kind", "sample")}
            if not a.no_cuda_eager:
                # the Blackwell bar of SURVEY.md 2a / 8d: the same unmodified reference with device=cuda (PyTorch eager kernels), on this GPU
                try:
                    r_ = time_reference(a, str(dev), 60, 5, budget_s=30.0, episodes=a.episodes)
                    cuda_eager = {k: r_[k] for k in ("value", "unit", "kind", "sample")}
                    cuda_eager["note"] = "reference FBDDPGAgent on device=cuda through PyTorch eager; host replay + EpisodeBatch.to(device) per step, as the reference runs it"
                except Exception as exc:   # noqa: BLE001 — a reported leg, never fatal
                    cuda_eager = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        cfg = workload(a)
        coll = getattr(agent, "collectives_mode", a.collectives)
        cfg.update({"parallelism": f"dp{world} (batch rows sharded; exchange of the F/B row blocks + reduction of the flat gradients, "
                                   f"collectives={coll})" if world > 1 else "single GPU", "per_gpu_batch": Bl, "cuda_graph": not a.no_graph,
                    "rng": "device Philox inside the step graph", "mlp_mode": a.mlp_mode,
                    "replay": f"value: {a.episodes} episodes resident in HBM (sharded {a.episodes // world}/rank); e2e: {E} episodes/rank in host memory",
                    "l2": "inputs exceed L2: each step gathers random rows of a replay far larger than the 126 MB L2; weights and "
                          "activations are re-used step to step exactly as in training"})
        line = {"metric": METRIC, "value": a.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "steps": e2e_steps, "value_with_prefetch": e2e_steps / (e2e_ms_prefetch * 1e-3),
                        "value_python_sample": e2e_steps / (e2e_ms_python * 1e-3),
                        "path": "FBDDPGAgent.update(host_replay, step) at default flags, host_replay = a buffer in host memory with the reference "
                                "ReplayBuffer's layout: numpy index draws (the reference's order) -> fb_host_gather_rows into pinned packed rows -> "
                                "one H2D -> step graph (device RNG) -> metrics block D2H, strictly in sequence every step; value_with_prefetch: "
                                "the opt-in agent.prefetch_host_batch=True (next step's sample + upload issued while this step runs on the GPU); "
                                "value_python_sample: the replay object's own numpy sample() instead of the library's host gather"},
                "gpu_launches": launches_per_step * a.steps, "launches_per_step": launches_per_step,
                "roofline": roofline, "cpu_baseline": cpu, "reference_cuda_eager": cuda_eager, "parity_check": parity,
                "breakdown_ms": {k: round(v["ms"], 4) for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1]["ms"])},
                "last_metrics": {k: m.get(k) for k in ("fb_loss", "actor_loss")}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    a = parse()
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 behind Python's back (NCCL prints
    # "NCCL version ..." there when NCCL_DEBUG is set on the box) are pointed at stderr for the whole run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w", buffering=1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
