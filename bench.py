#!/usr/bin/env python
"""bench.py — FB-DDPG gradient-steps/sec (batch=1024, z_dim=50) on N B200s (BASELINE.json metric).

One "step" = one `agent.update(replay, step)` with update_every_steps=1 (train_offline.py:59,118): replay sample,
z draw + mixing, update_fb (+Adam), update_actor (+Adam), both target soft updates.

  python bench.py [--gpus N] [--steps K] [--warmup W]                  our arm (CUDA step, device-resident replay)
  python bench.py --impl reference [--steps K] [--warmup W]            the reference algorithm on the host CPU cores
                                                                       (oracle/ port, all host threads; rank 0 only)
N > 1 is launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`; the GLOBAL batch stays
1024 (strong scaling, BASELINE.json north_star), each rank steps 1024/N rows against its own replay shard and the ranks
exchange the [batch, 6*z] embedding block (all-gather) and the two flat gradients (all-reduce) over NCCL.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
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
    p.add_argument("--episodes", type=int, default=5000, help="episodes of the synthetic replay (whole job)")
    p.add_argument("--episode-len", type=int, default=1000)
    p.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer end-to-end leg (0: min(steps, 100))")
    p.add_argument("--cpu-steps", type=int, default=12, help="timed steps of the cpu_baseline leg (rank 0, N=1)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--collectives", default="graph", choices=["graph", "torch"], help="multi-GPU exchange: NCCL calls captured in the step "
                   "graph on the library's communicator, or torch.distributed calls between graph segments")
    p.add_argument("--mlp-mode", default="tcgen05", choices=["tcgen05", "simt"], help="wide Linear products: tensor cores (3xTF32) or fp32 CUDA cores")
    return p.parse_args()


def workload(a: argparse.Namespace) -> dict:
    return {"workload": f"walker_walk fb_ddpg offline (BASELINE.json configs[1]): obs={a.obs_dim} act={a.action_dim} z={a.z_dim} "
                        f"batch={a.batch} hidden=1024 feature=512 backward_hidden=526, {a.episodes}x{a.episode_len}-step synthetic "
                        "replay resident in HBM, update_every_steps=1",
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
# the reference algorithm on the host CPU (oracle/ is test infrastructure: only this leg and --impl reference run it)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps_per_sec(a: argparse.Namespace, steps: int, warmup: int, budget_s: float = 150.0) -> dict:
    import numpy as np
    import torch
    from oracle import fb_oracle as O
    torch.manual_seed(1)
    np.random.seed(1)
    d = O.Dims(obs_dim=a.obs_dim, action_dim=a.action_dim, z_dim=a.z_dim, goal_dim=a.obs_dim)
    agent = O.OracleAgent(O.OracleConfig(dims=d, batch_size=a.batch, metrics=False))
    n_ep = 50
    replay = O.OracleReplay(n_ep, 0.98, 0.99)
    rng = np.random.RandomState(0)
    for _ in range(n_ep):
        replay.add_episode(O.synthetic_episode(rng, a.episode_len, d))
    t0 = time.perf_counter()
    for i in range(max(warmup, 1)):
        agent.update(replay, i)
    per = (time.perf_counter() - t0) / max(warmup, 1)
    done = steps
    if per * steps > budget_s:
        done = max(3, int(budget_s / per))
    t0 = time.perf_counter()
    for i in range(done):
        agent.update(replay, i)
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
            "steps": done, "ms_per_step": 1e3 * dt / done,
            "sample": f"{done} full agent.update() steps at batch={a.batch} (oracle/fb_oracle.py OracleAgent = the reference's "
                      f"torch-CPU op sequence, {torch.get_num_threads()} threads) on a {n_ep}x{a.episode_len}-step synthetic replay"}


def run_reference(a: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps_per_sec(a, a.steps, a.warmup)
    cfg = workload(a)
    cfg["parallelism"] = f"host CPU, {r['cores']} torch threads"
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
    """A host-memory replay with the reference's sample() contract (in_memory_replay_buffer.py:139-190: numpy index draws and
    fancy-index gathers returning numpy arrays), for the end-to-end leg: every step's batch is gathered on the host and
    crosses PCIe inside the timed region."""

    def __init__(self, obs, action, reward, discount, gamma: float) -> None:
        self.obs, self.action, self.reward, self.discount = obs, action, reward, discount
        self._discount, self._future = gamma, 1.0

    def sample(self, batch_size: int):
        import numpy as np
        from controllable_agent_b200 import EpisodeBatch
        E, R = self.obs.shape[:2]
        ep = np.random.randint(0, E, size=batch_size)
        t = np.random.randint(0, R - 1, size=batch_size) + 1
        return EpisodeBatch(obs=self.obs[ep, t - 1], action=self.action[ep, t], reward=self.reward[ep, t],
                            discount=self.discount[ep, t] * self._discount, next_obs=self.obs[ep, t])


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
    replay.load_storage(storage)
    del storage
    common = dict(obs_type="states", obs_shape=(a.obs_dim,), action_shape=(a.action_dim,), device=str(dev), num_expl_steps=0,
                  update_encoder=True, goal_space=None, update_every_steps=1, batch_size=a.batch, z_dim=a.z_dim,
                  use_cuda_graph=not a.no_graph)
    agent = FBDDPGAgent(use_tb=False, use_wandb=False, use_hiplog=False, rng_mode="device", mlp_mode=a.mlp_mode, collectives=a.collectives,
                        **common)
    eng = agent.engine

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

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
    rs = np.random.RandomState(7 + rank)
    Eh = min(E, 200)
    host = HostReplay(rs.standard_normal((Eh, R, a.obs_dim)).astype(np.float32), rs.uniform(-1, 1, (Eh, R, a.action_dim)).astype(np.float32),
                      rs.uniform(0, 1, (Eh, R, 1)).astype(np.float32), np.ones((Eh, R, 1), np.float32), 0.98)
    agent.cfg.use_tb = True          # metrics on: one D2H read of the step's losses per step

    def e2e_leg(prefetch: bool) -> float:
        agent.cfg.prefetch_host_batch = prefetch
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

    e2e_ms_serial = e2e_leg(False)   # sample -> upload -> step -> read, strictly in sequence (the reference's own order)
    e2e_ms = e2e_leg(True)           # the next batch is sampled and uploaded while the step runs (cfg.prefetch_host_batch)
    m = e2e_leg.metrics
    agent.cfg.prefetch_host_batch = False
    Bl = a.batch // world
    h2d = 4 * Bl * eng._row_pitch   # one copy of the packed batch rows [obs | action | reward, discount | next_obs], 16-byte aligned fields
    d2h = 4 * L.METRIC_COUNT
    agent.cfg.use_tb = False

    # ---- per-kernel timings (CUDA events between launches, eager) -> roofline of the dominant kernel -----------------
    line: dict = {}
    if rank == 0:
        ops = eng.profile_ops(L.PHASE_ALL & ~L.PHASE_METRICS, reps=5) if world == 1 else []
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
            # measured dense bf16 rate / 2 (kind::tf32 runs at half the bf16 rate) / 3 (three MMA chains per product).
            bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
            src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
            gk = by_kind["gemm_tc"]
            achieved = gk["flops"] / (gk["ms"] * 1e-3) / 1e12
            peak = bf16 / 6.0
            roofline = {"kernel": "k_gemm_tc (tcgen05 kind::tf32, 3xTF32 split, TMA + TMEM): all wide nn.Linear forward / dX / dW products",
                        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                        # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the step's 29 k_gemm_tc launches of the
                        # committed ncu --set full capture (profiles/r1d_ncu_key_metrics.csv); algorithmic operand + result bytes
                        # of the same launches: bytes_per_launch below
                        "traffic": 19.66e6, "traffic_unit": "bytes/launch (ncu, profiles/r1d_ncu_key_metrics.csv)",
                        "algorithmic_bytes_per_launch": gk["bytes"] / gk["launches"],
                        "peak_source": f"{src} = {bf16:.1f} TFLOP/s dense bf16; /2 for tf32, /3 for the three chains of an fp32-grade product "
                                       "(achieved counts each algorithmic fp32 FLOP once)",
                        "achieved_tensor_tflops_tf32": 3.0 * achieved, "frac_of_tf32_peak": 3.0 * achieved / (bf16 / 2.0),
                        "launches_per_step": gk["launches"], "avg_launch_us": 1e3 * gk["ms"] / gk["launches"],
                        "algorithmic_gflop_per_step": gk["flops"] / 1e9, "share_of_step": gk["ms"] / total_ms}
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
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            cpu = cpu_reference_steps_per_sec(a, a.cpu_steps, 2, budget_s=40.0)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "host_cpus", "kind", "sample")}
        cfg = workload(a)
        cfg.update({"parallelism": f"dp{world} (batch rows sharded; all-gather of the F/B row blocks + all-reduce of the flat gradients, "
                                   f"collectives={a.collectives})" if world > 1 else "single GPU", "per_gpu_batch": Bl, "cuda_graph": not a.no_graph,
                    "rng": "device Philox inside the step graph", "mlp_mode": a.mlp_mode,
                    "l2": "inputs exceed L2: each step gathers random rows of a replay far larger than the 126 MB L2; weights and "
                          "activations are re-used step to step exactly as in training"})
        line = {"metric": METRIC, "value": a.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "steps": e2e_steps, "value_without_prefetch": e2e_steps / (e2e_ms_serial * 1e-3),
                        "path": "FBDDPGAgent.update(host_replay, step): host numpy sample() -> pinned packed rows -> one H2D -> step graph (device RNG) -> "
                                "metrics block D2H, every step; value: agent.prefetch_host_batch=True (the next step's sample + upload is "
                                "issued while this step runs on the GPU), value_without_prefetch: strictly serial"},
                "gpu_launches": launches_per_step * a.steps, "launches_per_step": launches_per_step,
                "roofline": roofline, "cpu_baseline": cpu,
                "breakdown_ms": {k: round(v["ms"], 4) for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1]["ms"])},
                "last_metrics": {k: m.get(k) for k in ("fb_loss", "actor_loss")}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
