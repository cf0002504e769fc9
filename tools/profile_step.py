"""Per-launch table of one gradient step (CUDA events between launches) + an SGEMM sweep.  GPU only."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controllable_agent_b200 import FBDDPGAgent, ReplayBuffer, _lib as L  # noqa: E402


def _parse_overrides(items):
    """`--agent q_loss=true add_trunk=true norm_z=false q_loss_coef=0.5` -> FBDDPGAgent keyword overrides."""
    out = {}
    for it in items or []:
        k, v = it.split("=", 1)
        low = v.lower()
        out[k] = True if low == "true" else False if low == "false" else (float(v) if any(c in v for c in ".e") else int(v))
    return out


def step_table(batch: int, z_dim: int, reps: int, overrides=None, fused: bool = False) -> None:
    dev = torch.device("cuda")
    E, R, O_, A_ = 200, 1001, 24, 6
    replay = ReplayBuffer(E, 0.98, 0.99, device=dev)
    replay.load_storage({"observation": torch.randn(E, R, O_, device=dev), "action": torch.rand(E, R, A_, device=dev) * 2 - 1,
                         "reward": torch.rand(E, R, 1, device=dev), "discount": torch.ones(E, R, 1, device=dev)})
    agent = FBDDPGAgent(obs_type="states", obs_shape=(O_,), action_shape=(A_,), device="cuda", num_expl_steps=0, update_encoder=True,
                        goal_space=None, use_tb=False, use_wandb=False, use_hiplog=False, batch_size=batch, z_dim=z_dim, update_every_steps=1,
                        **(overrides or {}))
    for i in range(5):
        agent.update(replay, i)
    torch.cuda.synchronize()
    if fused:
        rows = agent.engine.fused_profile(L.PHASE_ALL, reps=reps)
        tot = sum(r["us"] for r in rows)
        print(f"batch={batch} z={z_dim} {overrides or ''}: fused execution, {len(set(r['unit'] for r in rows))} launches, "
              f"{len(rows)} stages + kernels, {tot / 1e3:.3f} ms summed")
        for i, r in enumerate(rows):
            where = f"unit {r['unit']:2d} stage {r['stage']:2d}" if r["stage"] >= 0 else f"unit {r['unit']:2d} kernel  "
            print(f"{i:3d} {where} {r['us']:8.1f} us  items {r['items']}  first {r['first']} x{r['count']}")
        return
    ops = agent.engine.profile_ops(L.PHASE_ALL, reps=reps)
    names = ["SAMPLE", "MIX", "FB_FWD", "FB_LOSS", "FB_BWD", "FB_ADAM", "ACTOR_FWD", "ACTOR_BWD", "ACTOR_ADAM", "METRICS"]
    bounds, acc = [], 0
    for ph in range(10):
        acc += agent.engine.launch_count(1 << ph, fused=False)
        bounds.append(acc)
    tot = sum(o["ms"] for o in ops)
    print(f"batch={batch} z={z_dim} {overrides or ''}: {len(ops)} launches, {tot:.3f} ms eager-serial")
    ph = 0
    for i, o in enumerate(ops):
        while i >= bounds[ph]:
            ph += 1
        tf = o["flops"] / (o["ms"] * 1e-3) / 1e12 if o["flops"] else 0.0
        gbs = o["bytes"] / (o["ms"] * 1e-3) / 1e9
        print(f"{i:3d} {names[ph]:10s} {o['kind']:11s} {o['ms'] * 1e3:8.1f} us  {o['flops'] / 1e9:7.3f} GFLOP {tf:6.1f} TF/s  {gbs:7.0f} GB/s")


def sgemm_sweep() -> None:
    lib = L.load()
    s = torch.cuda.current_stream().cuda_stream
    for (M, N, K, ak, bk) in [(4096, 4096, 4096, 1, 1), (4096, 4096, 4096, 1, 0), (4096, 4096, 4096, 0, 0), (1024, 1024, 1024, 1, 1),
                              (2048, 1024, 1024, 1, 1), (1024, 512, 1024, 1, 1), (1024, 1024, 1024, 0, 0), (1024, 50, 1024, 1, 1),
                              (1024, 1024, 74, 1, 1)]:
        A = torch.randn((M, K) if ak else (K, M), device="cuda")
        B = torch.randn((N, K) if bk else (K, N), device="cuda")
        Cm = torch.zeros(M, N, device="cuda")
        args = (A.data_ptr(), B.data_ptr(), Cm.data_ptr(), None, M, N, K, A.shape[1], B.shape[1], N, ak, bk, 0, 1, -1, s)
        for _ in range(3):
            L.check(lib.fb_sgemm(*args))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0.record()
            L.check(lib.fb_sgemm(*args))
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        # cuBLAS (torch.matmul, TF32 off) for comparison
        torch.backends.cuda.matmul.allow_tf32 = False
        At, Bt = (A if ak else A.t()), (B.t() if bk else B)
        for _ in range(3):
            At @ Bt
        bc = 1e9
        for _ in range(5):
            e0.record()
            At @ Bt
            e1.record()
            torch.cuda.synchronize()
            bc = min(bc, e0.elapsed_time(e1))
        fl = 2.0 * M * N * K
        print(f"sgemm M={M} N={N} K={K} ak={ak} bk={bk}: {best * 1e3:8.1f} us {fl / best / 1e9:6.1f} TF/s (incl. desc upload)   "
              f"cuBLAS fp32 {bc * 1e3:8.1f} us {fl / bc / 1e9:6.1f} TF/s")
    peak = C.c_double()
    L.check(lib.fb_fp32_peak_tflops(C.byref(peak), s))
    print(f"fp32 FMA-chain peak: {peak.value:.1f} TFLOP/s")


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--batch", type=int, default=1024)
    p.add_argument("--z-dim", type=int, default=50)
    p.add_argument("--reps", type=int, default=10)
    p.add_argument("--sweep", action="store_true")
    p.add_argument("--fused", action="store_true", help="per-stage table of the fused execution instead of the per-launch table")
    p.add_argument("--agent", nargs="*", default=[], help="FBDDPGAgent config overrides, e.g. q_loss=true add_trunk=true rand_weight=true")
    a = p.parse_args()
    step_table(a.batch, a.z_dim, a.reps, _parse_overrides(a.agent), fused=a.fused)
    if a.sweep:
        sgemm_sweep()
