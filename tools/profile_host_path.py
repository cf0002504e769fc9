"""Where does the host side of FBDDPGAgent.update(host_replay, step) spend its time?  cProfile over the end-to-end leg of bench.py
(host replay in the reference layout, metrics on, default flags).  GPU only."""
import argparse
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from controllable_agent_b200 import FBDDPGAgent  # noqa: E402


def main() -> None:
    a = argparse.Namespace(obs_dim=24, action_dim=6, z_dim=50, batch=1024, goal_space=None, episodes=2000, episode_len=1000)
    agent = FBDDPGAgent(obs_type="states", obs_shape=(a.obs_dim,), action_shape=(a.action_dim,), device="cuda", num_expl_steps=0,
                        update_encoder=True, goal_space=None, use_tb=True, use_wandb=False, use_hiplog=False, batch_size=a.batch,
                        z_dim=a.z_dim, update_every_steps=1)
    host = bench.HostReplay(bench.host_storage(a, a.episodes, seed=1), 0.98, a.episode_len)
    for i in range(20):
        agent.update(host, i)
    torch.cuda.synchronize()
    for label, steps in (("wall", 300),):
        t0 = time.perf_counter()
        for i in range(steps):
            agent.update(host, i)
        torch.cuda.synchronize()
        print(f"{label}: {1e3 * (time.perf_counter() - t0) / steps:.3f} ms / update")
    # the same loop without reading the metrics (no device sync per step): host time per update when the GPU is not waited for
    agent.cfg.use_tb = False
    t0 = time.perf_counter()
    for i in range(300):
        agent.update(host, i)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"no metrics read: host enqueue {1e3 * t_host / 300:.3f} ms / update, total {1e3 * (time.perf_counter() - t0) / 300:.3f} ms / update")
    agent.cfg.use_tb = True
    pr = cProfile.Profile()
    pr.enable()
    for i in range(300):
        agent.update(host, i)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
