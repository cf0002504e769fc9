"""Debug: which elements of the actor's first-layer weight gradient differ from the oracle (GPU)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_common import load_params, make_engine, read_tensors, rel
from oracle import fb_oracle as O
from controllable_agent_b200 import _lib as L

d = O.Dims()
B = 256
gen = torch.Generator().manual_seed(11)
actor = O.init_params(O.actor_spec(d), gen); fwd = O.init_params(O.forward_map_spec(d), gen); bwd = O.init_params(O.backward_map_spec(d), gen)
fwd_t = {k: v + 0.02 * torch.randn(v.shape, generator=gen) for k, v in fwd.items()}
bwd_t = {k: v + 0.02 * torch.randn(v.shape, generator=gen) for k, v in bwd.items()}
obs, next_obs = torch.randn(B, d.obs_dim, generator=gen), torch.randn(B, d.obs_dim, generator=gen)
action = torch.rand(B, d.action_dim, generator=gen) * 2 - 1
discount = torch.full((B, 1), 0.98)
z = O.sample_z(B, d.z_dim, gen)
nf, na = torch.randn(B, d.action_dim, generator=gen), torch.randn(B, d.action_dim, generator=gen)
fwd_t = {k: v + 0.02 * torch.randn(v.shape, generator=gen) for k, v in fwd.items()}
perm = torch.randperm(B, generator=gen)
mix_mask = (torch.rand(B, generator=gen) < 0.5)
eng = make_engine(d, B)
load_params(eng, fwd=fwd, bwd=bwd, actor=actor, fwd_tgt=fwd_t, bwd_tgt=bwd_t)
eng.set_scalars(0.2, 0.3, 1e-4, 1e-4, 1e-4, 0.01)
eng.set_indices(perm=perm, mix_mask=mix_mask.int())
eng.set_batch(obs, action, discount, next_obs); eng.set_z(z); eng.set_noise(nf, na)
eng.run(L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_FB_BWD | L.PHASE_METRICS)
eng.run(L.PHASE_FB_ADAM)
fwd1 = read_tensors(eng, L.NET_FORWARD, "param")
eng.run(L.PHASE_ACTOR_FWD | L.PHASE_ACTOR_BWD | L.PHASE_METRICS)
torch.cuda.synchronize()
z = eng.view("z").detach().cpu()
dt = torch.float64
c = lambda p: {k: v.to(dt) for k, v in p.items()}
ora = O.actor_loss_and_grads(c(actor), c(fwd1), obs.to(dt), z.to(dt), na.to(dt), 0.2, 0.3)
o32 = O.actor_loss_and_grads(actor, fwd1, obs, z, na, 0.2, 0.3)
got = read_tensors(eng, L.NET_ACTOR, "grad")
for name, ref in ora["grads_actor"].items():
    g = got[name].double().numpy(); r = ref.numpy()
    e = rel(g, r)
    diff = np.abs(g - r)
    idx = np.unravel_index(np.argmax(diff), diff.shape)
    if name == "obs_net.0.weight":
        bad = np.argwhere(diff > 1e-4 * np.abs(r).max())
        print("bad elements:", len(bad), bad[:20].tolist(), [(float(g[tuple(i)]), float(r[tuple(i)])) for i in bad[:6]])
    print(f"{name:24s} own {rel(o32['grads_actor'][name], ref):.2e} rel {e:.3e} max|diff| {diff.max():.3e} at {idx} got {g[idx]:.6e} ref {r[idx]:.6e} |ref|max {np.abs(r).max():.3e} n_bad {(diff > 1e-3 * np.abs(r).max()).sum()}")
