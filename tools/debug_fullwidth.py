"""Debug: compare engine intermediates with torch autograd at full width (GPU)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_common import load_params, make_engine, read_tensors, rel
from oracle import fb_oracle as O
from controllable_agent_b200 import _lib as L
import torch.nn.functional as F

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = O.Dims(hidden_dim=H, feature_dim=H // 2)
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
for mode in (1, 0):
    eng = make_engine(d, B, contract_mode=mode, mix_ratio=0.0)
    load_params(eng, fwd=fwd, bwd=bwd, actor=actor, fwd_tgt=fwd_t, bwd_tgt=bwd_t)
    eng.set_scalars(0.2, 0.3, 1e-4, 1e-4, 1e-4, 0.01)
    eng.set_batch(obs, action, discount, next_obs); eng.set_z(z); eng.set_noise(nf, na)
    eng.run(L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_FB_BWD | L.PHASE_METRICS)
    torch.cuda.synchronize()
    for dt in (torch.float32, torch.float64):
        c = lambda p: {k: v.to(dt) for k, v in p.items()}
        ora = O.fb_loss_and_grads(c(fwd), c(bwd), c(fwd_t), c(bwd_t), c(actor), obs.to(dt), action.to(dt), discount.to(dt), next_obs.to(dt),
                                  next_obs.to(dt), z.to(dt), nf.to(dt), 0.2, 0.3, 1.0, d.z_dim)
        print(f"--- contract_mode={mode} oracle dtype={dt}")
        for name in ("next_action", "tF1", "tF2", "tB", "F1", "F2", "B", "dF1", "dF2", "dB"):
            print(f"   {name:12s} {rel(eng.view(name), ora[name]):.3e}")
        for net, key in ((L.NET_FORWARD, "grads_forward"), (L.NET_BACKWARD, "grads_backward")):
            got = read_tensors(eng, net, "grad")
            for name, ref in ora[key].items():
                print(f"   {key}/{name:28s} {rel(got[name], ref):.3e}")
        if dt == torch.float64:
            o32 = O.fb_loss_and_grads(fwd, bwd, fwd_t, bwd_t, actor, obs, action, discount, next_obs, next_obs, z, nf, 0.2, 0.3, 1.0, d.z_dim)
            print("   oracle fp32 vs fp64:")
            for key in ("grads_forward", "grads_backward"):
                for name, ref in ora[key].items():
                    print(f"   {key}/{name:28s} {rel(o32[key][name], ref):.3e}")
    eng.close()
