"""ctypes binding of libfb_b200.so (C ABI: include/fb_b200.h).

The library is built in-tree by `build_library()` (nvcc, sm_100a only) and loaded from this package
directory.  There is no CPU route: if the shared object is missing or does not export the ABI this
module raises, it never falls back to PyTorch.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import typing as tp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libfb_b200.so")
SOURCES = [os.path.join(HERE, "csrc", n) for n in ("fb_b200.cu", "plan.cuh", "kernels.cuh", "gemm_simt.cuh", "gemm_tc.cuh", "contract_tc.cuh", "common.cuh")]
CONTRACT_TCGEN05, CONTRACT_SIMT = 0, 1
MLP_TCGEN05, MLP_SIMT = 0, 1
HEADER = os.path.join(ROOT, "include", "fb_b200.h")

FB_ABI_VERSION = 6
FB_OK = 0

NET_FORWARD, NET_BACKWARD, NET_ACTOR = 0, 1, 2

PHASE_SAMPLE = 1 << 0
PHASE_MIX = 1 << 1
PHASE_FB_FWD = 1 << 2
PHASE_FB_LOSS = 1 << 3
PHASE_FB_BWD = 1 << 4
PHASE_FB_ADAM = 1 << 5
PHASE_ACTOR_FWD = 1 << 6
PHASE_ACTOR_BWD = 1 << 7
PHASE_ACTOR_ADAM = 1 << 8
PHASE_METRICS = 1 << 9
PHASE_ALL = (1 << 10) - 1
PHASE_INFER_ACTOR, PHASE_INFER_B, PHASE_INFER_BN = 1 << 10, 1 << 11, 1 << 12   # inference plans (run on their own)
INFER_ROWS = 8
RUN_HOST_BATCH = 1 << 15   # modifier of PHASE_SAMPLE: batch rows supplied by the caller (fb_upload_batch), gather skipped
RUN_UNFUSED = 1 << 14      # modifier: every launch of the plan as a kernel of its own (default: fused segments, k_fused_stack)

# index of each scalar of the metrics block (FB_M_* in fb_b200.h) -> key of the dict FBDDPGAgent.update returns
METRIC_KEYS = ("target_M", "M1", "F1", "B", "B_norm", "z_norm", "fb_loss", "fb_diag", "fb_offdiag", "orth_loss",
               "orth_loss_diag", "orth_loss_offdiag", "orth_linf", "orth_l2", "actor_loss", "q", "actor_logprob")
# slots behind them, filled on every step but only reported for the matching config flag (cfg.q_loss / cfg.additional_metric)
OPTIONAL_METRIC_KEYS = ("q_loss", "q1_success")
METRIC_COUNT = 32
FS_TYPES = ("none", "gemm_tc", "ln_fwd", "ln_bwd", "transpose", "colsum", "l2_fwd", "l2_bwd", "stage_inputs", "z_final", "actor_out", "actor_q")
OP_KINDS = ("gemm", "layernorm", "elementwise", "colsum", "adam", "gather", "loss", "memset", "contract", "gemm_tc", "transpose", "collective")


class fb_config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("batch", C.c_int32), ("global_batch", C.c_int32), ("row_offset", C.c_int32),
                ("obs_dim", C.c_int32), ("action_dim", C.c_int32), ("z_dim", C.c_int32), ("goal_dim", C.c_int32),
                ("hidden_dim", C.c_int32), ("feature_dim", C.c_int32), ("backward_hidden_dim", C.c_int32),
                ("use_goal", C.c_int32), ("rng_device", C.c_int32), ("contract_mode", C.c_int32), ("mlp_mode", C.c_int32),
                ("ortho_coef", C.c_float), ("mix_ratio", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float), ("future_ratio", C.c_float),
                ("seed", C.c_uint64), ("q_loss", C.c_int32), ("q_loss_coef", C.c_float),
                ("no_norm_z", C.c_int32), ("rand_weight", C.c_int32),
                ("add_trunk", C.c_int32), ("no_preprocess", C.c_int32),
                ("boltzmann", C.c_int32), ("temp", C.c_float), ("log_std_min", C.c_float), ("log_std_max", C.c_float),
                ("fused_stacks", C.c_int32), ("debug_identity_b", C.c_int32)]


class fb_host_storage(C.Structure):
    _fields_ = [("observation", C.c_void_p), ("action", C.c_void_p), ("reward", C.c_void_p), ("discount", C.c_void_p), ("goal", C.c_void_p),
                ("rows_per_episode", C.c_int32), ("obs_dim", C.c_int32), ("action_dim", C.c_int32), ("goal_dim", C.c_int32),
                ("max_episodes", C.c_int32)]


class fb_step_scalars(C.Structure):
    _fields_ = [("stddev", C.c_float), ("stddev_clip", C.c_float), ("lr_forward", C.c_float), ("lr_backward", C.c_float),
                ("lr_actor", C.c_float), ("tau", C.c_float), ("replay_discount", C.c_float), ("replay_future", C.c_float),
                ("grad_scale", C.c_float)]


class fb_buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_param_fb", "d_grad_fb", "d_m_fb", "d_v_fb", "d_target_fb", "d_param_actor",
                                          "d_grad_actor", "d_m_actor", "d_v_actor", "d_workspace")] + [("workspace_bytes", C.c_size_t)]


class fb_replay_view(C.Structure):
    _fields_ = [("d_rows", C.c_void_p), ("d_episode_len", C.c_void_p),
                ("max_episodes", C.c_int32), ("rows_per_episode", C.c_int32), ("row_stride", C.c_int32), ("n_episodes", C.c_int32),
                ("off_obs", C.c_int32), ("off_action", C.c_int32), ("off_reward", C.c_int32), ("off_discount", C.c_int32),
                ("off_goal", C.c_int32), ("off_extra", C.c_int32), ("goal_dim", C.c_int32), ("extra_dim", C.c_int32)]


_vp, _i, _u32, _f, _sz = C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_size_t
_pi32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes): every entry point include/fb_b200.h declares
SIGNATURES: tp.Dict[str, tp.Tuple[tp.Any, tp.List[tp.Any]]] = {
    "fb_abi_version": (_i, []),
    "fb_error_string": (C.c_char_p, [_i]),
    "fb_create": (_i, [C.POINTER(fb_config), C.POINTER(_vp)]),
    "fb_destroy": (None, [_vp]),
    "fb_flat_size": (_sz, [_vp, _i]),
    "fb_num_tensors": (_i, [_vp, _i]),
    "fb_tensor_info": (_i, [_vp, _i, _i, C.POINTER(_sz), C.POINTER(_i), C.POINTER(_i), C.c_char_p, _sz]),
    "fb_workspace_bytes": (_sz, [_vp]),
    "fb_bind": (_i, [_vp, C.POINTER(fb_buffers), _vp]),
    "fb_bind_replay": (_i, [_vp, C.POINTER(fb_replay_view), _vp]),
    "fb_set_step_scalars": (_i, [_vp, C.POINTER(fb_step_scalars), _vp]),
    "fb_set_indices": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fb_set_batch": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fb_upload_batch": (_i, [_vp, _vp, _i, _vp]),
    "fb_host_gather_rows": (_i, [C.POINTER(fb_host_storage), _vp, _vp, _vp, _i, _f, _vp, _i]),
    "fb_nccl_unique_id": (_i, [C.c_char_p, _vp]),
    "fb_nccl_init": (_i, [_vp, C.c_char_p, _vp, _i, _i]),
    "fb_p2p_create": (_i, [_vp, _i, _i, _vp, C.POINTER(fb_buffers)]),
    "fb_p2p_attach": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "fb_p2p_arena": (_vp, [_vp]),
    "fb_p2p_status": (_i, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), _vp]),
    "fb_p2p_slice": (_i, [_vp, _i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "fb_set_future_mask": (_i, [_vp, _vp, _vp]),
    "fb_set_mix_weights": (_i, [_vp, _vp, _vp, _vp]),
    "fb_set_z": (_i, [_vp, _vp, _vp]),
    "fb_set_noise": (_i, [_vp, _vp, _vp, _vp]),
    "fb_run": (_i, [_vp, _u32, _i, _vp]),
    "fb_launch_count": (_i, [_vp, _u32]),
    "fb_profile_ops": (_i, [_vp, _u32, _i, _vp, C.POINTER(C.c_float), _pi32, C.POINTER(C.c_double), C.POINTER(C.c_double), _i]),
    "fb_fused_profile": (_i, [_vp, _u32, _i, _vp, C.POINTER(C.c_float), _pi32, _i]),
    "fb_metrics_ptr": (_vp, [_vp]),
    "fb_set_adam_steps": (_i, [_vp, C.c_int64, C.c_int64, _vp]),
    "fb_get_adam_steps": (_i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _vp]),
    "fb_gather_block": (_i, [_vp, C.POINTER(_i), C.POINTER(_vp), C.POINTER(_vp)]),
    "fb_workspace_view": (_i, [_vp, C.c_char_p, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "fb_batch_row_layout": (_i, [_i, _i, _i, _i, _i, _pi32, _pi32]),
    "fb_replay_gather": (_i, [C.POINTER(fb_replay_view), _i, _i, _vp, _vp, _vp, _i, _f, _vp, _i, _vp]),
    "fb_replay_pack_episode": (_i, [C.POINTER(fb_replay_view), _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fb_sgemm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fb_gemm_tc_bench": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_float), _vp]),
    "fb_fp32_peak_tflops": (_i, [C.POINTER(C.c_double), _vp]),
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-ldl", "-lgomp"]


def nccl_library_path() -> tp.Optional[bytes]:
    """The NCCL shared object torch uses (the wheel's nvidia/nccl/lib/libnccl.so.2), for fb_nccl_*; None -> soname lookup."""
    try:
        import nvidia.nccl as n
        for d in n.__path__:
            p = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p.encode()
    except ImportError:
        pass
    return None


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/fb_b200.cu into libfb_b200.so next to this file (nvcc cross-compiles without a GPU)."""
    deps = SOURCES + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(p) for p in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, SOURCES[0]]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib: tp.Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared object and type every entry point.  Raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(controllable_agent_b200 has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype, fn.argtypes = res, args
    if lib.fb_abi_version() != FB_ABI_VERSION:
        raise RuntimeError("libfb_b200.so ABI version mismatch: rebuild it")
    _lib = lib
    return lib


class FBError(RuntimeError):
    pass


def check(code: int, what: str = "") -> None:
    if code != FB_OK:
        msg = load().fb_error_string(code).decode()
        raise FBError(f"libfb_b200 {what}: {msg} (code {code})")
