"""Where does a k_gemm_tc tile spend its time?  Runs the tcgen05 grouped GEMM on the step's problem shapes with the
profiling knobs of fb_gemm_tc_bench.  GPU only."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controllable_agent_b200 import _lib as L  # noqa: E402

NOBUILD, ONECHAIN, NOEPI, PRE_B, PRE_A, PAIR = 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 22


def main(pair: int = 0) -> None:
    lib = L.load()
    s = torch.cuda.current_stream().cuda_stream
    ms = C.c_float()
    shapes = [(1024, 1024, 1024, 128, 4), (1024, 1024, 1024, 128, 1), (1024, 512, 1024, 128, 5), (1024, 1024, 1024, 64, 1),
              (1024, 50, 1024, 64, 4), (1024, 50, 1024, 32, 4), (1024, 1024, 50 // 4 * 4 + 4, 128, 4), (4096, 4096, 4096, 128, 1)]
    for (M, N, K, bn, nprob) in shapes:
        row = []
        for dbg in (0, PRE_B, PRE_A | PRE_B, NOBUILD, NOBUILD | ONECHAIN, NOBUILD | ONECHAIN | NOEPI, NOEPI):
            L.check(lib.fb_gemm_tc_bench(M, N, K, bn, nprob, 1, dbg | pair, 20, C.byref(ms), s))
            row.append(ms.value * 1e3)
        fl = 2.0 * M * N * K * nprob
        tiles = -(-M // 128) * -(-N // bn) * nprob
        print(("pairs " if pair else "") + f"M={M} N={N} K={K} bn={bn} x{nprob} ({tiles} tiles): build A+B {row[0]:7.1f} us ({fl / row[0] / 1e6:6.1f} TF/s) | "
              f"pre-split B {row[1]:7.1f} ({fl / row[1] / 1e6:6.1f}) | pre-split A+B {row[2]:7.1f} ({fl / row[2] / 1e6:6.1f}) | nobuild {row[3]:7.1f} | "
              f"nobuild+1chain {row[4]:7.1f} | +noepi {row[5]:7.1f} | noepi {row[6]:7.1f}", flush=True)


def splitk_sweep() -> None:
    lib = L.load()
    s = torch.cuda.current_stream().cuda_stream
    ms = C.c_float()
    for (M, N, K, bn, nprob) in [(1024, 50, 1024, 64, 4), (1024, 50, 1024, 32, 4), (2048, 6, 1024, 32, 1), (50, 1024, 1024, 128, 3), (1024, 512, 1024, 128, 1)]:
        row = []
        for sk in (1, 2, 4, 8):
            L.check(lib.fb_gemm_tc_bench(M, N, K, bn, nprob, sk, PRE_B, 20, C.byref(ms), s))
            row.append(f"x{sk} {ms.value * 1e3:6.1f}")
        print(f"split-K M={M} N={N} K={K} bn={bn} x{nprob}: " + " | ".join(row) + " us", flush=True)


def data_sweep() -> None:
    """Constant operands vs pseudo-random ones, 20 back-to-back launches (x20) vs one launch between the events (x1)."""
    lib = L.load()
    s = torch.cuda.current_stream().cuda_stream
    ms = C.c_float()
    RANDOM = 1 << 21
    for (M, N, K, bn, nprob) in [(1024, 1024, 1024, 128, 4), (2048, 512, 1024, 128, 4), (1024, 1024, 1024, 128, 1), (1024, 1024, 1024, 128, 5)]:
        row = []
        for dbg in (PRE_B, PRE_B | RANDOM):
            for reps in (20, 1):
                L.check(lib.fb_gemm_tc_bench(M, N, K, bn, nprob, 1, dbg, reps, C.byref(ms), s))
                row.append(f"{ms.value * 1e3:6.1f}")
        print(f"data M={M} N={N} K={K} x{nprob}: constant fill x20 {row[0]} x1 {row[1]} | random fill x20 {row[2]} x1 {row[3]} us", flush=True)


if __name__ == "__main__":
    if "--quick" not in sys.argv:
        data_sweep()
        splitk_sweep()
    main()
    main(PAIR)
