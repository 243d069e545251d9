"""Developer probe: the M=256 decode GEMMs (library code: cuBLAS through F.linear) with the weights
streaming from HBM (rotating over copies > L2), default heuristics vs torch TunableOp's per-shape choice.
  python tools/gemm_probe.py            -> one JSON line per shape
"""
import json
import sys

import torch
import torch.nn.functional as F

SHAPES = {   # name: (M, N, K)
    "0.6b qkv": (256, 4096, 1024), "0.6b o": (256, 1024, 2048), "0.6b gate_up": (256, 6144, 1024),
    "0.6b down": (256, 1024, 3072),
    "32b/8 qkv": (256, 1280, 5120), "32b/8 o": (256, 5120, 1024), "32b/8 gate_up": (256, 6400, 5120),
    "32b/8 down": (256, 5120, 3200),
}


def bench(M, N, K, iters=60):
    copies = max(2, int(300e6 // (N * K * 2)) + 1)
    ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) for _ in range(copies)]
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    for w in ws[:3]:
        F.linear(x, w)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):                  # as in the decode step: no launch overhead in the number
        for i in range(iters):
            F.linear(x, ws[i % copies])
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    base = {n: bench(*s) for n, s in SHAPES.items()}
    tuned = {}
    try:
        import torch.cuda.tunable as tunable
        tunable.enable(True)
        tunable.tuning_enable(True)
        tunable.set_max_tuning_iterations(30)
        for n, (M, N, K) in SHAPES.items():      # tune eagerly (graph capture must not tune)
            x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
            w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
            F.linear(x, w)
        torch.cuda.synchronize()
        tunable.tuning_enable(False)
        tuned = {n: bench(*s) for n, s in SHAPES.items()}
    except Exception as e:
        print(json.dumps({"tunable_op": f"unavailable: {type(e).__name__}: {e}"[:300]}))
    for n, (M, N, K) in SHAPES.items():
        gb = N * K * 2 / 1e9
        print(json.dumps({"gemm": n, "M": M, "N": N, "K": K, "us_default": round(base[n], 2),
                          "weight_GBps_default": round(gb / base[n] * 1e6, 0),
                          "us_tunable": round(tuned[n], 2) if n in tuned else None}))


if __name__ == "__main__":
    sys.exit(main())
