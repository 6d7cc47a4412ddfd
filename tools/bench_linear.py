"""Microbenchmark of the tcgen05 Linear kernels on the Swin-UNet shapes (bs16): us, TF/s, GB/s per product,
next to torch.matmul (cuBLAS TF32).  usage: python tools/bench_linear.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cv_ssl_mis_b200 import ops

B = 16
SHAPES = [("qkv1", B * 3136, 288, 96), ("fc1_1", B * 3136, 384, 96), ("fc2_1", B * 3136, 96, 384), ("qkv2", B * 784, 576, 192),
          ("fc1_2", B * 784, 768, 192), ("fc1_3", B * 196, 1536, 384), ("fc2_3", B * 196, 384, 1536), ("fc1_4", B * 49, 3072, 768),
          ("exp4", B * 3136, 1536, 96)]


def timeit(fn, n=20):
    """device time per call: n calls captured in one CUDA graph (no host launch overhead), replayed 3 times"""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n) * 1e3


torch.backends.cuda.matmul.allow_tf32 = True
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, M, O, I in SHAPES:
    if only and name != only:
        continue
    x = torch.randn(M, I, device="cuda"); w = torch.randn(O, I, device="cuda"); b = torch.randn(O, device="cuda")
    dy = torch.randn(M, O, device="cuda"); y = torch.empty(M, O, device="cuda"); dx = torch.empty(M, I, device="cuda")
    dw = torch.empty(O, I, device="cuda")
    ws = torch.empty(max(ops.linear_wgrad_workspace_bytes(M, O, I) // 4, 4), device="cuda")
    if only:                                   # single launches for ncu
        ops.linear_fwd(x, None, w, b, y, M, O); ops.linear_dgrad(dy, w, dx, None, False, M, O)
        ops.linear_wgrad(x, None, dy, dw, ws, M, O); ops.colsum(dy, M, O, b, ws2 := torch.empty(ops.colsum_workspace_bytes(M, O) // 4 + 4, device="cuda"))
        torch.cuda.synchronize()
        continue
    fl = 2.0 * M * O * I
    by_f = 4.0 * (M * I + M * O + O * I)
    t = {"fwd": timeit(lambda: ops.linear_fwd(x, None, w, b, y, M, O)),
         "dgrad": timeit(lambda: ops.linear_dgrad(dy, w, dx, None, False, M, O)),
         "wgrad": timeit(lambda: ops.linear_wgrad(x, None, dy, dw, ws, M, O)),
         "cublas_fwd": timeit(lambda: torch.addmm(b, x, w.t(), out=y)),
         "cublas_wgrad": timeit(lambda: torch.mm(dy.t(), x, out=dw))}
    print(f"{name:6s} M={M:6d} O={O:4d} I={I:4d} | " + " | ".join(
        f"{k} {v:7.1f}us {fl / v * 1e-6:6.1f}TF {by_f / v * 1e-3:6.0f}GB/s" for k, v in t.items()), flush=True)
