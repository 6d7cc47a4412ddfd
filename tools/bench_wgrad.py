"""Per-layer timing of the 3x3 weight-gradient kernels (row-ring tcgen05 vs the mma.sync tile kernel) at the BASELINE
config-2 layer shapes.  usage: python tools/bench_wgrad.py [reps] [row,tile]   (run under ncu for counters)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cv_ssl_mis_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["row", "tile"]
SHAPES = [(24, 256, 256, 16, 0, 16), (24, 256, 256, 16, 16, 16), (24, 128, 128, 32, 0, 32), (24, 128, 128, 32, 32, 32),
          (24, 64, 64, 32, 0, 64), (24, 64, 64, 64, 0, 64), (24, 64, 64, 64, 64, 64), (24, 32, 32, 64, 0, 128),
          (24, 32, 32, 128, 0, 128), (24, 32, 32, 128, 128, 128), (24, 16, 16, 128, 0, 256), (24, 16, 16, 256, 0, 256)]
flush = torch.empty(64 * 1024 * 1024, device="cuda")       # 256 MB > L2
for (n, h, w, c0, c1, cout) in SHAPES:
    d = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    M, cin = n * h * w, c0 + c1
    x0 = torch.randn(M, c0, device="cuda")
    x1 = torch.randn(M, c1, device="cuda") if c1 else None
    dy = torch.randn(M, cout, device="cuda")
    dw = torch.empty(cout, cin, 3, 3, device="cuda")
    db = torch.empty(cout, device="cuda")
    flops = 2.0 * M * 9 * cin * cout
    byts = 4.0 * (M * cin + M * cout)
    line = f"{n}x{h}x{w} {cin:3d}->{cout:3d}: "
    res = {}
    for name in which:
        if name == "row":
            if not ops.conv_row_wgrad_supported(d):
                continue
            ws = torch.empty(ops.conv_row_wgrad_workspace_bytes(d) // 4 + 4, device="cuda")
            fn = lambda: ops.conv_row_wgrad(d, x0, x1, dy, ws, dw)
        else:
            ws = torch.empty(ops.conv_tile_wgrad_workspace_bytes(d) // 4 + 4, device="cuda")
            fn = lambda: ops.conv_tile_wgrad(d, x0, x1, dy, ws, dw, db)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_(); flush.zero_()          # > L2, and long enough for the host to queue the launches behind it
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        us = tot / reps * 1e3
        res[name] = dw.clone()
        line += f" {name} {us:7.1f} us {flops / us / 1e6:6.1f} TF {byts / us / 1e3:6.0f} GB/s |"
    if len(res) == 2:
        a, b = res["row"], res["tile"]
        line += f" maxdiff/scale {float((a - b).abs().max() / b.abs().max()):.2e}"
    print(line, flush=True)
