"""Per-layer timing of the 3x3 conv kernels (tcgen05 v1/v2, mma.sync tile) at the BASELINE config-2 layer shapes.
usage: python tools/bench_conv.py [reps]   (run under ncu for counters)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cv_ssl_mis_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["row", "blk", "umma2", "tile"]
SHAPES = [(24, 256, 256, 16, 0, 16), (24, 256, 256, 16, 16, 16), (24, 128, 128, 16, 0, 32), (24, 128, 128, 32, 0, 32),
          (24, 128, 128, 32, 0, 64), (24, 128, 128, 32, 0, 16), (24, 64, 64, 32, 0, 64), (24, 64, 64, 64, 0, 64), (24, 64, 64, 64, 64, 64),
          (24, 32, 32, 64, 0, 128),
          (24, 32, 32, 128, 0, 128), (24, 16, 16, 256, 0, 256), (24, 32, 32, 128, 128, 128)]
for (n, h, w, c0, c1, cout) in SHAPES:
    d = ops.conv_desc(n, 1, h, w, c0, c1, cout, 3, 1, 1, 2)
    M, cin = n * h * w, c0 + c1
    x0 = torch.randn(M, c0, device="cuda")
    x1 = torch.randn(M, c1, device="cuda") if c1 else None
    wgt = torch.randn(cout, cin, 3, 3, device="cuda") * (cin * 9) ** -0.5
    bias = torch.zeros(cout, device="cuda")
    y = torch.empty(M, cout, device="cuda")
    flops = 2.0 * M * 9 * cin * cout
    byts = 4.0 * (M * cin + M * cout)
    line = f"{n}x{h}x{w} {cin:3d}->{cout:3d}: "
    for name in which:
        if name == "row":
            if not ops.conv_row_supported(d, False):
                continue
            wt = torch.empty(ops.conv_row_packed_floats(d, False), device="cuda")
            ops.conv_row_pack_weights(d, False, wgt, wt)
            part = torch.empty(ops.conv_row_stats_blocks(d) * 2 * cout, dtype=torch.float64, device="cuda")
            fn = lambda: ops.conv_row_fwd(d, x0, x1, wt, bias, y, part)
        elif name == "blk":
            if not ops.conv_blk_supported(d, False):
                continue
            wt = torch.empty(9 * cout * cin, device="cuda")
            ops.conv_blk_pack_weights(wgt, wt, False, cout, cin)
            part = torch.empty(ops.conv_blk_stats_blocks(d) * 2 * cout, dtype=torch.float64, device="cuda")
            fn = lambda: ops.conv_blk_fwd(d, x0, x1, wt, bias, y, part)
        elif name.startswith("umma"):
            wt = torch.empty(ops.conv_umma_packed_floats(False, cout, cin, 9), device="cuda")
            ops.conv_umma_pack_weights(wgt, wt, False, cout, cin, 9)
            fn = lambda: ops.conv_umma_fwd(d, x0, x1, wt, bias, y)
        else:
            wt = torch.empty(ops.conv_tile_packed_floats(False, cout, cin, 9), device="cuda")
            ops.conv_tile_pack_weights(wgt, wt, False, cout, cin, 9)
            fn = lambda: ops.conv_tile_fwd(d, x0, x1, wt, bias, y)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()                       # graph replay: no host launch gaps in the timing
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        line += f" {name} {us:7.1f} us {flops / us / 1e6:6.1f} TF {byts / us / 1e3:6.0f} GB/s |"
    print(line)
