"""Per-layer timing of the 3x3x3 convolution kernels (halo-block tcgen05 vs mma.sync tile) at the VNet layer shapes of
BASELINE config 4 (112 x 112 x 80 patch, 16 filters).  usage: python tools/bench_conv3d.py [reps] [blk,tile] [fwd|dgrad]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cv_ssl_mis_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["blk", "tile"]
dgrad = len(sys.argv) > 3 and sys.argv[3] == "dgrad"
SHAPES = [(4, 96, 96, 96, 16), (4, 48, 48, 48, 32), (4, 56, 56, 40, 32), (4, 28, 28, 20, 64), (4, 14, 14, 10, 128), (4, 7, 7, 5, 256), (2, 56, 56, 40, 32), (2, 28, 28, 20, 64)]
for (n, dd, h, w, c) in SHAPES:
    d = ops.conv_desc(n, dd, h, w, c, 0, c, 3, 1, 1, 3)
    M = n * dd * h * w
    x0 = torch.randn(M, c, device="cuda")
    wgt = torch.randn(c, c, 3, 3, 3, device="cuda") * (c * 27) ** -0.5
    bias = torch.zeros(c, device="cuda")
    y = torch.empty(M, c, device="cuda")
    flops, byts = 2.0 * M * 27 * c * c, 8.0 * M * c
    line = f"{n}x{dd}x{h}x{w} {c:3d}->{c:3d}: "
    for name in which:
        if name == "blk":
            if not ops.conv_blk_supported(d, dgrad):
                continue
            wt = torch.empty(27 * c * c, device="cuda")
            ops.conv_blk_pack_weights(wgt, wt, ops.conv_blk_supported(d, dgrad) - 8, c, c, 27)
            part = torch.empty(ops.conv_blk_stats_blocks(d) * 2 * c, dtype=torch.float64, device="cuda")
            fn = (lambda: ops.conv_blk_dgrad(d, x0, wt, y, None, False)) if dgrad else (lambda: ops.conv_blk_fwd(d, x0, None, wt, bias, y, part))
        else:
            wt = torch.empty(ops.conv_tile_packed_floats(dgrad, c, c, 27), device="cuda")
            ops.conv_tile_pack_weights(wgt, wt, dgrad, c, c, 27)
            fn = (lambda: ops.conv_tile_dgrad(d, x0, wt, y, None, False)) if dgrad else (lambda: ops.conv_tile_fwd(d, x0, None, wt, bias, y))
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()
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
    print(line, flush=True)
