"""Pair mode vs the 64-byte-row mode of csrc/conv_row.cu on the same inputs: outputs, statistics and data gradients must agree to fp32 round-off."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_ssl_mis_b200 import ops
torch.manual_seed(0)
for (n,h,w,c0,c1,cout) in [(24,256,256,16,0,16),(12,256,256,16,16,16)]:
    d = ops.conv_desc(n,1,h,w,c0,c1,cout,3,1,1,2)
    M=n*h*w; cin=c0+c1
    x0 = torch.randn(M,c0, device='cuda'); x1 = torch.randn(M,c1,device='cuda') if c1 else None
    wgt = torch.randn(cout,cin,3,3,device='cuda')*(cin*9)**-0.5; bias=torch.randn(cout,device='cuda'); dy=torch.randn(M,cout,device='cuda')
    res={}
    for nopair in ('0','1'):
        os.environ['B200_ROW_NOPAIR']=nopair
        wt = torch.empty(ops.conv_row_packed_floats(d,False),device='cuda'); ops.conv_row_pack_weights(d,False,wgt,wt)
        y=torch.full((M,cout),float('nan'),device='cuda'); nb=ops.conv_row_stats_blocks(d)
        part=torch.zeros(nb*2*cout,dtype=torch.float64,device='cuda')
        ops.conv_row_fwd(d,x0,x1,wt,bias,y,part)
        wb = torch.empty(ops.conv_row_packed_floats(d,True),device='cuda'); ops.conv_row_pack_weights(d,True,wgt,wb)
        dx0=torch.full((M,c0),float('nan'),device='cuda'); dx1=torch.full((M,c1),float('nan'),device='cuda') if c1 else None
        ops.conv_row_dgrad(d,dy,wb,dx0,dx1)
        torch.cuda.synchronize()
        res[nopair]=(y,part.view(nb,2,cout).sum(0),dx0,dx1)
    a,b=res['0'],res['1']
    print((n,h,w,c0,c1,cout),'y maxdiff',float((a[0]-b[0]).abs().max()),'nan',int(torch.isnan(a[0]).sum()),
          'stats rel',float(((a[1]-b[1]).abs()/b[1].abs()).max()),'dx0',float((a[2]-b[2]).abs().max()),
          'dx1',float((a[3]-b[3]).abs().max()) if c1 else None)
    bad=(a[0]-b[0]).abs().max(1).values>1e-4
    if bad.any():
        idx=bad.nonzero().flatten()
        pix=idx%(h*w); print('  bad pixels',int(bad.sum()),'rows(h) sample',sorted(set((pix//w).tolist()))[:20],'cols sample',sorted(set((pix%w).tolist()))[:20])
    bad=(a[2]-b[2]).abs().max(1).values>1e-4
    if bad.any():
        idx=bad.nonzero().flatten()
        pix=idx%(h*w); print('  bad dx pixels',int(bad.sum()),'rows(h) sample',sorted(set((pix//w).tolist()))[:20],'cols sample',sorted(set((pix%w).tolist()))[:20])
