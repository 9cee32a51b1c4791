import torch, sys, math
sys.path.insert(0,'.')
from biomedkg_b200 import ops
from biomedkg_b200.ops import _p, _stream, call, lib, _ws
for (N,D) in [(200,128),(5,64)]:
    h1=torch.randn(N,D,device='cuda'); h2=torch.randn(N,D,device='cuda')
    scale=math.sqrt(ops.LOG2E/0.2)
    z=torch.empty(2*N,D,dtype=torch.bfloat16,device='cuda'); inv=torch.empty(2*N,device='cuda')
    call("bmkg_l2norm_scale",_p(h1),N,D,scale,_p(z),_p(inv),_stream())
    call("bmkg_l2norm_scale",_p(h2),N,D,scale,z.data_ptr()+N*D*2,inv.data_ptr()+N*4,_stream())
    loss=torch.empty((),device='cuda'); inv_r=torch.empty(lib.bmkg_infonce_padded_rows(N),device='cuda')
    ws=_ws(lib.bmkg_infonce_workspace_bytes(N,D),'cuda')
    call("bmkg_infonce_fwd",_p(z),N,D,_p(loss),_p(inv_r),_p(ws),ws.numel(),_stream())
    torch.cuda.synchronize(); print(N,D,'fwd loss',float(loss))
    from oracle import pygcl
    print('ref', float(pygcl.infonce_l2l_closed_form(h1.cpu().double(),h2.cpu().double(),0.2)))
    g=torch.ones((),device='cuda'); dz=torch.empty(2*N,D,device='cuda')
    try:
        call("bmkg_infonce_bwd",_p(z),_p(inv_r),_p(g),N,D,_p(dz),_stream()); torch.cuda.synchronize(); print('bwd main thread ok', float(dz.abs().sum()))
    except Exception as e: print('bwd main thread FAILED', e)
