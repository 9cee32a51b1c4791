import torch, sys
sys.path.insert(0,'.')
from biomedkg_b200 import ops
from biomedkg_b200._cabi import lib
h1=torch.randn(300,128,device='cuda',requires_grad=True); h2=torch.randn(300,128,device='cuda',requires_grad=True)
l=ops.infonce_loss(h1,h2,0.2)
try:
    l.backward(); torch.cuda.synchronize(); print('autograd bwd ok', float(h1.grad.abs().sum()))
except Exception as e:
    print('autograd bwd failed:', e, 'status', lib.bmkg_last_driver_status())
