import os, sys
sys.path.insert(0, "/root/repo")
import torch
from tests import gpu_ops as G
from tests.gpu_ops import DT, _p, _s, check, lib
n=int(sys.argv[1]) if len(sys.argv)>1 else 1024
cin,cout,h=48,80,56
x=torch.randn(n,cin,h,h,device="cuda"); w=(torch.randn(cout,cin,3,3,device="cuda")*0.05).contiguous(); b=torch.zeros(cout,device="cuda")
X=G.PF8.from_nchw(x,"bf16"); O=G.PF8(n,cout,h,h,"bf16")
nb=int(lib().mil_conv_workspace_bytes(n,cin,h,h,cout,h,h,3)); ws=torch.zeros(nb,dtype=torch.uint8,device="cuda")
def conv(epi,bias):
    check(lib().mil_conv_pf8(DT["bf16"],2,0,_p(X.buf),n,cin,h,h,_p(w),cout,cin,3,1,_p(bias),None,None,_p(O.buf),h,h,epi,_p(ws),nb,_s()),"conv")
for name,epi,bias in (("fwd",0,b),("plain",2,None)):
    for _ in range(3): conv(epi,bias)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): conv(epi,bias)
    e1.record(); torch.cuda.synchronize()
    us=e0.elapsed_time(e1)/10*1e3
    byt=n*h*h*(cin+cout)*2
    print(name, n, "tiles", round(us,1),"us", round(byt/us/1e6,2),"TB/s algorithmic")
