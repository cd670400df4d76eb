import os, sys
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
from anyref_b200 import ops
M=65536
x = torch.randn(M,1280,device="cuda")*2+0.3
g = torch.rand(1280,device="cuda")+0.5; b=torch.randn(1280,device="cuda")*0.1
out=torch.empty(M,1280,device="cuda",dtype=torch.bfloat16)
ref=F.layer_norm(x,(1280,),g,b,1e-6)
ops.layernorm(x,g,b,1e-6,torch.bfloat16,out=out)
print("err", (out.float()-ref).abs().max().item())
o32=ops.layernorm(x,g,b,1e-6,torch.float32)
print("err32", (o32-ref).abs().max().item())
for _ in range(10): ops.layernorm(x,g,b,1e-6,torch.bfloat16,out=out)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): ops.layernorm(x,g,b,1e-6,torch.bfloat16,out=out)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/100
print(f"LN {ms*1e3:.1f} us  {M*1280*6/ms/1e6:.0f} GB/s")
