"""Which side bounds the conv mainloop?  Time a layer with the TMA loads and/or the MMAs stubbed out
(isb_conv_desc.debug_flags).  dbg0 = normal, dbg1 = no TMA, dbg2 = no MMA, dbg3 = neither (pure fixed cost:
launch, prologue, barrier hand-shakes, epilogue).  Result of round 1: dbg3 is ~60 % of dbg0 on every layer."""
import sys, os
sys.path.insert(0,'/root/repo')
import torch
from tools.tune_conv import bench
from ishapediting_b200.ops import CudaOps
ops=CudaOps(torch.device("cuda",0),"bf16")
dev=ops.device
for (H,k,Cin,Cout,bn,sp) in [(128,3,256,256,128,1),(128,3,256,256,256,1),(8,3,1024,1024,128,8),(32,3,512,512,128,4)]:
    K=k*k*Cin
    a=torch.randn(1,H,H,Cin,device=dev).to(torch.bfloat16); out=torch.empty(1,H,H,Cout,device=dev); bias=torch.randn(Cout,device=dev)
    n=max(4,min(32,int(300e6//(Cout*K*2))))
    ws=[torch.randn(Cout,K,device=dev).to(torch.bfloat16) for _ in range(n)]
    res=[]
    for dbg in (0,1,2,3):
        for stg in (3,6):
            t,_=bench(ops,a,ws,bias,k,out,{"block_n":bn,"split_k":sp,"stages":stg,"two_cta":2,"debug":dbg},False)
            res.append(f"dbg{dbg}/st{stg}: {t:.1f}")
    print(f"H={H} k{k} {Cin}->{Cout} bn={bn} split={sp}: "+"  ".join(res), flush=True)
