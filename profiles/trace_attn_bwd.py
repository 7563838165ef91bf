import sys, os, ctypes
sys.path.insert(0, "/root/repo")
import torch
from musicstyletransfer_b200 import ops, lib
dev="cuda"
B,T,H,dh=2048,65,8,32
qkv=torch.randn(B*T,3*H*dh,device=dev); mask=torch.ones(B*T,device=dev)
dctx=torch.randn(B*T,H*dh,device=dev); dqkv=torch.empty_like(qkv); db=torch.zeros(3*H*dh,device=dev)
tr=torch.zeros(2*16*9,dtype=torch.int64,device=dev)
L=lib.load()
for use_db in (True, False):
    for _ in range(2): ops.attention_tc_bwd(qkv,mask,dctx,dqkv,B,T,H,dh,dbias=db if use_db else None)
    L.msx_attention_tc_set_trace(ctypes.c_void_p(tr.data_ptr()))
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); ops.attention_tc_bwd(qkv,mask,dctx,dqkv,B,T,H,dh,dbias=db if use_db else None); e1.record()
    torch.cuda.synchronize()
    L.msx_attention_tc_set_trace(None)
    print("dbias",use_db,"kernel ms",e0.elapsed_time(e1))
    t=tr.cpu().view(2,16,9)
    for g in range(2):
        print(" group",g)
        for n in range(2,8):
            r=t[g,n]; base=int(r[0])
            print("  item",n,"top@%d"%(int(r[0])-int(t[g,2,0])),[int(r[i])-base for i in range(1,9)])
