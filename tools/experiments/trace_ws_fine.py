import ctypes, os, sys
ROOT="/root/repo"; PKG=ROOT+"/flash-attention-v2-rdna3-minimal_b200"
os.environ["FA_FWD_SM100_LIB"]=PKG+"/lib/libfa_fwd_sm100_trace.so"
sys.path.insert(0,PKG)
import torch, numpy as np
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction
_capi.set_kernel(_capi.FA_KERNEL_WS)
q,k,v=(torch.rand(1,16,16384,128,dtype=torch.float16,device="cuda") for _ in range(3))
buf=torch.zeros(5*128*8,dtype=torch.int64,device="cuda")
for _ in range(2): FlashAttentionFunction.apply(q,k,v,None,False)
torch.cuda.synchronize()
_capi.lib.fa_trace_set.argtypes=[ctypes.c_void_p]; _capi.lib.fa_trace_set(buf.data_ptr())
FlashAttentionFunction.apply(q,k,v,None,False); torch.cuda.synchronize()
t=buf.cpu().view(5,128,8).numpy().astype(np.int64)
m,f=t[2],t[3]
for j in range(60,64):
    b=m[j,0]
    print(j,"mma",[int(x-b) for x in m[j]],"fine[S1 mma issued, s_full commit, after issue_s, before relV, after relV]",[int(f[j,i]-b) for i in (0,1,2,3,4)],"next0",int(m[j+1,0]-b))
