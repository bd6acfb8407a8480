import torch, sys, os
sys.path.insert(0,os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helping_hand_for_egocentric_videos_b200 import _lib as L
B,T,n,H=16,16,256,16
qkv=(torch.randn(B*(1+T*n),3*H*64,device='cuda')*0.5).bfloat16(); o=torch.empty(B*(1+T*n),H*64,device='cuda',dtype=torch.bfloat16)
lib=L.load()
L.check(lib.hh_attention(L.ptr(qkv),L.ptr(o),B,T,n,H,0,L.stream_ptr()),'a'); torch.cuda.synchronize()
