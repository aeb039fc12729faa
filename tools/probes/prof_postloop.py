import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[os.path.join(ROOT,'video-k-net_b200'), os.path.join(ROOT,'oracle'), ROOT]
import torch
from vknet import ops, _lib
dev=torch.device('cuda:0')
g=torch.Generator().manual_seed(4)
K,M,H,W=100,17,375,1242
base=torch.rand(K+M,H//15+1,W//18+1,generator=g)*0.3
masks=torch.nn.functional.interpolate(base[None],size=(H,W),mode='bilinear',align_corners=False)[0]
for k in range(K+M):
    y0,x0=int(torch.randint(0,H-60,(1,),generator=g)),int(torch.randint(0,W-200,(1,),generator=g))
    hh,ww=int(torch.randint(20,60,(1,),generator=g)),int(torch.randint(40,200,(1,),generator=g))
    masks[k,y0:y0+hh,x0:x0+ww]=0.55+0.45*torch.rand(hh,ww,generator=g)
scores=0.2+0.8*torch.rand(K+M,generator=g)
labels=torch.cat([torch.randint(0,2,(K,),generator=g),torch.arange(M)+2])
md,sd,ld=masks.to(dev),scores.to(dev),labels.to(dev)
f=lambda: ops.panoptic_merge(md[:K],ld[:K],sd[:K],md[K:],ld[K:],sd[K:],2,0.3,0.5)
for _ in range(3): f()
with _lib.profile() as p: f()
print('panoptic',[(n,round(t*1e3,1)) for n,t in p.records])
n=100; C=256
xy=torch.rand(n,2,generator=g)*300
bb=torch.cat([xy,xy+40+60*torch.rand(n,2,generator=g),torch.rand(n,1,generator=g)],1).to(dev)
tl=torch.randint(0,2,(n,),generator=g).to(dev); emb=torch.randn(n,C,generator=g).to(dev)
memo=(torch.randint(0,2,(30,),generator=g).to(dev),torch.randn(30,C,generator=g).to(dev),torch.arange(30).to(dev))
f2=lambda: ops.track_match(bb,tl,emb,memo[0],memo[1],memo[2],30,0.3,0.5,0.35,0.5,0.3,0.7,True)
for _ in range(3): f2()
with _lib.profile() as p: f2()
print('track',[(n_,round(t*1e3,1)) for n_,t in p.records])
import time
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(20): f2()
torch.cuda.synchronize(); print('track wall us', (time.perf_counter()-t0)/20*1e6)
