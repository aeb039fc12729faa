import os, sys, torch
sys.path[:0]=['/root/repo/video-k-net_b200','/root/repo/oracle','/root/repo/tests','/root/repo']
import knet_oracle as ko
from helpers import build_heads
dev=torch.device('cuda:0')
os.environ['VKN_ROWS_TC_MIN']='1'
B,N,C,H,W,S=int(os.environ.get("DBG_B","64")),100,256,200,88,2
cfg=ko.default_cfg(num_classes=19,in_channels=C,feedforward_channels=2048)
sds=[ko.round_state_dict_bf16(ko.random_state_dict(cfg,seed=50+s)) for s in range(S)]
heads=build_heads('KernelUpdateHead',cfg,sds,dev,dtype=torch.bfloat16)
g=torch.Generator(device=dev).manual_seed(7)
x=torch.randn(B,C,H,W,generator=g,device=dev); pfd=torch.randn(B,N,C,generator=g,device=dev)
mb=pfd.bmm(x.view(B,C,-1)).view(B,N,H,W).bfloat16(); xb=x.bfloat16()
res={}
for mode in ('0','1'):
    os.environ['VKN_MASK_F16']=mode
    obj,m=pfd,mb; outs=[]
    for h in heads:
        cls,m,obj=h(xb,obj,m); outs.append((cls.clone(),m.clone(),obj.clone()))
    res[mode]=outs
for s in range(S):
    a,b=res['0'][s][1].float(),res['1'][s][1].float()
    d=(a-b).abs()
    print('stage',s,'obj equal',torch.equal(res['0'][s][2],res['1'][s][2]),'max diff',d.max().item(),'scale',a.abs().max().item(),'n diff',int((d>0).sum()), 'n big', int((d>2**-6*a.abs().clamp_min(1)).sum()))
    if d.max()>0.5:
        idx=(d==d.max()).nonzero()[0].tolist(); print('  at',idx, a[tuple(idx)].item(), b[tuple(idx)].item())
        bb,nn=idx[0],idx[1]
        print('  per-kernel max diff for that frame:', d[bb].amax(dim=(1,2))[:12].tolist())
        print('  per-frame max diff:', d.amax(dim=(1,2,3)).tolist())
from helpers import bf16_ulp
for mode in ('0','1'):
    outs=res[mode]
    for s in range(S):
        obj_in = pfd if s==0 else outs[s-1][2].reshape(B,N,C)
        m_in = mb if s==0 else outs[s-1][1]
        nf=B if s==1 else 2
        want=ko.kernel_update_head_forward(sds[s],cfg,xb[:nf].float().cpu(),obj_in[:nf].float().cpu().reshape(nf,N,C,1,1),m_in[:nf].float().cpu())
        got=outs[s][1][:nf].float().cpu(); ref=ko.round_bf16(want[1])
        fl=2.0**-16*ref.abs().max().item()
        err=(got-ref).abs(); tol=bf16_ulp(torch.maximum(ref.abs(),got.abs()))*(1+1e-6)+fl
        r=(err/tol)
        print('mode',mode,'stage',s,'worst',r.max().item(),'n>1',int((r>1).sum()),'obj err',(outs[s][2][:nf].cpu().reshape(want[2].shape)-want[2]).abs().max().item())
        if r.max()>1:
            idx=(r==r.max()).nonzero()[0].tolist(); print('   at',idx,got[tuple(idx)].item(),ref[tuple(idx)].item(), want[1][tuple(idx)].item())
            bad=(r>1).nonzero(); print('   kernels hit',sorted(set(bad[:,1].tolist()))[:20],'frames',sorted(set(bad[:,0].tolist())))
            px=(bad[:,2]*W+bad[:,3]); print('   pixel tiles hit', sorted(set((px//128).tolist()))[:30], 'n', len(bad))
m0=res['1'][0][1].float()
sl=(m0>0)&(m0<2e-7)
print('sliver logits in mode-1 stage-0 output:', int(sl.sum()), sl.nonzero()[:5].tolist(), m0[sl][:5].tolist(), 'sigmoid>0.5:', (torch.sigmoid(m0[sl].cpu())>0.5).tolist()[:5])
m0=res['0'][0][1].float(); sl=(m0>0)&(m0<2e-7); print('mode-0:', int(sl.sum()))
