import sys, torch
sys.path[:0]=['/root/repo/video-k-net_b200','/root/repo/oracle','/root/repo']
import knet_oracle as ko
from vknet import ops
dev=torch.device('cuda:0')
g = torch.Generator().manual_seed(4)
K, M, H, W, nthing = 100, 17, 375, 1242, 2
base = torch.rand(K + M, H // 15 + 1, W // 18 + 1, generator=g)
masks = torch.nn.functional.interpolate(base[None], size=(H, W), mode='bilinear', align_corners=False)[0]
masks[5] = masks[4]
scores = torch.rand(K + M, generator=g)
scores[5] = scores[4]
labels = torch.cat([torch.randint(0, nthing, (K,), generator=g), torch.arange(M) + nthing])
want_seg, want_info, want_kept = ko.panoptic_merge_joint(masks[:K], labels[:K], scores[:K], masks[K:], labels[K:], scores[K:], nthing, 0.3, 0.5)
seg, info, kept = ops.panoptic_merge(masks[:K].to(dev), labels[:K].to(dev), scores[:K].to(dev), masks[K:].to(dev), labels[K:].to(dev), scores[K:].to(dev), nthing, 0.3, 0.5)
d = seg.cpu()!=want_seg
print('mismatch px', int(d.sum()), 'nseg', len(info), len(want_info), kept==want_kept)
owner_ref=(scores.view(-1,1,1)*masks).argmax(0)
print(info[:3]); print(want_info[:3])
for a,b in zip(info,want_info):
    if {k:v for k,v in a.items() if k!='score'}!={k:v for k,v in b.items() if k!='score'}: print('diff',a,b); break
