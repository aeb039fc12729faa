"""Cycle breakdown of the row-GEMM epilogue (needs a library built with -DVKN_EPI_PROF: tools/epi_prof.sh).
Counters are summed over thread 64 (epilogue warp 0, lane 0) of every CTA of the chain launches of one stage."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), ROOT]
import torch  # noqa: E402
import bench  # noqa: E402
import vknet  # noqa: E402
from vknet import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, 'tools', 'probes', 'libvknet_prof.so')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda:0')
torch.manual_seed(0)
h = vknet.build_head(dict(type='KernelUpdateHead', **bench.head_cfg()))
h.init_weights()
h = h.to(dev).bfloat16().eval()
xs, pfs, ms = zip(*[bench.dummy_inputs(torch, s) for s in range(min(B, 8))])
rep = (B + len(xs) - 1) // len(xs)
x = torch.cat(xs).repeat(rep, 1, 1, 1)[:B].to(dev).bfloat16()
pf = torch.cat(pfs).repeat(rep, 1, 1)[:B].to(dev)
m = torch.cat(ms).repeat(rep, 1, 1, 1)[:B].to(dev).bfloat16()
for _ in range(3):
    h(x, pf, m)
torch.cuda.synchronize()
L = _lib.lib()
out = (C.c_ulonglong * 16)()
L.vkn_debug_epi_prof(None, 1)
h(x, pf, m)
torch.cuda.synchronize()
L.vkn_debug_epi_prof(out, 0)
names = ['tmem ld', 'add', 'fp32 staging wait', 'plane staging wait', 'post', 'chunk store (all)', '-', 'LN barrier',
         'post: sigmoid', 'post: mul', 'post: add2', 'store: fp32 sts+fence+tma', 'store: plane split', 'store: split..plane tma (cumulative)', '-', '-']
ncta = (B * 100 + 255) // 256
for n, v in zip(names, out):
    print('%-22s %10.1f us per CTA' % (n, v / 1.9e3 / ncta))
