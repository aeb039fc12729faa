"""Diagnostic for the tcgen05/TMA engines: runs each engine on seeded inputs and prints error tables
against a float64 torch reference computed on the GPU (no asserts; run each engine in its own process
so that a device-side trap in one does not poison the other).

    python tools/tc_debug.py pool|mask|stage [B N C H W]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-k-net_b200'), os.path.join(ROOT, 'oracle')]

import torch  # noqa: E402

import knet_oracle as ko  # noqa: E402
import vknet  # noqa: E402
from vknet import _lib, ops  # noqa: E402


def build(C, Fh, dev):
    cfg = ko.default_cfg(num_classes=19, in_channels=C, feedforward_channels=Fh)
    sd = ko.round_state_dict_bf16(ko.random_state_dict(cfg, seed=5))
    h = vknet.build_head(dict(type='KernelUpdateHead', **cfg))
    h.load_state_dict(sd, strict=True)
    return cfg, sd, h.to(dev).bfloat16().eval()


def report(name, got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    err = (got - ref).abs()
    print('%-28s max|err| %.3e  mean|err| %.3e  max|ref| %.3e  rel %.3e' % (
        name, err.max().item(), err.mean().item(), ref.abs().max().item(),
        err.max().item() / max(ref.abs().max().item(), 1e-30)), flush=True)
    if err.max().item() > 1e-2 * ref.abs().max().item():
        idx = torch.nonzero(err > 1e-2 * ref.abs().max()).tolist()[:6]
        for i in idx:
            print('    at', i, 'got', got[tuple(i)].item(), 'ref', ref[tuple(i)].item())


def main():
    what = sys.argv[1]
    B, N, C, H, W = (int(v) for v in sys.argv[2:7]) if len(sys.argv) >= 7 else (1, 100, 256, 200, 88)
    dev = torch.device('cuda:0')
    cfg, sd, h = build(C, 256, dev)
    x, pf, mask = ko.dummy_inputs(B, N, C, H, W, seed=9)
    xb, mb = x.to(dev).bfloat16(), mask.to(dev).bfloat16()
    xd = xb.double()
    ftw = sd['feat_transform.conv.weight'].to(dev).double().reshape(C, C)
    ftb = sd['feat_transform.conv.bias'].to(dev).double()
    xt = torch.einsum('oc,bchw->bohw', ftw, xd) + ftb[None, :, None, None]
    print('shape B%d N%d C%d %dx%d  tc_supported=%s' % (B, N, C, H, W, H * W % 8 == 0), flush=True)
    if what in ('pool', 'stage'):
        ref = torch.einsum('bnhw,bchw->bnc', (mb.double() > 0).double(), xt)
        for eng, nm in ((_lib.ENGINE_SIMT, 'pool simt'), (_lib.ENGINE_TC, 'pool tcgen05')):
            h.engine = eng
            got = ops.mask_pool(h, xb, mb)
            torch.cuda.synchronize()
            report(nm, got, ref)
    if what in ('mask', 'stage'):
        g = torch.Generator().manual_seed(3)
        mk = torch.randn(B, N, C, generator=g).to(dev)
        ref = torch.einsum('bnc,bchw->bnhw', mk.double(), xt)
        for eng, nm in ((_lib.ENGINE_SIMT, 'mask conv simt'), (_lib.ENGINE_TC, 'mask conv tcgen05')):
            h.engine = eng
            got = ops.mask_gemm(h, xb, mk)
            torch.cuda.synchronize()
            report(nm + ' (bf16 out)', got, ref.bfloat16())
            a = got.float().argmax(1)
            b = ref.bfloat16().float().argmax(1)
            print('    argmax mismatches vs fp64->bf16 reference: %d of %d' % (int((a != b).sum()), a.numel()))
    if what == 'stage':
        want = ko.kernel_update_head_forward(sd, cfg, ko.round_bf16(x), pf, ko.round_bf16(mask))
        for eng, nm in ((_lib.ENGINE_SIMT, 'simt'), (_lib.ENGINE_TC, 'tcgen05')):
            h.engine = eng
            cls, nm_, obj = h(xb, pf.to(dev), mb)
            torch.cuda.synchronize()
            report('stage cls ' + nm, cls, want[0])
            report('stage obj ' + nm, obj, want[2])
            report('stage mask ' + nm, nm_, ko.round_bf16(want[1]))
    print('done', what, flush=True)


if __name__ == '__main__':
    main()
