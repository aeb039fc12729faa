"""The S-stage loop of KernelIterHead.simple_test (knet/det/kernel_iter_head.py:246-253) as ONE
C-ABI call (`vkn_iter_forward`), optionally captured in a CUDA graph.

The callers (the five *IterHead classes) are the boundary and are not re-implemented; this class is
the opt-in fast path an iter head uses instead of calling `mask_head[stage](...)` S times: it
returns what simple_test reads -- the LAST stage's (cls_score, mask_preds, object_feats).
"""
import ctypes as C

import torch

from . import _lib


class KernelIterLoop:
    def __init__(self, heads):
        self.heads = list(heads)
        if not self.heads:
            raise ValueError('need at least one stage')
        self._ws = _lib.Workspace()
        self._graph = None
        self._static = None

    def _pack(self, device):
        packed = [h.packed_weights(device) for h in self.heads]
        wds = {p[2] for p in packed}
        if len(wds) != 1:
            raise _lib.VknError('all stages must store weights in the same dtype')
        arr = (_lib.VknHeadW * len(packed))(*[p[0] for p in packed])
        return arr, wds.pop()

    @torch.no_grad()
    def forward(self, x, proposal_feat, mask_preds, out=None):
        """x [B,C,H,W], proposal_feat [B,N,C,1,1] | [B,N,C], mask_preds [B,N,H,W]
        -> (cls_score [B,N,ncls], mask_preds [B,N,H,W], object_feats [B,N,C,1,1]) of the last stage."""
        h0 = self.heads[0]
        x, pf, mask_preds, B, N, H, W, xd = h0._prepare(x, proposal_feat, mask_preds)
        arr, wd = self._pack(x.device)
        shape = h0._shape(B, N, H, W, xd, wd)
        dev, Cc = x.device, h0.in_channels
        if out is None:
            out = (torch.empty(B, N, h0.fc_cls.out_features, dtype=torch.float32, device=dev),
                   torch.empty(B, N, H, W, dtype=x.dtype, device=dev),
                   torch.empty(B, N, Cc, dtype=torch.float32, device=dev))
        cls, new_mask, obj = out
        ws, wsb = self._ws.get(shape, dev)
        _lib.check(_lib.lib().vkn_iter_forward(shape, arr, len(self.heads), _lib.ptr(x), _lib.ptr(pf),
                                               _lib.ptr(mask_preds), _lib.ptr(cls), _lib.ptr(new_mask),
                                               _lib.ptr(obj), ws, wsb, _lib.stream_ptr()))
        return cls, new_mask, obj.reshape(B, N, Cc, 1, 1)

    __call__ = forward

    # ---- CUDA-graph mode: static buffers, one replay per frame batch ------------------------------
    @torch.no_grad()
    def capture(self, x, proposal_feat, mask_preds):
        """Capture the loop for these shapes/dtypes.  Afterwards `replay(x, pf, mask)` copies the new
        inputs into the static buffers (device-to-device or host-to-device) and replays the graph."""
        h0 = self.heads[0]
        x, pf, mask_preds, B, N, H, W, xd = h0._prepare(x, proposal_feat, mask_preds)
        dev, Cc = x.device, h0.in_channels
        st = dict(x=x.clone(), pf=pf.clone(), mask=mask_preds.clone(),
                  cls=torch.empty(B, N, h0.fc_cls.out_features, dtype=torch.float32, device=dev),
                  out_mask=torch.empty(B, N, H, W, dtype=x.dtype, device=dev),
                  obj=torch.empty(B, N, Cc, dtype=torch.float32, device=dev))
        self._static = st
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):       # warm-up outside capture (workspace growth, attribute sets)
            for _ in range(2):
                self.forward(st['x'], st['pf'], st['mask'], out=(st['cls'], st['out_mask'], st['obj']))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward(st['x'], st['pf'], st['mask'], out=(st['cls'], st['out_mask'], st['obj']))
        self._graph = g
        return self

    @torch.no_grad()
    def replay(self, x=None, proposal_feat=None, mask_preds=None):
        st = self._static
        if x is not None:
            st['x'].copy_(x, non_blocking=True)
        if proposal_feat is not None:
            st['pf'].copy_(proposal_feat.reshape(st['pf'].shape), non_blocking=True)
        if mask_preds is not None:
            st['mask'].copy_(mask_preds, non_blocking=True)
        self._graph.replay()
        B, N, Cc = st['obj'].shape
        return st['cls'], st['out_mask'], st['obj'].reshape(B, N, Cc, 1, 1)
