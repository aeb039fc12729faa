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
        # vkn_iter_forward runs every stage with ONE VknShape (built from stage 0): stages with different settings would
        # silently run with stage 0's values, where stage-by-stage module calls honour each head's own config
        h0 = self.heads[0]
        for i, h in enumerate(self.heads[1:], 1):
            for attr in ('in_channels', 'feedforward_channels', 'with_ffn', 'num_heads', 'hard_mask_thr', 'engine',
                         'conv_kernel_size'):
                if getattr(h, attr) != getattr(h0, attr):
                    raise NotImplementedError('KernelIterLoop: stage %d differs from stage 0 in %s (%r vs %r); run such heads '
                                              'stage by stage' % (i, attr, getattr(h, attr), getattr(h0, attr)))
            if h.fc_cls.out_features != h0.fc_cls.out_features:
                raise NotImplementedError('KernelIterLoop: stage %d has a different number of classes' % i)
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


class FramesInFlight:
    """Throughput form of the loop: ONE CUDA graph that runs `branches` independent frame batches
    concurrently (fork/join inside the capture), each on its own stream with its own workspace.

    The loop of one frame is a chain of small latency-bound kernels that occupies a fraction of the
    148 SMs; frames are independent, so several are kept in flight.  With `host_io=True`
    the graph also contains the host<->device copies: pinned host inputs -> static device buffers before
    the loop and the result tuple -> pinned host buffers after it (one launch = the whole e2e step).
    """

    def __init__(self, heads, branches=4, batch=1):
        self.heads = list(heads)
        self.branches = branches
        self.batch = batch
        self.loops = [KernelIterLoop(self.heads) for _ in range(branches)]
        self.static = []
        self.host_in = []
        self.host_out = []
        self.graph = None

    @torch.no_grad()
    def capture(self, inputs, host_io=False):
        """inputs: list (len == branches) of (x [b,C,H,W], proposal_feat [b,N,C], mask_preds [b,N,H,W])
        device tensors (host_io=False) or pinned host tensors (host_io=True)."""
        assert len(inputs) == self.branches
        h0 = self.heads[0]
        dev = next(h0.parameters()).device
        streams = [torch.cuda.Stream(device=dev) for _ in range(self.branches)]
        for (x, pf, mask), lp in zip(inputs, self.loops):
            xd, pfd, md = x.to(dev, non_blocking=False), pf.to(dev), mask.to(dev)
            xd, pfd, md, B, N, H, W, _ = h0._prepare(xd, pfd, md)
            st = dict(x=xd.clone(), pf=pfd.clone(), mask=md.clone(),
                      cls=torch.empty(B, N, h0.fc_cls.out_features, dtype=torch.float32, device=dev),
                      out_mask=torch.empty(B, N, H, W, dtype=xd.dtype, device=dev),
                      obj=torch.empty(B, N, h0.in_channels, dtype=torch.float32, device=dev))
            self.static.append(st)
            if host_io:
                self.host_in.append((x, pf.reshape(st['pf'].shape), mask))
                self.host_out.append(tuple(torch.empty(t.shape, dtype=t.dtype).pin_memory()
                                           for t in (st['cls'], st['out_mask'], st['obj'])))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up outside capture
            for st, lp in zip(self.static, self.loops):
                lp.forward(st['x'], st['pf'], st['mask'], out=(st['cls'], st['out_mask'], st['obj']))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream(dev)
            fork = torch.cuda.Event()
            fork.record(cur)
            joins = []
            for i, (st, lp, s) in enumerate(zip(self.static, self.loops, streams)):
                s.wait_event(fork)
                with torch.cuda.stream(s):
                    if host_io:
                        hx, hpf, hm = self.host_in[i]
                        st['x'].copy_(hx, non_blocking=True)
                        st['pf'].copy_(hpf, non_blocking=True)
                        st['mask'].copy_(hm, non_blocking=True)
                    lp.forward(st['x'], st['pf'], st['mask'], out=(st['cls'], st['out_mask'], st['obj']))
                    if host_io:
                        for dst, src in zip(self.host_out[i], (st['cls'], st['out_mask'], st['obj'])):
                            dst.copy_(src, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s)
                    joins.append(ev)
            for ev in joins:
                cur.wait_event(ev)
        self.graph = g
        self._streams = streams
        return self

    def replay(self):
        self.graph.replay()
        return [(st['cls'], st['out_mask'], st['obj']) for st in self.static]

    @property
    def frames_per_replay(self):
        return self.branches * self.batch
