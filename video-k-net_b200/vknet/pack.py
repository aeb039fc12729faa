"""Turns module parameters into the POD weight structs of include/vknet.h.

Matrices go to the weight storage dtype (fp32 or bf16), vectors (bias, LayerNorm affine) are
always fp32.  bf16 storage is only ever chosen when the parameters already ARE bf16 (see
`weight_dtype_of`), so packing never changes a value.
"""
import ctypes as C

import torch

from . import _lib


class Packer:
    def __init__(self, device, w_dtype_code):
        self.device = device
        self.wt = torch.bfloat16 if w_dtype_code == _lib.VKN_BF16 else torch.float32
        self.keep = []      # tensors the structs point into

    def mat(self, t):
        t = t.detach().to(device=self.device, dtype=self.wt).contiguous()
        self.keep.append(t)
        return C.c_void_p(t.data_ptr())

    def vec(self, t):
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.keep.append(t)
        return C.c_void_p(t.data_ptr())


def weight_dtype_of(params):
    """bf16 storage iff every floating parameter is stored in bf16 (module.bfloat16())."""
    dts = {p.dtype for p in params}
    if dts == {torch.bfloat16}:
        return _lib.VKN_BF16
    if dts == {torch.float32}:
        return _lib.VKN_F32
    raise _lib.VknError('parameters must be uniformly float32 or bfloat16, got %s' % sorted(map(str, dts)))


def pack_updator(pk, m):
    """m: a KernelUpdator module (knet/kernel_updator.py:36-54 parameter names)."""
    w = _lib.VknUpdatorW()
    w.dyn_w, w.dyn_b = pk.mat(m.dynamic_layer.weight), pk.vec(m.dynamic_layer.bias)
    w.inp_w, w.inp_b = pk.mat(m.input_layer.weight), pk.vec(m.input_layer.bias)
    w.ig_w, w.ig_b = pk.mat(m.input_gate.weight), pk.vec(m.input_gate.bias)
    w.ug_w, w.ug_b = pk.mat(m.update_gate.weight), pk.vec(m.update_gate.bias)
    w.norm_in_g, w.norm_in_b = pk.vec(m.norm_in.weight), pk.vec(m.norm_in.bias)
    w.norm_out_g, w.norm_out_b = pk.vec(m.norm_out.weight), pk.vec(m.norm_out.bias)
    w.inorm_in_g, w.inorm_in_b = pk.vec(m.input_norm_in.weight), pk.vec(m.input_norm_in.bias)
    w.inorm_out_g, w.inorm_out_b = pk.vec(m.input_norm_out.weight), pk.vec(m.input_norm_out.bias)
    w.fc_w, w.fc_b = pk.mat(m.fc_layer.weight), pk.vec(m.fc_layer.bias)
    w.fc_norm_g, w.fc_norm_b = pk.vec(m.fc_norm.weight), pk.vec(m.fc_norm.bias)
    return w


def pack_attn(pk, att, norm):
    w = _lib.VknAttnW()
    w.in_w, w.in_b = pk.mat(att.attn.in_proj_weight), pk.vec(att.attn.in_proj_bias)
    w.out_w, w.out_b = pk.mat(att.attn.out_proj.weight), pk.vec(att.attn.out_proj.bias)
    w.norm_g, w.norm_b = pk.vec(norm.weight), pk.vec(norm.bias)
    return w


def pack_ffn(pk, ffn, norm):
    w = _lib.VknFfnW()
    w.w1, w.b1 = pk.mat(ffn.layers[0][0].weight), pk.vec(ffn.layers[0][0].bias)
    w.w2, w.b2 = pk.mat(ffn.layers[1].weight), pk.vec(ffn.layers[1].bias)
    w.norm_g, w.norm_b = pk.vec(norm.weight), pk.vec(norm.bias)
    return w


def pack_head(pk, head):
    """head: KernelUpdateHead-like module (state_dict contract SURVEY.md Appendix C)."""
    Cc = head.in_channels
    w = _lib.VknHeadW()
    if head.feat_transform is not None:
        ftw = head.feat_transform.conv.weight.detach().reshape(Cc, Cc)
        ftb = head.feat_transform.conv.bias.detach()
    else:
        p0 = head.fc_mask.weight
        ftw = torch.eye(Cc, device=p0.device, dtype=p0.dtype)
        ftb = torch.zeros(Cc, device=p0.device, dtype=p0.dtype)
    w.ft_w, w.ft_b = pk.mat(ftw), pk.vec(ftb)
    w.ft_wt_ext = pk.mat(torch.cat([ftw.t(), ftb[None, :]], dim=0))
    w.upd = pack_updator(pk, head.kernel_update_conv)
    w.attn = pack_attn(pk, head.attention, head.attention_norm)
    if head.with_ffn:
        w.ffn = pack_ffn(pk, head.ffn, head.ffn_norm)
    ncls, nmask = len(head.cls_fcs) // 3, len(head.mask_fcs) // 3
    if ncls > _lib.VKN_MAX_FCS or nmask > _lib.VKN_MAX_FCS:
        raise _lib.VknError('at most %d cls/mask FC layers are supported' % _lib.VKN_MAX_FCS)
    w.num_cls_fcs, w.num_mask_fcs = ncls, nmask
    for i in range(ncls):
        w.cls_fc_w[i] = pk.mat(head.cls_fcs[3 * i].weight)
        w.cls_ln_g[i], w.cls_ln_b[i] = pk.vec(head.cls_fcs[3 * i + 1].weight), pk.vec(head.cls_fcs[3 * i + 1].bias)
    for i in range(nmask):
        w.mask_fc_w[i] = pk.mat(head.mask_fcs[3 * i].weight)
        w.mask_ln_g[i], w.mask_ln_b[i] = pk.vec(head.mask_fcs[3 * i + 1].weight), pk.vec(head.mask_fcs[3 * i + 1].bias)
    if head.fc_cls is not None:          # KernelUpdateHeadVideo(with_cls=False) has no cls branch
        w.fc_cls_w, w.fc_cls_b = pk.mat(head.fc_cls.weight), pk.vec(head.fc_cls.bias)
    w.fc_mask_w, w.fc_mask_b = pk.mat(head.fc_mask.weight), pk.vec(head.fc_mask.bias)
    return w


def attach_frame_chain_pack(pk, w, shape):
    """VknHeadW.fc_pack: the weights re-laid for the single-frame row engine (csrc/framechain.cu), built once per weight
    version by the library itself (vkn_frame_chain_pack); heads the engine does not apply to keep fc_pack = NULL."""
    if torch.device(pk.device).type != 'cuda':
        return
    L = _lib.lib()
    n = C.c_size_t(0)
    _lib.check(L.vkn_frame_chain_pack_bytes(C.byref(shape), C.byref(w), C.byref(n)))
    if n.value == 0:
        return
    buf = torch.empty(n.value, dtype=torch.uint8, device=pk.device)
    with torch.cuda.device(pk.device):
        _lib.check(L.vkn_frame_chain_pack(C.byref(shape), C.byref(w), _lib.ptr(buf), n.value, _lib.stream_ptr(pk.device)))
    pk.keep.append(buf)
    w.fc_pack = buf.data_ptr()


def pack_link(pk, updator, att, att_norm, ffn, ffn_norm):
    w = _lib.VknLinkW()
    w.has_updator = int(updator is not None)
    if updator is not None:
        w.upd = pack_updator(pk, updator)
    w.attn = pack_attn(pk, att, att_norm)
    w.ffn = pack_ffn(pk, ffn, ffn_norm)
    return w
