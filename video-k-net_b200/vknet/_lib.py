"""ctypes binding of libvknet.so (the C ABI declared in include/vknet.h).

There is deliberately no fallback: if the shared library is missing or a call fails, this
module raises.  PyTorch is used only as the owner of device memory and streams; every pointer
that crosses the boundary is a raw `data_ptr()`.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvknet.so')

VKN_F32, VKN_BF16 = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2
VKN_MAX_FCS = 4

_vp = C.c_void_p


class VknShape(C.Structure):
    _fields_ = [('B', C.c_int32), ('N', C.c_int32), ('C', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('ffn_dim', C.c_int32), ('num_classes', C.c_int32), ('num_heads', C.c_int32),
                ('x_dtype', C.c_int32), ('w_dtype', C.c_int32), ('with_ffn', C.c_int32),
                ('engine', C.c_int32), ('frames_per_set', C.c_int32), ('mask_thr_logit', C.c_float)]


class VknUpdatorW(C.Structure):
    _fields_ = [(n, _vp) for n in (
        'dyn_w', 'dyn_b', 'inp_w', 'inp_b', 'ig_w', 'ig_b', 'ug_w', 'ug_b',
        'norm_in_g', 'norm_in_b', 'norm_out_g', 'norm_out_b', 'inorm_in_g', 'inorm_in_b',
        'inorm_out_g', 'inorm_out_b', 'fc_w', 'fc_b', 'fc_norm_g', 'fc_norm_b')]


class VknAttnW(C.Structure):
    _fields_ = [(n, _vp) for n in ('in_w', 'in_b', 'out_w', 'out_b', 'norm_g', 'norm_b')]


class VknFfnW(C.Structure):
    _fields_ = [(n, _vp) for n in ('w1', 'b1', 'w2', 'b2', 'norm_g', 'norm_b')]


class VknHeadW(C.Structure):
    _fields_ = [('ft_w', _vp), ('ft_b', _vp), ('ft_wt_ext', _vp),
                ('upd', VknUpdatorW), ('attn', VknAttnW), ('ffn', VknFfnW),
                ('num_cls_fcs', C.c_int32), ('num_mask_fcs', C.c_int32),
                ('cls_fc_w', _vp * VKN_MAX_FCS), ('cls_ln_g', _vp * VKN_MAX_FCS), ('cls_ln_b', _vp * VKN_MAX_FCS),
                ('fc_cls_w', _vp), ('fc_cls_b', _vp),
                ('mask_fc_w', _vp * VKN_MAX_FCS), ('mask_ln_g', _vp * VKN_MAX_FCS), ('mask_ln_b', _vp * VKN_MAX_FCS),
                ('fc_mask_w', _vp), ('fc_mask_b', _vp), ('fc_pack', _vp)]


class VknMlpLayer(C.Structure):
    _fields_ = [('w', _vp), ('b', _vp), ('ln_g', _vp), ('ln_b', _vp), ('in_dim', C.c_int32), ('out_dim', C.c_int32),
                ('relu', C.c_int32)]


class VknLinkW(C.Structure):
    _fields_ = [('has_updator', C.c_int32), ('upd', VknUpdatorW), ('attn', VknAttnW), ('ffn', VknFfnW)]


class VknError(RuntimeError):
    pass


_lib = None

# every symbol include/vknet.h declares (tests check that the library exports all of them)
SYMBOLS = ('vkn_version', 'vkn_last_error', 'vkn_kernel_names', 'vkn_launch_count', 'vkn_profile_begin',
           'vkn_profile_end', 'vkn_debug_timestamps', 'vkn_workspace_bytes', 'vkn_mask_pool',
           'vkn_kernel_update', 'vkn_mhsa_ln', 'vkn_ffn_ln', 'vkn_heads', 'vkn_mask_gemm',
           'vkn_stage_forward', 'vkn_iter_forward', 'vkn_init_proposals', 'vkn_link_attend', 'vkn_rescale_masks',
           'vkn_panoptic_merge', 'vkn_mask_boxes', 'vkn_mlp', 'vkn_track_match', 'vkn_frame_chain_pack_bytes',
           'vkn_frame_chain_pack', 'vkn_match_cost_workspace_bytes', 'vkn_match_cost')


def lib():
    """Load libvknet.so once.  Raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VknError('libvknet.so not found at %s -- build it with `python -c "import __graft_entry__ as g; '
                       'g.build()"`; there is no CPU or PyTorch fallback for this path' % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.vkn_version.restype = C.c_int
    L.vkn_last_error.restype = C.c_char_p
    L.vkn_kernel_names.restype = C.c_char_p
    L.vkn_launch_count.restype = C.c_ulonglong
    L.vkn_profile_begin.restype = C.c_int
    L.vkn_profile_end.restype = C.c_int
    L.vkn_profile_end.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
    S, sz = C.POINTER(VknShape), C.c_size_t
    L.vkn_workspace_bytes.argtypes = [S, C.POINTER(sz)]
    L.vkn_mask_pool.argtypes = [S, C.POINTER(VknHeadW), _vp, _vp, _vp, _vp, sz, _vp]
    L.vkn_kernel_update.argtypes = [S, C.POINTER(VknUpdatorW), _vp, _vp, _vp, _vp, sz, _vp]
    L.vkn_mhsa_ln.argtypes = [S, C.POINTER(VknAttnW), _vp, _vp, _vp, _vp, sz, _vp]
    L.vkn_ffn_ln.argtypes = [S, C.POINTER(VknFfnW), _vp, _vp, _vp, sz, _vp]
    L.vkn_heads.argtypes = [S, C.POINTER(VknHeadW), _vp, _vp, _vp, _vp, sz, _vp]
    L.vkn_mask_gemm.argtypes = [S, C.POINTER(VknHeadW), _vp, _vp, _vp, _vp, sz, _vp]
    L.vkn_stage_forward.argtypes = [S, C.POINTER(VknHeadW), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, sz, _vp]
    L.vkn_iter_forward.argtypes = [S, C.POINTER(VknHeadW), C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, sz, _vp]
    L.vkn_init_proposals.argtypes = [S, _vp, _vp, _vp, _vp, _vp, _vp, _vp, sz, _vp]
    L.vkn_link_attend.argtypes = [S, C.POINTER(VknLinkW), _vp, _vp, _vp, _vp, _vp, sz, _vp]
    L.vkn_rescale_masks.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_float, _vp, _vp, _vp]
    L.vkn_panoptic_merge.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _vp, _vp, _vp, _vp,
                                     _vp, _vp, sz, _vp]
    L.vkn_mask_boxes.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]
    L.vkn_mlp.argtypes = [C.POINTER(VknMlpLayer), C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, sz, _vp]
    L.vkn_track_match.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int64, _vp, _vp, _vp,
                                  _vp, sz, _vp]
    L.vkn_frame_chain_pack_bytes.argtypes = [S, C.POINTER(VknHeadW), C.POINTER(sz)]
    L.vkn_frame_chain_pack.argtypes = [S, C.POINTER(VknHeadW), _vp, sz, _vp]
    L.vkn_match_cost_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(sz)]
    L.vkn_match_cost.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), _vp, _vp, sz, _vp]
    L.vkn_debug_timestamps.restype = C.c_int
    L.vkn_debug_timestamps.argtypes = [_vp, C.c_size_t]
    for name in SYMBOLS[7:]:
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise VknError('libvknet call failed (%d): %s' % (rc, lib().vkn_last_error().decode()))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """the current CUDA stream of `device` (default: the current device) as the void* the C ABI takes"""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dtype_code(dt):
    if dt == torch.float32:
        return VKN_F32
    if dt == torch.bfloat16:
        return VKN_BF16
    raise VknError('unsupported dtype %s (float32 and bfloat16 only)' % dt)


def make_shape(B, N, Cc, H, W, ffn_dim, num_classes, num_heads, x_dtype, w_dtype, with_ffn=True,
               engine=ENGINE_AUTO, mask_thr_logit=0.0, frames_per_set=1):
    return VknShape(B, N, Cc, H, W, ffn_dim, num_classes, num_heads, x_dtype, w_dtype, int(bool(with_ffn)),
                    engine, int(frames_per_set), float(mask_thr_logit))


def workspace_bytes(shape):
    n = C.c_size_t(0)
    check(lib().vkn_workspace_bytes(C.byref(shape), C.byref(n)))
    return n.value


class Workspace:
    """Caller-owned scratch, grown on demand and reused across calls on one device."""

    def __init__(self):
        self.buf = None

    def get(self, shape, device):
        need = workspace_bytes(shape)
        if self.buf is None or self.buf.numel() < need or self.buf.device != device:
            self.buf = torch.empty(need + 256, dtype=torch.uint8, device=device)
        off = (-self.buf.data_ptr()) % 256
        return C.c_void_p(self.buf.data_ptr() + off), self.buf.numel() - off


def kernel_names():
    return lib().vkn_kernel_names().decode().split('\n')


def launch_count():
    return int(lib().vkn_launch_count())


class profile:
    """with profile() as p: ...calls...  ->  p.records = [(kernel_name, ms), ...] in launch order."""
    MAX = 512

    def __enter__(self):
        check(lib().vkn_profile_begin())
        self.records = []
        return self

    def __exit__(self, *exc):
        names = (C.c_char_p * self.MAX)()
        ms = (C.c_float * self.MAX)()
        n = C.c_int(0)
        check(lib().vkn_profile_end(names, ms, self.MAX, C.byref(n)))
        self.records = [(names[i].decode(), float(ms[i])) for i in range(n.value)]
        return False
