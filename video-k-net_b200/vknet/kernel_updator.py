"""KernelUpdator -- drop-in for knet/kernel_updator.py:7-94 (registered under the same
TRANSFORMER_LAYER key, same constructor kwargs, same state_dict keys, same forward contract).
The math runs in libvknet.so (`vkn_kernel_update`); there is no PyTorch fallback.
"""
import torch
import torch.nn as nn

from . import _lib, pack
from .bricks import make_ln
from .registry import TRANSFORMER_LAYER


@TRANSFORMER_LAYER.register_module(force=True)
class KernelUpdator(nn.Module):

    def __init__(self, in_channels=256, feat_channels=64, out_channels=None, input_feat_shape=3,
                 gate_sigmoid=True, gate_norm_act=False, activate_out=False,
                 act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN')):
        super().__init__()
        self.in_channels = in_channels
        self.feat_channels = feat_channels
        self.out_channels_raw = out_channels
        self.gate_sigmoid = gate_sigmoid
        self.gate_norm_act = gate_norm_act
        self.activate_out = activate_out
        if isinstance(input_feat_shape, int):
            input_feat_shape = [input_feat_shape] * 2
        self.input_feat_shape = input_feat_shape
        self.act_cfg = act_cfg
        self.norm_cfg = norm_cfg
        self.out_channels = out_channels if out_channels else in_channels
        if act_cfg.get('type', 'ReLU') != 'ReLU':
            raise NotImplementedError('only ReLU is on the shipped path')
        self.num_params_in = self.feat_channels
        self.num_params_out = self.feat_channels
        self.dynamic_layer = nn.Linear(self.in_channels, self.num_params_in + self.num_params_out)
        self.input_layer = nn.Linear(self.in_channels, self.num_params_in + self.num_params_out, 1)
        self.input_gate = nn.Linear(self.in_channels, self.feat_channels, 1)
        self.update_gate = nn.Linear(self.in_channels, self.feat_channels, 1)
        if self.gate_norm_act:
            self.gate_norm = make_ln(norm_cfg, self.feat_channels)
        self.norm_in = make_ln(norm_cfg, self.feat_channels)
        self.norm_out = make_ln(norm_cfg, self.feat_channels)
        self.input_norm_in = make_ln(norm_cfg, self.feat_channels)
        self.input_norm_out = make_ln(norm_cfg, self.feat_channels)
        self.activation = nn.ReLU(inplace=True)
        self.fc_layer = nn.Linear(self.feat_channels, self.out_channels, 1)
        self.fc_norm = make_ln(norm_cfg, self.out_channels)
        self._ws = _lib.Workspace()

    def check_supported(self):
        """The CUDA path covers the configuration every shipped config uses
        (configs/det/_base_/models/knet_kitti_step_s3_r50_fpn.py:110-117)."""
        if not (self.in_channels == self.feat_channels == self.out_channels):
            raise NotImplementedError('KernelUpdator CUDA path needs in == feat == out channels')
        if not self.gate_sigmoid or self.gate_norm_act or self.activate_out:
            raise NotImplementedError('only gate_sigmoid=True, gate_norm_act=False, activate_out=False is shipped')

    @torch.no_grad()
    def forward(self, update_feature, input_feature):
        """update_feature [..., C] (pooled feature), input_feature [B, N, K*K, C] with K == 1.
        Returns [P, 1, C] like the reference (knet/kernel_updator.py:94)."""
        self.check_supported()
        Cc = self.in_channels
        uf = update_feature.reshape(-1, Cc)
        P = uf.shape[0]
        inp = input_feature.reshape(P, -1, Cc)
        if inp.shape[1] != 1:
            raise NotImplementedError('conv_kernel_size != 1 is not on the shipped path')
        if not uf.is_cuda:
            raise _lib.VknError('vknet has no CPU path: inputs must live on a CUDA device')
        uf = uf.to(torch.float32).contiguous()
        inp = inp.reshape(P, Cc).to(torch.float32).contiguous()
        wd = pack.weight_dtype_of(self.parameters())
        pk = pack.Packer(uf.device, wd)
        w = pack.pack_updator(pk, self)
        shape = _lib.make_shape(1, P, Cc, 1, 1, 32, 1, 8 if Cc % 8 == 0 and Cc // 8 <= 32 else Cc // 32,
                                _lib.VKN_F32, wd)
        out = torch.empty(P, Cc, dtype=torch.float32, device=uf.device)
        ws, wsb = self._ws.get(shape, uf.device)
        _lib.check(_lib.lib().vkn_kernel_update(shape, w, _lib.ptr(uf), _lib.ptr(inp), _lib.ptr(out), ws, wsb,
                                                _lib.stream_ptr()))
        return out.reshape(P, 1, Cc)
