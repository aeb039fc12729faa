"""Parameter containers with the state_dict key layout of the mmcv bricks the reference builds
(SURVEY.md Appendix B/C).  They hold weights only; the math runs in libvknet.so.

  MultiheadAttention -> attn.in_proj_weight / attn.in_proj_bias / attn.out_proj.{weight,bias}
  FFN                -> layers.0.0.{weight,bias} / layers.1.{weight,bias}
  ConvModule (1x1)   -> conv.{weight,bias}
"""
import math

import torch.nn as nn


class MultiheadAttentionParams(nn.Module):
    def __init__(self, embed_dims, num_heads, attn_drop=0.0, **kwargs):
        super().__init__()
        if attn_drop not in (0, 0.0):
            raise NotImplementedError('attention dropout != 0 is not on the shipped path (configs use 0.0)')
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, 0.0)


class FFNParams(nn.Module):
    def __init__(self, embed_dims, feedforward_channels, num_fcs=2, act_cfg=None, ffn_drop=0.0, **kwargs):
        super().__init__()
        ffn_drop = kwargs.pop('dropout', ffn_drop)
        if num_fcs != 2:
            raise NotImplementedError('num_ffn_fcs != 2 is not on the shipped path')
        if ffn_drop not in (0, 0.0):
            raise NotImplementedError('FFN dropout != 0 is not on the shipped path (configs use 0.0)')
        if act_cfg is not None and act_cfg.get('type', 'ReLU') != 'ReLU':
            raise NotImplementedError('only ReLU FFNs are on the shipped path')
        self.embed_dims = embed_dims
        self.feedforward_channels = feedforward_channels
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(0.0)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.0))


_UNSET = object()


class Conv1x1Params(nn.Module):
    """mmcv ConvModule for the only form the shipped configs build: 1x1, stride 1, conv_cfg Conv2d, bias, NO norm and NO
    activation.  ConvModule's own default is act_cfg=dict(type='ReLU'): a config that omits act_cfg gets a ReLU in the
    reference, which the algebraic fold of feat_transform cannot express -- so `act_cfg=None` must be
    given explicitly (all 29 shipped configs do); anything else raises."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, conv_cfg=None,
                 norm_cfg=None, act_cfg=_UNSET, bias='auto', **kwargs):
        super().__init__()
        if kernel_size != 1 or stride != 1 or padding != 0:
            raise NotImplementedError('feat_transform must be the 1x1 / stride-1 conv of the shipped configs')
        if act_cfg is _UNSET:
            raise NotImplementedError("feat_transform_cfg without act_cfg: mmcv's ConvModule would add its default ReLU, which "
                                      'is not on the shipped path -- pass act_cfg=None as the shipped configs do')
        if norm_cfg is not None or act_cfg is not None:
            raise NotImplementedError('feat_transform with norm/activation is not on the shipped path')
        if conv_cfg is not None and conv_cfg.get('type', 'Conv2d') not in ('Conv2d', 'Conv'):
            raise NotImplementedError('feat_transform conv_cfg type %r is not on the shipped path' % conv_cfg.get('type'))
        if bias not in ('auto', True):
            raise NotImplementedError('feat_transform without bias is not on the shipped path')
        if kwargs:
            raise NotImplementedError('unsupported feat_transform_cfg keys: %s' % sorted(kwargs))
        self.conv = nn.Conv2d(in_channels, out_channels, 1, bias=True)


def make_ln(cfg, n):
    if cfg is not None and cfg.get('type', 'LN') != 'LN':
        raise NotImplementedError('only LayerNorm (type="LN") is on the shipped path')
    return nn.LayerNorm(n)


def bias_init_with_prob(p):
    return float(-math.log((1 - p) / p))
