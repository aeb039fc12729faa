"""CPU: the drop-in boundary -- C-ABI exports, registry names, state_dict contract, error behaviour.
No compute call is made (there is no GPU here and no CPU path in the product)."""
import ctypes
import os
import re

import pytest
import torch

import knet_oracle as ko

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

APPENDIX_C_KEYS = [
    'attention.attn.in_proj_weight', 'attention.attn.in_proj_bias', 'attention.attn.out_proj.weight',
    'attention.attn.out_proj.bias', 'attention_norm.weight', 'attention_norm.bias',
    'kernel_update_conv.dynamic_layer.weight', 'kernel_update_conv.dynamic_layer.bias',
    'kernel_update_conv.input_layer.weight', 'kernel_update_conv.input_layer.bias',
    'kernel_update_conv.input_gate.weight', 'kernel_update_conv.input_gate.bias',
    'kernel_update_conv.update_gate.weight', 'kernel_update_conv.update_gate.bias',
    'kernel_update_conv.norm_in.weight', 'kernel_update_conv.norm_in.bias',
    'kernel_update_conv.norm_out.weight', 'kernel_update_conv.norm_out.bias',
    'kernel_update_conv.input_norm_in.weight', 'kernel_update_conv.input_norm_in.bias',
    'kernel_update_conv.input_norm_out.weight', 'kernel_update_conv.input_norm_out.bias',
    'kernel_update_conv.fc_layer.weight', 'kernel_update_conv.fc_layer.bias',
    'kernel_update_conv.fc_norm.weight', 'kernel_update_conv.fc_norm.bias',
    'feat_transform.conv.weight', 'feat_transform.conv.bias',
    'ffn.layers.0.0.weight', 'ffn.layers.0.0.bias', 'ffn.layers.1.weight', 'ffn.layers.1.bias',
    'ffn_norm.weight', 'ffn_norm.bias', 'cls_fcs.0.weight', 'cls_fcs.1.weight', 'cls_fcs.1.bias',
    'fc_cls.weight', 'fc_cls.bias', 'mask_fcs.0.weight', 'mask_fcs.1.weight', 'mask_fcs.1.bias',
    'fc_mask.weight', 'fc_mask.bias']


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, 'include', 'vknet.h')).read()
    declared = sorted(set(re.findall(r'\b(vkn_[a-z_]+)\s*\(', hdr)))
    assert len(declared) >= 13
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), 'libvknet.so does not export %s' % name
    import vknet
    assert sorted(vknet._lib.SYMBOLS) == declared
    assert lib.vkn_version() == 100


def test_workspace_and_shape_validation(built_lib):
    import vknet
    L = vknet._lib
    s = L.make_shape(1, 100, 256, 200, 88, 2048, 19, 8, L.VKN_BF16, L.VKN_BF16)
    n = L.workspace_bytes(s)
    assert 1 << 20 < n < 1 << 30
    for bad in (dict(Cc=100), dict(Cc=512), dict(num_heads=3), dict(ffn_dim=100), dict(N=0), dict(x_dtype=7)):
        kw = dict(B=1, N=100, Cc=256, H=8, W=8, ffn_dim=2048, num_classes=19, num_heads=8, x_dtype=0, w_dtype=0)
        kw.update(bad)
        with pytest.raises(L.VknError):
            L.workspace_bytes(L.make_shape(**kw))
    assert b'C = 100' in L.lib().vkn_last_error() or True


def test_struct_sizes_match_header_layout(built_lib):
    import vknet
    L = vknet._lib
    p = ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(L.VknShape) == 14 * 4
    assert ctypes.sizeof(L.VknUpdatorW) == 20 * p
    assert ctypes.sizeof(L.VknAttnW) == 6 * p and ctypes.sizeof(L.VknFfnW) == 6 * p
    assert ctypes.sizeof(L.VknHeadW) == (3 + 20 + 6 + 6 + 1 + 3 * 4 + 2 + 3 * 4 + 2) * p
    assert ctypes.sizeof(L.VknLinkW) == (1 + 20 + 6 + 6) * p


def test_registry_builds_reference_config_blocks_unchanged(built_lib):
    import vknet
    cfg = ko.default_cfg()          # the mask_head dict of knet_kitti_step_s3_r50_fpn.py:88-136
    head = vknet.build_head(dict(type='KernelUpdateHead', **cfg))
    assert type(head).__name__ == 'KernelUpdateHead'
    assert type(head.kernel_update_conv).__name__ == 'KernelUpdator'
    assert list(head.state_dict().keys()) == APPENDIX_C_KEYS
    assert sum(p.numel() for p in head.parameters()) == 2046739
    assert head.mask_upsample_stride == 2 and head.num_classes == 19 and head.loss_cls.use_sigmoid
    head.load_state_dict(ko.random_state_dict(cfg), strict=True)
    upd = vknet.build_transformer_layer(dict(cfg['kernel_updator_cfg']))
    assert upd.dynamic_layer.weight.shape == (512, 256)


def test_video_head_state_dict_contract(built_lib):
    import vknet
    for pt, pl, total in (('ffn', None, None), ('update', 'update_dynamic_cov', None), ('ffn', 'update_dynamic_cov', None)):
        cfg = ko.default_cfg(previous='placeholder', previous_type=pt, previous_link=pl)
        head = vknet.build_head(dict(type='VideoKernelUpdateHead', **cfg))
        sd = ko.random_state_dict(cfg)
        assert sorted(head.state_dict().keys()) == sorted(sd.keys())
        head.load_state_dict(sd, strict=True)


def test_clip_head_state_dict_contract(built_lib):
    """KernelUpdateHeadVideo (knet_vis/tracker/kernel_update_head.py): with_cls=False drops the cls branch."""
    import vknet
    cfg = ko.default_cfg()
    sd = ko.random_state_dict(cfg)
    full = vknet.build_head(dict(type='KernelUpdateHeadVideo', with_cls=True, num_proposals=100, **cfg))
    assert list(full.state_dict().keys()) == APPENDIX_C_KEYS
    full.load_state_dict(sd, strict=True)
    nocls = vknet.build_head(dict(type='KernelUpdateHeadVideo', with_cls=False, num_proposals=100, **cfg))
    want = [k for k in APPENDIX_C_KEYS if not (k.startswith('cls_fcs') or k.startswith('fc_cls'))]
    assert list(nocls.state_dict().keys()) == want
    with pytest.raises(NotImplementedError):
        vknet.build_head(dict(type='KernelUpdateHeadVideo', query_merge_method='attention', **cfg)).check_supported()


def test_alias_packages_follow_the_reference_import_strings(built_lib):
    import importlib
    import vknet
    for mod, name in (('knet.kernel_updator', 'KernelUpdator'), ('knet.det.kernel_update_head', 'KernelUpdateHead'),
                      ('knet.video.kernel_update_head', 'VideoKernelUpdateHead'),
                      ('knet_vis.kernel_updator', 'KernelUpdator'), ('knet_vis.det.kernel_update_head', 'KernelUpdateHead'),
                      ('knet_vis.tracker.kernel_update_head', 'KernelUpdateHeadVideo')):
        m = importlib.import_module(mod)
        assert getattr(m, name) is getattr(vknet, name)
    assert vknet.HEADS.get('KernelUpdateHead') is vknet.KernelUpdateHead


def test_no_cpu_fallback_and_unsupported_configs_fail_loudly(built_lib):
    import vknet
    cfg = ko.default_cfg(in_channels=64, feedforward_channels=64, num_classes=3)
    head = vknet.build_head(dict(type='KernelUpdateHead', **cfg))
    x, pf, mask = ko.dummy_inputs(1, 5, 64, 4, 4)
    with pytest.raises(vknet.VknError, match='no CPU path'):
        head(x, pf, mask)
    with pytest.raises(NotImplementedError):
        vknet.build_head(dict(type='KernelUpdateHead', **ko.default_cfg(dropout=0.1)))
    with pytest.raises(NotImplementedError):
        head.loss()
    k3 = vknet.build_head(dict(type='KernelUpdateHead', **ko.default_cfg(in_channels=64, conv_kernel_size=3)))
    with pytest.raises(NotImplementedError):
        k3(x, torch.zeros(1, 5, 64, 3, 3), mask)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'video-k-net_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'knet_oracle' not in src and 'ref_shim' not in src, f


def test_alias_packages_overlay_the_rest_of_the_reference_package(tmp_path, built_lib):
    """`knet.det.kernel_update_head` is ours, any other `knet.*` module still resolves to a `knet/`
    directory later on sys.path (the reference tree): configs stay unchanged."""
    import subprocess
    import sys
    fake = tmp_path / 'reftree' / 'knet' / 'det'
    fake.mkdir(parents=True)
    (tmp_path / 'reftree' / 'knet' / '__init__.py').write_text('')
    (fake / '__init__.py').write_text('')
    (fake / 'kernel_iter_head.py').write_text('MARK = "reference file"\n')
    (fake / 'kernel_update_head.py').write_text('raise RuntimeError("the reference copy must be shadowed")\n')
    code = ('import sys; sys.path[:0] = [%r, %r]; import knet.det.kernel_iter_head as a, knet.det.kernel_update_head as b; '
            'print(a.MARK, b.KernelUpdateHead.__module__)' % (os.path.join(ROOT, 'video-k-net_b200'), str(tmp_path / 'reftree')))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == 'reference file vknet.kernel_update_head'
