"""Plugin boundary: the reference resolves its heads by NAME through mmcv/mmdet registries
(`@HEADS.register_module()` knet/det/kernel_update_head.py:16, `@TRANSFORMER_LAYER.register_module()`
knet/kernel_updator.py:7) from config dicts.  When mmcv/mmdet are importable we register into the
real registries (force=True: knet/ and knet_vis/ both claim the same keys); otherwise a minimal
local registry with the same `register_module` / `build` surface is used, so the config blocks of
configs/det/_base_/models/knet_kitti_step_s3_r50_fpn.py:88-136 still build unchanged.
"""


class _LocalRegistry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError('%s is already registered in %s' % (key, self.name))
            self.module_dict[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        cfg = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        typ = cfg.pop('type')
        cls = self.module_dict[typ] if isinstance(typ, str) else typ
        return cls(**cfg)


try:  # real registries when the reference's dependencies are installed
    from mmcv.cnn.bricks.transformer import TRANSFORMER_LAYER  # type: ignore
    from mmdet.models.builder import HEADS  # type: ignore
    HAVE_MM = True
except Exception:  # noqa: BLE001 - any import problem means "not available"
    TRANSFORMER_LAYER = _LocalRegistry('transformer_layer')
    HEADS = _LocalRegistry('head')
    HAVE_MM = False


def build_transformer_layer(cfg, default_args=None):
    if HAVE_MM:
        from mmcv.cnn.bricks.transformer import build_transformer_layer as _b
        return _b(cfg, default_args)
    return TRANSFORMER_LAYER.build(cfg, default_args)


def build_head(cfg):
    if HAVE_MM:
        from mmdet.models.builder import build_head as _b
        return _b(cfg)
    return HEADS.build(cfg)


class _LossStub:
    """Only `use_sigmoid` of loss_cls is read on the inference path
    (knet/det/kernel_update_head.py:136, knet/det/kernel_iter_head.py:255-258)."""

    def __init__(self, use_sigmoid=False, **kwargs):
        self.use_sigmoid = use_sigmoid
        self.cfg = dict(kwargs, use_sigmoid=use_sigmoid)

    def __call__(self, *a, **k):
        raise NotImplementedError('training losses are outside this package (inference hot path only); '
                                  'install mmdet to build real losses')


def build_loss(cfg):
    if cfg is None:
        return None
    if HAVE_MM:
        try:
            from mmdet.models.builder import build_loss as _b
            return _b(cfg)
        except Exception:  # custom losses of the reference tree may not be registered
            pass
    cfg = dict(cfg)
    cfg.pop('type', None)
    return _LossStub(**cfg)
