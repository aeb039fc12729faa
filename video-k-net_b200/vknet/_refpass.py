"""Pass-through to the reference's own training-side methods (loss / get_targets).

The drop-in modules replace `forward`; everything the reference computes in pure torch for TRAINING stays the reference's
code: the same-named source file is located in the reference tree (the first `<sys.path entry>/<relative file>` that is not
this package's alias), imported once under a private module name, and its functions are called with the drop-in module as
`self` (identical attribute names).  Nothing is copied.  Requires the reference checkout on sys.path and its dependencies
(mmcv / mmdet); otherwise NotImplementedError explains what is missing.
"""
import importlib.util
import os
import sys

_cache = {}
_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))     # .../video-k-net_b200


def _find(relpath):
    for base in sys.path:
        cand = os.path.abspath(os.path.join(base or '.', relpath))
        if os.path.isfile(cand) and not cand.startswith(_HERE + os.sep):
            return cand
    return None


def reference_function(relpath, cls_name, fn_name):
    key = (relpath, cls_name)
    if key not in _cache:
        path = _find(relpath)
        if path is None:
            raise NotImplementedError('%s.%s is the reference\'s own pure-torch code and is passed through to it: put the '
                                      'Video-K-Net checkout on sys.path (looked for %s)' % (cls_name, fn_name, relpath))
        from . import registry
        saved = {}
        for reg in (registry.HEADS, registry.TRANSFORMER_LAYER):       # the reference registers the same keys without force
            md = getattr(reg, 'module_dict', None) or getattr(reg, '_module_dict', None)
            if md is not None:
                saved[id(md)] = (md, dict(md))
        name = '_vknet_refpass_' + relpath.replace('/', '_').replace('.py', '')
        try:
            for md, _ in saved.values():
                md.clear()
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
        except ImportError as e:
            sys.modules.pop(name, None)
            raise NotImplementedError('%s.%s is passed through to the reference (%s), which needs its own dependencies: %s'
                                      % (cls_name, fn_name, path, e))
        finally:
            for md, snapshot in saved.values():
                md.clear()
                md.update(snapshot)
        _cache[key] = getattr(mod, cls_name)
    fn = getattr(_cache[key], fn_name, None)
    if fn is None:
        raise NotImplementedError('the reference class %s has no method %s' % (cls_name, fn_name))
    return fn
