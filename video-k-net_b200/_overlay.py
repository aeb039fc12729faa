"""Package overlay helper for the alias packages `knet` / `knet_vis`.

The alias packages shadow only the three hot-path modules.  Every other submodule of the reference's
package of the same name (e.g. `knet.det.kernel_iter_head`) must keep resolving to the reference tree,
so each alias package appends the same-named directories found later on sys.path to its `__path__`.
"""
import os
import sys


def extend_path(pkg_path, pkg_name):
    here = [os.path.abspath(p) for p in pkg_path]
    out = list(pkg_path)
    rel = pkg_name.replace('.', os.sep)
    for base in sys.path:
        cand = os.path.abspath(os.path.join(base or '.', rel))
        if os.path.isdir(cand) and cand not in here and cand not in out:
            out.append(cand)
    return out
