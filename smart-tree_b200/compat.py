"""Register the B200 stand-ins under the third-party names the reference imports, so its own
modules (smart_tree.model.model, smart_tree.skeleton.graph, ...) can run unchanged:
    import smart_tree_b200.compat as compat; compat.install()
Nothing is registered if the real package is importable, unless force=True."""
import importlib
import sys


def install(force: bool = False):
    from . import frnn as _frnn
    from . import spconv as _spconv
    from .spconv import pytorch as _pt
    from .spconv import utils as _utils

    def have(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    done = []
    if force or not have("spconv"):
        _pt.utils = _utils
        _spconv.pytorch = _pt
        sys.modules["spconv"] = _spconv
        sys.modules["spconv.pytorch"] = _pt
        sys.modules["spconv.pytorch.utils"] = _utils
        done.append("spconv")
    if force or not have("frnn"):
        sys.modules["frnn"] = _frnn
        done.append("frnn")
    return done
