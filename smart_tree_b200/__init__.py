"""Import alias: the product package lives in the directory ``smart-tree_b200/`` (the
name the build contract fixes), which is not a legal Python identifier.  This stub makes
it importable as ``smart_tree_b200`` by pointing ``__path__`` at that directory and
executing its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "smart-tree_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
