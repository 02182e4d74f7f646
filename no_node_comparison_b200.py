"""Import shim: the package directory is named `no-node-comparison_b200/` (not a valid Python
identifier), so this module exposes it as the package `no_node_comparison_b200`."""
import os as _os

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "no-node-comparison_b200")
__path__ = [_dir]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(_dir, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
