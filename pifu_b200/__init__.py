"""Import shim: the product package lives in ``rgb-d-pifuhd_b200/`` (a directory name
that is not a valid Python identifier), so ``import pifu_b200`` resolves to it here.

Every submodule (``pifu_b200.mesh_util``, ``pifu_b200.PIFuMRNet`` ...) is a file of
``rgb-d-pifuhd_b200/``; this file only redirects the package search path.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "rgb-d-pifuhd_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
