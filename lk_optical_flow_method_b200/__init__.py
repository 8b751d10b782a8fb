"""Importable alias of the product package.

The product lives in ``lk-optical-flow-method_b200/`` (the directory name the build contract asks
for).  A hyphen cannot appear in a Python ``import`` statement, so this shim gives that directory the
importable name ``lk_optical_flow_method_b200``: it points ``__path__`` at it and executes its
``__init__.py`` in this namespace.  There is no code of its own here.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "lk-optical-flow-method_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
