"""Importable alias of the `fenicsx-fus_b200/` package directory (a hyphen cannot be imported).

`import fenicsx_fus_b200` executes fenicsx-fus_b200/__init__.py with this package's __path__
pointing there, so submodules (capi, build) resolve to the real files.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "fenicsx-fus_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
