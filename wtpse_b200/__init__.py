"""Importable alias of the package directory ``wt-pse-code_b200/`` (a hyphen cannot be imported).

``import wtpse_b200`` executes ``wt-pse-code_b200/__init__.py`` with this module's ``__path__``
pointing at that directory, so ``wtpse_b200.functional`` etc. resolve to the files there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "wt-pse-code_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
