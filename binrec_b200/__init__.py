"""Importable alias for the hyphenated package directory ``binary-recommendation_b200/``.

A directory name with a hyphen cannot be imported by name, so this module points its
``__path__`` at that directory and executes its ``__init__.py``.  Everything lives there;
``import binrec_b200.BPRModel`` resolves to ``binary-recommendation_b200/BPRModel.py``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "binary-recommendation_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _f
