"""Importable alias of the product package.

The package directory is named after the reference repository
(`dealii-galerkin-difference-methods_b200/`), which is not a valid Python
identifier; this shim makes it importable as `gdm_b200`.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "dealii-galerkin-difference-methods_b200"))

from .api import *  # noqa: E402,F401,F403
from . import api as _api  # noqa: E402

__all__ = _api.__all__
