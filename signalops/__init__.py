"""Importable name for the host-side package.

The package body lives in ``signaloperators.jl_b200/host`` (the directory name the
build contract asks for is not a valid Python identifier), so this shim only
extends ``__path__`` and re-exports the public API.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__.append(_os.path.join(_os.path.dirname(_here), "signaloperators.jl_b200", "host"))

from .api import *  # noqa: F401,F403,E402
from .api import __all__  # noqa: F401,E402
