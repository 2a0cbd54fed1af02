"""Importable alias of the ``continual-skeletons_b200/`` package directory (a hyphen is not a
valid Python identifier).  ``import continual_skeletons_b200`` resolves every submodule from
there; nothing lives in this directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "continual-skeletons_b200"))

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
