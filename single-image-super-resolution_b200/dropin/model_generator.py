"""Drop-in for the reference's top-level ``model_generator`` module: put this directory first on sys.path and
``from model_generator import ...`` in config.py / visualisation.py resolves to the B200-native classes."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
import sisr_b200  # noqa: E402,F401
from sisr_b200.model_generator import *  # noqa: E402,F401,F403
from sisr_b200.model_generator import __dict__ as _d  # noqa: E402
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
