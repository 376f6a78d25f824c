"""Importable alias of the product package.

The product directory is named `smart-nar_fast_tts_b200/` (the reference repo's name + _b200); a hyphen
cannot appear in a Python import, so this one-file package forwards `import smart_nar_fast_tts_b200` to it.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "smart-nar_fast_tts_b200"))
from ._pkg import *  # noqa: F401,F403,E402
from ._pkg import __all__  # noqa: E402
