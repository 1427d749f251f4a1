"""Import shim: the product package lives in ``multi-modal-emotion_b200/`` (the name the build contract fixes), which
is not a valid Python identifier.  This package re-points its ``__path__`` there, so
``import multi_modal_emotion_b200.tav`` loads ``multi-modal-emotion_b200/tav.py``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "multi-modal-emotion_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
