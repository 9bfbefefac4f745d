"""Import alias for the product package.

The package directory is ``em-pose_b200/`` (the name the project layout prescribes), which is not
an importable identifier; this stub makes it importable as ``empose_b200`` by pointing
``__path__`` at that directory and executing its ``__init__``.  No code lives here.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'em-pose_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f
