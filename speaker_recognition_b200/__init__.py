"""Import shim: the package directory is named `speaker-recognition_b200/` (the repo layout the build
contract asks for), which is not a valid Python identifier.  `import speaker_recognition_b200` resolves
here and forwards the package path to the real directory, so submodules
(`speaker_recognition_b200.feature.MFCC`, `.gmm.pygmm`, `.gmmset`, `.interface`, ...) load from it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "speaker-recognition_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
