"""Import shim: the package directory is named `wavecap-sdr_b200/` (repo layout contract), which is
not a valid Python identifier. Importing `wavecap_sdr_b200` loads that directory as a package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "wavecap-sdr_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
