"""Import the UNMODIFIED reference package from /root/reference (authoring container only).

bopy's modules import GPy, scipydirect, dppy, pyDOE and sobol_seq at module level
(bopy/surrogate.py:5, bopy/optimizer.py:6-7, bopy/initial_design.py:4-5); none of them is
installed here and none is on the ScipyGPSurrogate -> LCB/EI/POI path, so empty stand-ins are
registered in sys.modules before the import.  Used only by tools/make_golden.py.
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def import_reference():
    gpy = _stub("GPy")
    gpy.models = _stub("GPy.models", GPRegression=object)
    dppy = _stub("dppy")
    dppy.finite_dpps = _stub("dppy.finite_dpps", FiniteDPP=object)
    _stub("scipydirect", minimize=None)
    _stub("pyDOE")
    _stub("sobol_seq")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import bopy  # noqa: F401
    import bopy.acquisition
    import bopy.surrogate
    return bopy
