"""Import shim: `import sg2_b200 as sg2` == importlib.import_module("stylegan-for-facerec_b200")."""
import importlib as _il

_pkg = _il.import_module("stylegan-for-facerec_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
__all__ = list(_pkg.__all__)
package = _pkg
