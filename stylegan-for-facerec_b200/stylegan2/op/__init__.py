"""The op-level drop-in boundary: the three names the reference's model.py imports from its
`op` package (model.py:7), here bound to the sm_100a kernels behind the C ABI (include/sg2_b200.h) instead of the two
JIT-compiled pybind extensions."""
from . import fused_act as _fused_act
from . import upfirdn2d as _upfirdn2d_module

FusedLeakyReLU = _fused_act.FusedLeakyReLU
fused_leaky_relu = _fused_act.fused_leaky_relu
upfirdn2d = _upfirdn2d_module.upfirdn2d      # like the reference, the package attribute `upfirdn2d` is the FUNCTION

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d"]
