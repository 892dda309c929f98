#!/usr/bin/env python
"""Recipe: compile the REFERENCE's own two CUDA extensions, from the sources where they lie under
/root/reference/backbone/stylegan2/op, into oracle/_ref/ (git-ignored, travels to the GPU box).

Nothing is copied: nvcc/g++ read the reference sources in place, only the built pybind modules
(`fused.so`, `upfirdn2d.so`) land in oracle/_ref/.  They are the reference's native kernels
(fused_bias_act_kernel.cu:18-49, upfirdn2d_kernel.cu:52-137) and are used ONLY as a checker by
`tests/test_ref_ext_gpu.py` on the GPU box: our kernels vs the reference's kernels, same inputs.
The reference's build system is `torch.utils.cpp_extension.load` at import time
(op/fused_act.py:9-15, op/upfirdn2d.py:8-14); this recipe calls the same function with an explicit
build directory and the sm_100a arch flag, nothing else.
"""
import os
import sys

SRC = "/root/reference/backbone/stylegan2/op"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def main():
    if not os.path.isdir(SRC):
        print("reference sources not present; keeping prebuilt oracle/_ref as is")
        return 0
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    for name, files in (("fused", ["fused_bias_act.cpp", "fused_bias_act_kernel.cu"]),
                        ("upfirdn2d", ["upfirdn2d.cpp", "upfirdn2d_kernel.cu"])):
        if os.path.exists(os.path.join(OUT, name + ".so")):
            continue
        load(name, sources=[os.path.join(SRC, f) for f in files], build_directory=OUT,
             is_python_module=False, verbose=False)
    for junk in os.listdir(OUT):   # keep only the shared objects
        if not junk.endswith(".so"):
            p = os.path.join(OUT, junk)
            if os.path.isfile(p):
                os.remove(p)
    print("built:", sorted(os.listdir(OUT)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
