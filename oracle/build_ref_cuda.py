#!/usr/bin/env python
"""Compile the REFERENCE's own CUDA correlation package for sm_100 into oracle/_ref/ (build container only).

TEST INFRASTRUCTURE.  The sources are compiled where they lie under /root/reference/models/
correlation_package (correlation_cuda.cc, correlation_cuda_kernel.cu); nothing is copied into the
repo, the output (oracle/_ref/correlation_cuda*.so) is git-ignored and travels to the GPU box.  The
shipped egg only carries sm_50..sm_90 SASS (setup.py:7-18), so it cannot run on a B200; this build adds
-gencode arch=compute_100,code=sm_100 and nothing else.  Used by tests/test_gpu_vs_reference_cuda.py
and benchmarks/ref_cuda_compare.py as "the existing GPU kernel".
"""
import os
import sys

from torch.utils.cpp_extension import load

REF = os.environ.get("USF_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def main():
    src = os.path.join(REF, "models", "correlation_package")
    if not os.path.isdir(src):
        print("reference sources not found:", src)
        return 1
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ["CC"], os.environ["CXX"] = "/usr/bin/gcc", "/usr/bin/g++"
    load(name="correlation_cuda", sources=[os.path.join(src, "correlation_cuda.cc"),
                                           os.path.join(src, "correlation_cuda_kernel.cu")],
         extra_cuda_cflags=["-gencode", "arch=compute_100,code=sm_100", "-ccbin", "/usr/bin/g++"],
         extra_cflags=["-O3"], build_directory=OUT, is_python_module=False, verbose=False)
    print("built:", [f for f in os.listdir(OUT) if f.endswith(".so")])
    return 0


if __name__ == "__main__":
    sys.exit(main())
