"""Build the UNMODIFIED reference EMD extension into oracle/_ref/ (test infrastructure only).

The sources are compiled where they lie under /root/reference (never copied into this
repo): pointcloud_vision/loss/emd/emd.cpp + emd_cuda.cu, the pybind module `emd` that
pointcloud_vision/loss/emd/emd_module.py:25 imports.  The reference's own setup.py
(loss/emd/setup.py:5-13) passes no arch flags; we only add the sm_100a gencode so the cubin
runs on a B200.  Output: oracle/_ref/emd.so (git-ignored, travels to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's reference legs may load oracle/_ref.
"""
import os
import sys

REF_DIR = "/root/reference/pointcloud_vision/loss/emd"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def build(verbose: bool = False) -> str | None:
    """Returns the path of the built module, or None when /root/reference is absent."""
    so = os.path.join(OUT_DIR, "emd.so")
    srcs = [os.path.join(REF_DIR, "emd.cpp"), os.path.join(REF_DIR, "emd_cuda.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return so if os.path.exists(so) else None
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    load(name="emd", sources=srcs, build_directory=OUT_DIR, is_python_module=False,
         extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"], verbose=verbose)
    return so if os.path.exists(so) else None


def load_ref():
    """Import the prebuilt reference module (GPU box: prebuilt file only)."""
    so = os.path.join(OUT_DIR, "emd.so")
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location("emd", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
