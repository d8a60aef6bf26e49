"""Build recipe for oracle/_ref: the reference's OWN CPU implementation of the hot path.

TEST INFRASTRUCTURE ONLY.  Compiles, unmodified and where they lie, the reference sources
  /root/reference/gammagl/mpops/torch_ext/src/*.cpp   (autograd Functions + pybind module, operators.cpp:51-59)
  /root/reference/gammagl/mpops/torch_ext/cpu/*.cpp   (serial E x K loops)
into oracle/_ref/_torch_ext.so with torch.utils.cpp_extension (the sources depend on libtorch only).
Flags follow what the reference's setup.py effectively ships for a CPU build (setup.py:83-88): no
COMPILE_WITH_OMP (the macro is misspelt there, setup.py:50), no COMPILE_WITH_CUDA, -O2.
The OpenMP variant is not an alternative: with -fopenmp -DCOMPILE_WITH_OMP the sources do not compile (tried here with
g++: `'idx' has not been declared` at cpu/segment_max_cpu.cpp:46, `#pragma omp critical` in an invalid position at
cpu/segment_sum_cpu.cpp:52, segment_mean_cpu.cpp:55,71, segment_max_cpu.cpp:56), so the single-threaded edge loops are the
only CPU implementation of the path the reference can actually ship, and the one bench.py times.

No reference source is copied into this repo; only the built .so lands in oracle/_ref/ (git-ignored,
NOT gpurun-ignored, so it travels to the GPU box).  If /root/reference is absent (GPU box) the
prebuilt file is used as is.
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/gammagl/mpops/torch_ext'
OUT_DIR = os.path.join(HERE, '_ref')
SO_PATH = os.path.join(OUT_DIR, '_torch_ext.so')


def build(force=False, verbose=False):
    if os.path.exists(SO_PATH) and not force:
        return SO_PATH
    if not os.path.isdir(REF_SRC):
        return None
    from torch.utils.cpp_extension import load
    os.makedirs(OUT_DIR, exist_ok=True)
    sources = sorted(glob.glob(REF_SRC + '/src/*.cpp') + glob.glob(REF_SRC + '/cpu/*.cpp'))
    load(name='_torch_ext', sources=sources,
         extra_include_paths=[REF_SRC + '/include', REF_SRC + '/cpu', REF_SRC],
         extra_cflags=['-O2', '-std=c++17'],
         build_directory=OUT_DIR, verbose=verbose)
    return SO_PATH if os.path.exists(SO_PATH) else None


def load_ref():
    """Import the built reference extension (returns the module or None)."""
    path = build()
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location('_torch_ext', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    p = build(force='--force' in sys.argv, verbose=True)
    print('oracle/_ref:', p)
