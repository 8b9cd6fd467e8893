"""astre_b200 -- B200-native transform + visibility path for the Astre engine.

The product is ``libastre_gpu.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/astre_gpu.h``).  This package is the thin Python host side used by the
tests, the benchmark and multi-GPU orchestration; it never computes results on
the CPU and raises if the CUDA library is missing.
"""
from ._lib import load_library, library_path, AstreError  # noqa: F401
from .gpu import GpuContext, FRAME_TRANSFORM, FRAME_CULL, FRAME_SORT, FRAME_BASIS, FRAME_ALL  # noqa: F401
from . import scenes  # noqa: F401
