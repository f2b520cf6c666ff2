"""simplefluid_b200 -- B200-native SPH solver step behind SimpleFluid's solver-facing API.

The product is the C-ABI shared library ``simplefluid_b200/lib/libsf_b200.so`` (hand-written
sm_100a CUDA kernels; declared in ``include/sf_b200.h``) plus the C++ host facade in
``simplefluid_b200/host/`` (QtSPHSolver / Simulator / SceneManager shaped).  This Python package is
only a ctypes binding of that C-ABI, used by the tests and by ``bench.py``; it mirrors the reference
method names (``makeReady``, ``advanceFrame``, ``getParticles`` ...).

There is no CPU fallback: importing works anywhere, but every compute call raises ``SFError`` when
the library or a B200-class GPU is missing.
"""
from .binding import (  # noqa: F401
    PinnedArray, SCENES, SFError, SFParams, SPHSolver, build_library, default_params, library, library_path, scene_generate,
)

HAS_SLAB = True  # multi-GPU z-slab decomposition is built into the library

__all__ = ["HAS_SLAB", "PinnedArray", "SCENES", "SFError", "SFParams", "SPHSolver", "build_library", "default_params", "library",
           "library_path", "scene_generate"]
