"""B200-native LLaMA decoder hot path behind the reference's plugin / operator surface.

Layout: ``csrc/`` hand-written sm_100a kernels + C ABI + plugin and runtime C++;
``_lib.py`` ctypes binding of ``include/*.h``; ``ops.py`` (operator wrappers of
T/tensorrt_llm/functional.py / quantization/functional.py), ``quantization.py``, ``runtime.py``
(GenerationSession, SamplingConfig, KVCacheManager), ``builder.py``, ``ft_format.py`` and
``calibration.py`` mirror the reference's Python interface for this path only.
"""
from ._lib import lib, load_library, LibraryNotBuilt  # noqa: F401

__version__ = "0.1.0"
