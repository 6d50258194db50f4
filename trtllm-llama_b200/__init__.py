"""B200-native LLaMA decoder hot path behind the reference's plugin / operator surface.

Layout: ``csrc/`` hand-written sm_100a kernels + C ABI + plugin and runtime C++;
``_lib.py`` ctypes binding of ``include/*.h``; ``plugin.py`` / ``functional.py`` /
``quantization.py`` / ``runtime.py`` mirror the reference's Python operator interface
(T/tensorrt_llm/{plugin,functional,quantization,runtime}) for this path only.
"""
from ._lib import lib, load_library, LibraryNotBuilt  # noqa: F401

__version__ = "0.1.0"
