"""heongpu_b200 -- B200-native (sm_100a) RNS-FHE arithmetic engine behind the
HEonGPU operator surface.  The compute path is the CUDA library
``heongpu_b200/lib/libheon_b200.so`` (C ABI in ``include/heon_b200.h``); this
package is the Python host-side mirror used by the tests and ``bench.py``.
There is no CPU fallback: importing :mod:`heongpu_b200.api` fails loudly when
the CUDA library has not been built.
"""
from .api import (  # noqa: F401
    HEContext,
    Ciphertext,
    Relinkey,
    Galoiskey,
    HEArithmeticOperator,
    HeonError,
    lib,
    build_library,
)
