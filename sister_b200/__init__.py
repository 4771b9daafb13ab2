"""sister_b200 -- B200-native (sm_100a) 5-view disparity path behind the reference's helper-class API.

Python here is binding glue over the C ABI (include/sister_b200.h, built into sister_b200/libsister_b200.so by
sister_b200/csrc/Makefile). ``SisterMultiviewDisparities`` mirrors the reference class of the same name
(cpp/include/sister/SisterMultiviewDisparities.hpp:18-26): same constructor order, same method, same outputs.

There is no CPU fallback: if the CUDA library is missing or no sm_100 device is present, calls raise.
"""
from .api import (Engine, SisterError, SisterMultiviewDisparities, build_library, library_path,  # noqa: F401
                  MODE_ALL, MODE_HORIZONTAL, MODE_MULTIVIEW, MODE_VERTICAL, STAGE_NAMES)

__all__ = ["Engine", "SisterError", "SisterMultiviewDisparities", "build_library", "library_path",
           "MODE_ALL", "MODE_HORIZONTAL", "MODE_MULTIVIEW", "MODE_VERTICAL", "STAGE_NAMES"]
