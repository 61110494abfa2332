"""minicv_b200 -- B200-native match + two-view-estimation path behind MiniCV's native ABI.

The product is ``lib/libMiniCVNative.so`` (C ABI in ``include/MiniCVNative.h``, CUDA kernels
in ``csrc/``).  This Python package is the host-side mirror of the reference's F# wrappers
(``src/MiniCV/OpenCV.fs``) used by the tests and the benchmark; it never computes on the CPU.
"""
from . import native  # noqa: F401
from .matching import (ENGINE_AUTO, ENGINE_POPC, ENGINE_TENSOR, ENGINE_TENSOR_PACKED, KnnResult, NORM_HAMMING, NORM_L2,  # noqa: F401
                       gather_matches_device, match_features, match_knn2, set_hamming_engine,
                       match_knn2_batch, match_knn2_batch_device, match_knn2_batch_multi, set_l2_engine,
                       LargeDb, PinnedArray)
from . import estimation  # noqa: F401
from .estimation import (RecoverPoseConfig, findFundamentalMat, findHomography, fivepoint,  # noqa: F401
                         recoverPose, recoverPoses, recoverPoses2)
