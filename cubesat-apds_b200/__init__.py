"""dunk_b200: B200-native (sm_100a) implementation of DUNK's registration hot path.

Sub-modules mirror the reference crates on the path:
  feature_extraction  (feature_extraction/src/lib.rs)
  homographier        (homographier/src/homographier/mod.rs)
  feature_database    (feature_database/src/{keypointdb,imagedb,elevationdb,models}.rs, read/load side)
  image_extractor     (geotiff_extractor/src/image_extractor/mod.rs: the radiometric pre-step only)
Everything numeric happens in libdunk_b200.so (csrc/, C ABI in include/dunk_b200.h).
"""
from . import _lib
from ._lib import (DMATCH_DTYPE, KEYPOINT_DTYPE, POSE_DTYPE, REGISTRATION_DTYPE, TOP2_DTYPE, Context, DeviceBuffer, DunkError,
                   PinnedBuffer, default_context)
from . import feature_extraction
from . import _extract
from . import feature_database
from . import homographier
from . import image_extractor

__all__ = ["_lib", "Context", "DunkError", "default_context", "feature_extraction", "feature_database", "homographier", "image_extractor",
           "DMATCH_DTYPE", "KEYPOINT_DTYPE", "TOP2_DTYPE", "POSE_DTYPE", "REGISTRATION_DTYPE", "DeviceBuffer", "PinnedBuffer"]
