"""front_end_b200 -- B200-native stereo feature front-end (FAST -> top-N -> ORB / SURF -> stereo and
inter-frame matching) behind the C-ABI of include/fe_abi.h.  sm_100a CUDA only; no CPU fallback."""
from . import lib, shard
from .frontend import FrontEnd, get_lib
from .lib import (DESC_BRIEF16, DESC_BRIEF32, DESC_BRIEF64, DESC_FREAK, DESC_ORB256, DESC_SURF64, DESC_SURF128, FAST_5_8, FAST_7_12, FAST_9_16, KPOINT, MASK_EPIPOLAR,
                  MASK_NONE, MASK_WINDOW, MATCH, MATCH_CROSSCHECK, MATCH_RATIO, NORM_HAMMING, NORM_HAMMING2, NORM_L2, FeError,
                  match_cfg)

__all__ = ["FrontEnd", "get_lib", "lib", "shard", "match_cfg", "FeError", "KPOINT", "MATCH"]
