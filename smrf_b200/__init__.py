"""smrf_b200 -- B200-native (sm_100a) drop-in for SMRF's spatial distribution hot path.

    from smrf_b200.spatial import IDW, DK, GRID      # same constructors / methods as smrf.spatial.*
    from smrf_b200.utils import set_min_max

All arithmetic runs in libsmrfb.so (CUDA, C-ABI in include/smrfb.h).  Build it with
`python -m smrf_b200.build`.  There is no CPU fallback.
"""
__version__ = "0.1.0"
