"""orbslam_jpminipc_b200 — B200-native (sm_100a) ORB front end behind the reference's
ORBextractor / ORBmatcher interface (caomw/ORBSLAM_jpMiniPC, ORB-SLAM v1).

Only the hot path lives here: csrc/ (hand-written CUDA kernels + the C ABI of
include/orb_b200.h) and thin host-side mirrors of the reference classes.  There is no CPU
fallback: importing works without a GPU, but every compute call needs liborb_b200.so and a
CUDA device and raises otherwise.
"""
from ._lib import KP_DTYPE, OrbError, SO_PATH, lib  # noqa: F401
from .extractor import ORBextractor  # noqa: F401
from .matcher import Frame, ORBmatcher  # noqa: F401
from .vocabulary import ORBVocabulary  # noqa: F401
