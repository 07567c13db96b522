"""ctypes binding of oracle/_ref/libref_dropin.so with the SAME python surface as oracle/pyref.py.

libref_dropin.so = the reference's unmodified src/Frame.cc, KeyFrame.cc, MapPoint.cc, Map.cc, KeyFrameDatabase.cc and DBoW2 containers,
compiled with this repository's include/ORBextractor.h, ORBmatcher.h and ORBVocabulary.h swapped in for the reference's
(-DORB_B200_WITH_REFERENCE_TYPES, recipe in oracle/Makefile) and linked against liborb_b200.so: every call the reference's code makes
to those three classes runs on the GPU.  tests/test_gpu_dropin.py drives this module and oracle/pyref.py (the reference's own
ORBextractor.cc / ORBmatcher.cc) with identical inputs.  TEST INFRASTRUCTURE ONLY."""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("oracle._pyref_dropin", os.path.join(_HERE, "pyref.py"))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
_m._SO = os.path.join(_HERE, "_ref", "libref_dropin.so")
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})


def available():
    return os.path.exists(_m._SO)
