"""ctypes loader for liborb_b200.so (the C ABI declared in include/orb_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable the
import / first call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ORB_B200_LIB") or os.path.join(_HERE, "liborb_b200.so")      # override: kernel-variant experiments only

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

ORB_OK, ORB_ERR_INVALID, ORB_ERR_GEOMETRY, ORB_ERR_CAPACITY, ORB_ERR_CUDA, ORB_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
GRID_COLS, GRID_ROWS = 64, 48

# every symbol include/orb_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "orb_error_string", "orb_last_cuda_error", "orb_abi_version", "orb_create", "orb_default_context", "orb_destroy", "orb_nlevels",
    "orb_scale_factor", "orb_keypoint_capacity", "orb_set_descriptor_fma", "orb_extract", "orb_extract_batch", "orb_extract_batch_device", "orb_extract_batch_async", "orb_wait",
    "orb_last_launch_count", "orb_profile_enable", "orb_profile_read", "orb_profile_stage_name", "orb_debug_level_info", "orb_debug_level_plane", "orb_descriptor_distance",
    "orb_hamming_knn2", "orb_hamming_knn2_device", "orb_set_knn_engine", "orb_knn2_merge_device", "orb_match_ratio",
    "orb_frame_grid_build", "orb_search_by_projection", "orb_search_window", "orb_search_window_best", "orb_search_for_initialization", "orb_search_by_bow", "orb_search_by_bow_kf", "orb_search_for_triangulation", "orb_host_alloc", "orb_host_free", "orb_host_alloc_input",
    "orb_measure_popc_peak",
    "orb_distinctive_descriptors", "orb_cvt_gray", "orb_extract_batch_color", "orb_undistort_keypoints", "orb_image_bounds",
    "orb_db_read_descriptors", "orb_db_write_descriptors", "orb_db_read_keypoints", "orb_db_write_keypoints",
    "orb_vocab_create", "orb_vocab_load_text", "orb_vocab_destroy", "orb_vocab_info", "orb_vocab_transform_features",
    "orb_vocab_transform_batch", "orb_bow_score_db", "orb_bow_detect_candidates",
    "orb_comm_init", "orb_comm_unique_id", "orb_comm_init_rank", "orb_comm_destroy", "orb_comm_size", "orb_comm_transport",
    "orb_comm_context", "orb_comm_set_extractor", "orb_comm_db_upload", "orb_comm_db_attach", "orb_knn2_sharded",
    "orb_knn2_sharded_device", "orb_extract_batch_multi",
]


class FrameView(C.Structure):
    _fields_ = [("n", C.c_int32), ("kps", C.c_void_p), ("desc", C.c_void_p),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("min_x", C.c_int32), ("max_x", C.c_int32), ("min_y", C.c_int32), ("max_y", C.c_int32),
                ("nlevels", C.c_int32), ("scale_factor", C.c_float),
                ("cell_start", C.c_void_p), ("cell_items", C.c_void_p)]


class WindowQuerySet(C.Structure):
    _fields_ = [("n", C.c_int32), ("active", C.c_void_p), ("desc", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("xyz", C.c_void_p), ("Tcw16", C.c_void_p), ("check_bounds", C.c_int32), ("radius", C.c_void_p),
                ("radius_const", C.c_float), ("min_level", C.c_void_p), ("max_level", C.c_void_p), ("angle", C.c_void_p)]


class FeatVecView(C.Structure):
    _fields_ = [("nnodes", C.c_int32), ("node_id", C.c_void_p), ("start", C.c_void_p), ("items", C.c_void_p)]


class OrbError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        L = lib()
        msg = L.orb_error_string(status).decode()
        if status == ORB_ERR_CUDA:
            msg += " [" + L.orb_last_cuda_error().decode() + "]"
        super().__init__("%s: %s (%d)" % (where, msg, status))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError("liborb_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C orbslam_jpminipc_b200/csrc`); there is no CPU fallback")
    L = C.CDLL(SO_PATH)
    vp, i32, f32, i64, sz = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t
    L.orb_error_string.restype = C.c_char_p
    L.orb_error_string.argtypes = [i32]
    L.orb_last_cuda_error.restype = C.c_char_p
    L.orb_create.restype = vp
    L.orb_create.argtypes = [i32, i32, f32, i32, i32, i32, i32, i32, i32]
    L.orb_destroy.argtypes = [vp]
    L.orb_default_context.restype = vp
    L.orb_default_context.argtypes = []
    L.orb_nlevels.argtypes = [vp]
    L.orb_scale_factor.restype = f32
    L.orb_scale_factor.argtypes = [vp]
    L.orb_keypoint_capacity.argtypes = [vp]
    L.orb_set_descriptor_fma.argtypes = [vp, C.c_int]
    L.orb_extract.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, C.POINTER(C.c_int)]
    L.orb_extract_batch.argtypes = [vp, vp, i32, i32, i32, i32, sz, vp, vp, i32, vp]
    L.orb_extract_batch_async.argtypes = [vp, vp, i32, i32, i32, i32, sz, vp, vp, i32, vp, C.POINTER(C.c_longlong)]
    L.orb_wait.argtypes = [vp, C.c_longlong]
    L.orb_extract_batch_device.argtypes = [vp, vp, i32, i32, i32, i32, sz, vp, vp, i32, vp, vp]
    L.orb_last_launch_count.argtypes = [vp]
    L.orb_profile_enable.argtypes = [vp, i32]
    L.orb_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.orb_profile_stage_name.restype = C.c_char_p
    L.orb_profile_stage_name.argtypes = [i32]
    L.orb_debug_level_info.argtypes = [vp, i32, i32, vp]
    L.orb_debug_level_plane.argtypes = [vp, i32, i32, i32, vp, sz]
    L.orb_descriptor_distance.argtypes = [vp, vp]
    L.orb_hamming_knn2.argtypes = [vp, vp, i32, vp, i64, vp, vp, vp]
    L.orb_hamming_knn2_device.argtypes = [vp, vp, i32, vp, i64, i32, i32, vp, vp, vp, vp]
    L.orb_set_knn_engine.argtypes = [vp, i32]
    L.orb_knn2_merge_device.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    L.orb_match_ratio.argtypes = [vp, vp, vp, vp, i32, f32, i32, vp, C.POINTER(C.c_int)]
    L.orb_frame_grid_build.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    L.orb_search_by_projection.argtypes = [vp, C.POINTER(FrameView), C.POINTER(FrameView), vp, vp, vp, vp,
                                           f32, i32, vp, C.POINTER(C.c_int)]
    L.orb_search_window.argtypes = [vp, C.POINTER(FrameView), C.POINTER(WindowQuerySet), i32, f32, i32, i32, vp, C.POINTER(C.c_int)]
    L.orb_search_window_best.argtypes = [vp, C.POINTER(FrameView), C.POINTER(WindowQuerySet), vp, vp]
    L.orb_search_for_initialization.argtypes = [vp, C.POINTER(FrameView), C.POINTER(FrameView), vp, i32, f32, i32, vp, C.POINTER(C.c_int)]
    L.orb_search_by_bow.argtypes = [vp, C.POINTER(FeatVecView), vp, vp, vp, i32,
                                    C.POINTER(FeatVecView), vp, vp, i32, f32, i32, vp, C.POINTER(C.c_int)]
    L.orb_search_by_bow_kf.argtypes = [vp, C.POINTER(FeatVecView), vp, vp, vp, i32,
                                       C.POINTER(FeatVecView), vp, vp, vp, i32, f32, i32, vp, C.POINTER(C.c_int)]
    L.orb_search_for_triangulation.argtypes = [vp, C.POINTER(FeatVecView), vp, vp, vp, i32, C.POINTER(FeatVecView), vp, vp, vp, i32, vp, vp, i32, i32,
                                               vp, C.POINTER(C.c_int)]
    L.orb_host_alloc.restype = vp
    L.orb_host_alloc.argtypes = [sz]
    L.orb_host_alloc_input.restype = vp
    L.orb_host_alloc_input.argtypes = [sz]
    L.orb_host_free.argtypes = [vp]
    L.orb_measure_popc_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.orb_distinctive_descriptors.argtypes = [vp, vp, vp, i32, vp, vp]
    L.orb_cvt_gray.argtypes = [vp, vp, i32, i32, i32, sz, sz, i32, vp, sz, sz]
    L.orb_extract_batch_color.argtypes = [vp, vp, i32, i32, i32, sz, sz, i32, vp, vp, i32, vp]
    L.orb_undistort_keypoints.argtypes = [vp, vp, i32, f32, f32, f32, f32, vp, i32, vp]
    L.orb_image_bounds.argtypes = [vp, i32, i32, f32, f32, f32, f32, vp, i32, vp]
    L.orb_db_read_descriptors.argtypes = [C.c_char_p, vp, i64, vp, i32, C.POINTER(i64), C.POINTER(C.c_int32)]
    L.orb_db_read_keypoints.argtypes = [C.c_char_p, vp, i64, vp, i32, C.POINTER(i64), C.POINTER(C.c_int32)]
    L.orb_db_write_descriptors.argtypes = [C.c_char_p, vp, vp, i32]
    L.orb_db_write_keypoints.argtypes = [C.c_char_p, vp, vp, i32]
    L.orb_vocab_create.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, C.POINTER(vp)]
    L.orb_vocab_load_text.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.orb_vocab_destroy.restype = None
    L.orb_vocab_destroy.argtypes = [vp]
    L.orb_vocab_info.argtypes = [vp] + [C.POINTER(C.c_int)] * 4
    L.orb_vocab_transform_features.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
    L.orb_vocab_transform_batch.argtypes = [vp, vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.orb_bow_score_db.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, i32, vp, vp, C.POINTER(C.c_int)]
    L.orb_bow_detect_candidates.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, i32, C.c_float, vp, vp, vp, vp, vp, C.POINTER(C.c_int)]
    L.orb_comm_init.restype = vp
    L.orb_comm_init.argtypes = [i32]
    L.orb_comm_unique_id.argtypes = [vp]
    L.orb_comm_init_rank.restype = vp
    L.orb_comm_init_rank.argtypes = [vp, i32, i32, vp]
    L.orb_comm_destroy.restype = None
    L.orb_comm_destroy.argtypes = [vp]
    L.orb_comm_size.argtypes = [vp]
    L.orb_comm_transport.restype = C.c_char_p
    L.orb_comm_transport.argtypes = [vp]
    L.orb_comm_context.restype = vp
    L.orb_comm_context.argtypes = [vp, i32]
    L.orb_comm_set_extractor.argtypes = [vp, i32, f32, i32, i32, i32, i32, i32, i32]
    L.orb_comm_db_upload.argtypes = [vp, vp, i64]
    L.orb_comm_db_attach.argtypes = [vp, i32, vp, i64, i64]
    L.orb_knn2_sharded.argtypes = [vp, vp, i32, vp, vp, vp]
    L.orb_knn2_sharded_device.argtypes = [vp, vp, i32, vp, i64, i64, vp, vp, vp, vp]
    L.orb_extract_batch_multi.argtypes = [vp, vp, i32, i32, i32, i32, sz, vp, vp, i32, vp]
    _lib = L
    return L


def check(status, where):
    if status != ORB_OK:
        raise OrbError(status, where)


def ptr(a):
    """numpy array -> host pointer, torch tensor -> its data pointer, int -> as is."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())      # torch tensor
