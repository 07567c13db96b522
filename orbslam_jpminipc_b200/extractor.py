"""ORBextractor — python mirror of ORB_SLAM::ORBextractor (reference include/ORBextractor.h:32-77)
on top of the C ABI.  Same constructor arguments, same call: extractor(image, mask) -> (keypoints,
descriptors); keypoints is a numpy record array bit-compatible with cv::KeyPoint, descriptors is
N x 32 uint8.  All compute happens in liborb_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, lib, ptr


class ORBextractor:
    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, scoreType=1, fastTh=20,
                 device=0, max_width=1920, max_height=1200, max_batch=64, desc_fma=False):
        self.nfeatures, self.nlevels, self.device = nfeatures, nlevels, device
        self.max_batch = max_batch
        self._h = lib().orb_create(device, nfeatures, scaleFactor, nlevels, scoreType, fastTh,
                                   max_width, max_height, max_batch)
        if not self._h:
            raise RuntimeError("orb_create failed: " + lib().orb_last_cuda_error().decode())
        self.capacity = lib().orb_keypoint_capacity(self._h)
        if desc_fma:                                         # reproduce a reference built with FMA contraction (orb_b200.h)
            check(lib().orb_set_descriptor_fma(self._h, 1), "orb_set_descriptor_fma")

    def close(self):
        if getattr(self, "_h", None):
            lib().orb_destroy(self._h)
            self._h = None

    __del__ = close

    def GetLevels(self):
        return lib().orb_nlevels(self._h)

    def GetScaleFactor(self):
        return lib().orb_scale_factor(self._h)

    # ---- frame plumbing around the extractor (reference src/Tracking.cc:200-212, src/Frame.cc:289-349)
    def cvt_gray(self, images, order="RGB"):
        """cvtColor(RGB2GRAY / BGR2GRAY) of Tracking::GrabImage for an (N,)H x W x 3 uint8 array."""
        from ._lib import check, ptr
        a = np.ascontiguousarray(images, np.uint8)
        single = a.ndim == 3
        if single:
            a = a[None]
        n, h, w, _ = a.shape
        out = np.zeros((n, h, w), np.uint8)
        check(lib().orb_cvt_gray(self._h, ptr(a), n, w, h, a.strides[1], a.strides[0], 0 if order == "RGB" else 1, ptr(out), w, w * h),
              "orb_cvt_gray")
        return out[0] if single else out

    def extract_color(self, image, order="RGB"):
        """Colour frame -> (keypoints, descriptors), like GrabImage + Frame::Frame."""
        from ._lib import check, ptr
        a = np.ascontiguousarray(image, np.uint8)
        h, w, _ = a.shape
        kps = np.zeros(self.capacity, KP_DTYPE); desc = np.zeros((self.capacity, 32), np.uint8); cnt = np.zeros(1, np.int32)
        check(lib().orb_extract_batch_color(self._h, ptr(a), 1, w, h, a.strides[0], a.strides[0] * h, 0 if order == "RGB" else 1,
                                            ptr(kps), ptr(desc), self.capacity, ptr(cnt)), "orb_extract_batch_color")
        return kps[:cnt[0]].copy(), desc[:cnt[0]].copy()

    def undistort_keypoints(self, kps, K, dist):
        """Frame::UndistortKeyPoints: K = (fx, fy, cx, cy), dist = (k1, k2, p1, p2[, k3 ...]) float32."""
        from ._lib import check, ptr
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        d = np.ascontiguousarray(dist, np.float32)
        out = np.zeros_like(kps)
        check(lib().orb_undistort_keypoints(self._h, ptr(kps), len(kps), K[0], K[1], K[2], K[3], ptr(d), len(d), ptr(out)),
              "orb_undistort_keypoints")
        return out

    def image_bounds(self, w, h, K, dist):
        """Frame::ComputeImageBounds -> (mnMinX, mnMaxX, mnMinY, mnMaxY)."""
        from ._lib import check, ptr
        d = np.ascontiguousarray(dist, np.float32)
        b = np.zeros(4, np.int32)
        check(lib().orb_image_bounds(self._h, w, h, K[0], K[1], K[2], K[3], ptr(d), len(d), ptr(b)), "orb_image_bounds")
        return b

    def extract_batch_async(self, images, kps, desc, counts):
        """Enqueue a batch (uint8 array N x H x W, ideally pinned) into caller-owned output arrays (N x capacity keypoints,
        N x capacity x 32 descriptors, N counts) and return a ticket for wait(); see orb_extract_batch_async."""
        import ctypes as C
        from ._lib import check, ptr
        n, h, w = images.shape
        t = C.c_longlong(-1)
        check(lib().orb_extract_batch_async(self._h, ptr(images), n, w, h, images.strides[1], images.strides[0], ptr(kps), ptr(desc),
                                            self.capacity, ptr(counts), C.byref(t)), "orb_extract_batch_async")
        return t.value

    def wait(self, ticket):
        from ._lib import check
        check(lib().orb_wait(self._h, ticket), "orb_wait")

    # ORBextractor::operator()(image, mask, keypoints, descriptors)
    def __call__(self, image, mask=None):
        image = np.asarray(image)
        if image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1, "CV_8UC1 expected"
        h, w = image.shape
        cap = self.capacity
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        check(lib().orb_extract(self._h, ptr(image), w, h, image.strides[0], ptr(kps), ptr(desc), cap, C.byref(n)),
              "orb_extract")
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images):
        """images: (n, h, w) uint8 host array -> list of (keypoints, descriptors)."""
        images = np.ascontiguousarray(images, np.uint8)
        n, h, w = images.shape
        cap = self.capacity
        kps = np.zeros((n, cap), KP_DTYPE)
        desc = np.zeros((n, cap, 32), np.uint8)
        counts = np.zeros(n, np.int32)
        check(lib().orb_extract_batch(self._h, ptr(images), n, w, h, w, h * w, ptr(kps), ptr(desc), cap, ptr(counts)),
              "orb_extract_batch")
        return [(kps[i, :counts[i]].copy(), desc[i, :counts[i]].copy()) for i in range(n)]

    def extract_batch_device(self, d_images, d_kps, d_desc, d_counts, stream=0):
        """Device-resident batch (torch CUDA tensors or raw device pointers); asynchronous on `stream`.
        d_images: (n, h, w) uint8; d_kps: (n, cap, 7) int32/float32 view of orb_keypoint; d_desc: (n, cap, 32) uint8."""
        n, h, w = d_images.shape
        cap = d_desc.shape[1]
        check(lib().orb_extract_batch_device(self._h, ptr(d_images), n, w, h, w, h * w, ptr(d_kps), ptr(d_desc), cap,
                                             ptr(d_counts), C.c_void_p(stream)), "orb_extract_batch_device")

    def last_launch_count(self):
        return lib().orb_last_launch_count(self._h)

    # inspection hooks used by the parity tests
    def level_info(self, level, frame=0):
        info = np.zeros(10, np.int32)
        check(lib().orb_debug_level_info(self._h, frame, level, ptr(info)), "orb_debug_level_info")
        return dict(zip(["w", "h", "stride", "nDesired", "cols", "rows", "cellW", "cellH", "nfCell", "nKept"],
                        [int(v) for v in info]))

    def level_plane(self, level, blurred=False, frame=0):
        i = self.level_info(level, frame)
        buf = np.zeros((i["h"] + 32, i["stride"]), np.uint8)
        check(lib().orb_debug_level_plane(self._h, frame, level, int(blurred), ptr(buf), buf.nbytes), "orb_debug_level_plane")
        return buf[:, :i["w"] + 32].copy()
