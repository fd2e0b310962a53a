"""ctypes binding of libxfeat_b200.so (include/xfeat_b200.h) -- what tests and bench.py call.

There is no fallback of any kind: if the shared library has not been built, or no sm_100 GPU is
present, constructing `XFeatB200` raises.  The library itself links only the CUDA runtime.
"""
import ctypes
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_long, c_size_t, c_uint8, c_void_p
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libxfeat_b200.so"
WEIGHTS_PATH = _PKG / "weights" / "xfeat_b200.bin"
INT_MAX = 2 ** 31 - 1

# every symbol include/xfeat_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "xfb_create", "xfb_destroy", "xfb_last_error", "xfb_set_stream", "xfb_extract", "xfb_extract_batch",
    "xfb_extract_batch_device", "xfb_distance_matrix", "xfb_distance_matrix_device", "xfb_distance_pairs", "xfb_distance_pairs_device",
    "xfb_match", "xfb_match_device", "xfb_image_bounds", "xfb_keypoint_geometry", "xfb_keypoint_geometry_device", "xfb_vocab_load", "xfb_bow_transform", "xfb_bow_transform_device", "xfb_bow_transform_frames_device",
    "xfb_submit", "xfb_wait", "xfb_match_frames", "xfb_match_frame_pairs", "xfb_match_frame_pairs_device", "xfb_profile_enable", "xfb_profile_read", "xfb_profile_tag_name",
    "xfb_debug_match_error", "xfb_debug_force_simt", "xfb_debug_read", "xfb_debug_read_stats", "xfb_debug_post", "xfb_debug_candidates", "xfb_launch_count",
]

_lib = None


def image_bounds(cam, w, h):
    """xfb_image_bounds: fills cam[10:14] (min_x, min_y, max_x, max_y) of a float32[14] camera array in place."""
    cam = np.ascontiguousarray(cam, np.float32)
    rc = load_library().xfb_image_bounds(_ptr(cam), int(w), int(h))
    if rc != 0:
        raise RuntimeError("xfb_image_bounds failed (%d)" % rc)
    return cam


def load_library(path=LIB_PATH):
    global _lib
    if _lib is not None:
        return _lib
    if not Path(path).exists():
        raise RuntimeError("libxfeat_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "-- there is no CPU / PyTorch fallback" % path)
    lib = ctypes.CDLL(str(path))
    lib.xfb_create.argtypes = [POINTER(c_void_p), c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int]
    lib.xfb_create.restype = c_int
    lib.xfb_destroy.argtypes = [c_void_p]
    lib.xfb_destroy.restype = None
    lib.xfb_last_error.argtypes = [c_void_p]
    lib.xfb_last_error.restype = c_char_p
    lib.xfb_set_stream.argtypes = [c_void_p, c_void_p]
    lib.xfb_extract.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.xfb_extract_batch.argtypes = [c_void_p, c_void_p, c_int, c_size_t, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                      c_void_p, c_void_p]
    lib.xfb_extract_batch_device.argtypes = lib.xfb_extract_batch.argtypes
    lib.xfb_distance_matrix.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]
    lib.xfb_distance_matrix_device.argtypes = lib.xfb_distance_matrix.argtypes
    lib.xfb_distance_pairs.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]
    lib.xfb_distance_pairs_device.argtypes = lib.xfb_distance_pairs.argtypes
    lib.xfb_image_bounds.argtypes = [c_void_p, c_int, c_int]
    lib.xfb_keypoint_geometry.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p] + [c_void_p] * 4
    lib.xfb_keypoint_geometry_device.argtypes = lib.xfb_keypoint_geometry.argtypes
    lib.xfb_vocab_load.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]
    lib.xfb_bow_transform.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    lib.xfb_bow_transform_device.argtypes = lib.xfb_bow_transform.argtypes
    lib.xfb_bow_transform_frames_device.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
    lib.xfb_match.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int] + [c_void_p] * 5
    lib.xfb_match_device.argtypes = lib.xfb_match.argtypes
    lib.xfb_submit.argtypes = [c_void_p, c_int, c_void_p, c_int, c_size_t, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_int] + [c_void_p] * 5
    lib.xfb_wait.argtypes = [c_void_p, c_int]
    lib.xfb_match_frames.argtypes = [c_void_p, c_int, c_int, c_int] + [c_void_p] * 5
    lib.xfb_match_frame_pairs.argtypes = [c_void_p, c_void_p, c_int, c_int] + [c_void_p] * 5
    lib.xfb_match_frame_pairs_device.argtypes = lib.xfb_match_frame_pairs.argtypes
    lib.xfb_profile_enable.argtypes = [c_void_p, c_int]
    lib.xfb_profile_read.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.xfb_profile_tag_name.argtypes = [c_int]
    lib.xfb_profile_tag_name.restype = c_char_p
    lib.xfb_debug_match_error.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]
    lib.xfb_debug_force_simt.argtypes = [c_void_p, c_int]
    lib.xfb_debug_read.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_size_t, c_void_p]
    lib.xfb_debug_read.restype = c_long
    lib.xfb_debug_read_stats.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_size_t]
    lib.xfb_debug_read_stats.restype = c_long
    lib.xfb_debug_post.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                   c_void_p]
    lib.xfb_debug_candidates.argtypes = [c_void_p, c_int]
    lib.xfb_launch_count.argtypes = [c_void_p]
    lib.xfb_launch_count.restype = c_long
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


class XFBError(RuntimeError):
    pass


class XFeatB200:
    """One context = one GPU, one host thread (mirrors one reference XFextractor instance)."""

    def __init__(self, max_h=480, max_w=640, max_batch=1, max_topk=4096, device=0, weights=WEIGHTS_PATH):
        self.lib = load_library()
        blob = Path(weights).read_bytes()
        h = c_void_p()
        rc = self.lib.xfb_create(ctypes.byref(h), blob, len(blob), device, max_h, max_w, max_batch, max_topk)
        if rc != 0:
            raise XFBError("xfb_create failed (%d): %s" % (rc, self.lib.xfb_last_error(None).decode()))
        self.h = h
        self.max_batch, self.max_topk = max_batch, max_topk

    def close(self):
        if getattr(self, "h", None):
            self.lib.xfb_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc, what):
        if rc < 0:
            raise XFBError("%s failed (%d): %s" % (what, rc, self.lib.xfb_last_error(self.h).decode()))
        return rc

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.xfb_set_stream(self.h, c_void_p(cuda_stream_ptr)), "xfb_set_stream")

    # ---- extract ---------------------------------------------------------------------------------
    def extract(self, frames, topk, nms_thr=0.05):
        """frames: uint8 [H,W] or [B,H,W] (host).  Returns dict of host arrays."""
        f = np.ascontiguousarray(frames, dtype=np.uint8)
        single = f.ndim == 2
        if single:
            f = f[None]
        B, H, W = f.shape
        nv = np.zeros(B, np.int32)
        xy = np.zeros((B, topk, 2), np.float32)
        sc = np.zeros((B, topk), np.float32)
        ds = np.zeros((B, topk, 64), np.float32)
        if single:
            rc = self.lib.xfb_extract(self.h, _ptr(f), H, W, W, topk, nms_thr, _ptr(nv), _ptr(xy), _ptr(sc), _ptr(ds))
        else:
            rc = self.lib.xfb_extract_batch(self.h, _ptr(f), B, H * W, H, W, W, topk, nms_thr, _ptr(nv), _ptr(xy), _ptr(sc), _ptr(ds))
        self._check(rc, "xfb_extract")
        out = {"n_valid": nv, "kpts": xy, "scores": sc, "desc": ds}
        if single:
            out = {k: v[0] for k, v in out.items()}
        return out

    def extract_ptrs(self, gray_ptr, batch, frame_stride, H, W, stride, topk, nms_thr, nv_ptr, xy_ptr, sc_ptr, ds_ptr, device=False):
        """Raw-pointer form (host pinned buffers or device tensors' data_ptr())."""
        fn = self.lib.xfb_extract_batch_device if device else self.lib.xfb_extract_batch
        self._check(fn(self.h, c_void_p(gray_ptr), batch, frame_stride, H, W, stride, topk, nms_thr, c_void_p(nv_ptr), c_void_p(xy_ptr),
                       c_void_p(sc_ptr), c_void_p(ds_ptr)), "xfb_extract_batch%s" % ("_device" if device else ""))

    # ---- match -----------------------------------------------------------------------------------
    def distance_matrix(self, A, B):
        A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
        out = np.zeros((A.shape[0], B.shape[0]), np.int32)
        self._check(self.lib.xfb_distance_matrix(self.h, _ptr(A), A.shape[0], _ptr(B), B.shape[0], _ptr(out)), "xfb_distance_matrix")
        return out

    def distance_pairs(self, A, B, idx_a, idx_b):
        """DescriptorDistance(A[idx_a[p]], B[idx_b[p]]) for a list of pairs (xfb_distance_pairs)."""
        A = np.ascontiguousarray(A, np.float32).reshape(-1, 64); B = np.ascontiguousarray(B, np.float32).reshape(-1, 64)
        ia = np.ascontiguousarray(idx_a, np.int32); ib = np.ascontiguousarray(idx_b, np.int32)
        out = np.zeros(ia.shape[0], np.int32)
        self._check(self.lib.xfb_distance_pairs(self.h, _ptr(A), A.shape[0], _ptr(B), B.shape[0], _ptr(ia), _ptr(ib), ia.shape[0], _ptr(out)),
                    "xfb_distance_pairs")
        return out

    def keypoint_geometry(self, xy, depth, cam):
        """(un_xy, depth, uright, cell) per keypoint (xfb_keypoint_geometry); cam: float32[14] = fx fy cx cy k1 k2 p1 p2 k3 bf minx miny maxx maxy."""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); n = xy.shape[0]
        cam = np.ascontiguousarray(cam, np.float32)
        un = np.zeros((n, 2), np.float32); kd = np.zeros(n, np.float32); ur = np.zeros(n, np.float32); cell = np.zeros(n, np.int32)
        if depth is None:
            dp, h, w, st = None, 0, 0, 0
        else:
            depth = np.ascontiguousarray(depth, np.float32); dp, (h, w), st = _ptr(depth), depth.shape, depth.shape[1]
        self._check(self.lib.xfb_keypoint_geometry(self.h, _ptr(xy), n, dp, h, w, st, _ptr(cam), _ptr(un), _ptr(kd), _ptr(ur), _ptr(cell)),
                    "xfb_keypoint_geometry")
        return un, kd, ur, cell

    def vocab_load(self, node_desc, child_start, child_index, L):
        nd = np.ascontiguousarray(node_desc, np.uint8).reshape(-1, 32)
        cs = np.ascontiguousarray(child_start, np.int32); ci = np.ascontiguousarray(child_index, np.int32)
        self._check(self.lib.xfb_vocab_load(self.h, _ptr(nd), _ptr(cs), _ptr(ci), nd.shape[0], ci.shape[0], int(L)), "xfb_vocab_load")

    def bow_transform(self, desc, levelsup):
        """(leaf node, node at level L - levelsup) per descriptor row (xfb_bow_transform)."""
        D = np.ascontiguousarray(desc, np.float32).reshape(-1, 64)
        leaf = np.zeros(D.shape[0], np.int32); nid = np.zeros(D.shape[0], np.int32)
        self._check(self.lib.xfb_bow_transform(self.h, _ptr(D), D.shape[0], int(levelsup), _ptr(leaf), _ptr(nid)), "xfb_bow_transform")
        return leaf, nid

    def bow_transform_frames(self, levelsup, leaf_ptr, nid_ptr):
        self._check(self.lib.xfb_bow_transform_frames_device(self.h, int(levelsup), c_void_p(leaf_ptr), c_void_p(nid_ptr)), "xfb_bow_transform_frames_device")

    def match(self, A, B, group_a=None, group_b=None, init=INT_MAX):
        A = np.ascontiguousarray(A, np.float32).reshape(-1, 64); B = np.ascontiguousarray(B, np.float32).reshape(-1, 64)
        n1, n2 = A.shape[0], B.shape[0]
        ga = None if group_a is None else np.ascontiguousarray(group_a, np.int32)
        gb = None if group_b is None else np.ascontiguousarray(group_b, np.int32)
        bi = np.full(n1, -1, np.int32); bd = np.full(n1, init, np.int32); sd = np.full(n1, init, np.int32)
        ri = np.full(n2, -1, np.int32); rd = np.full(n2, init, np.int32)
        self._check(self.lib.xfb_match(self.h, _ptr(A), n1, _ptr(B), n2, _ptr(ga), _ptr(gb), int(init), _ptr(bi), _ptr(bd), _ptr(sd),
                                       _ptr(ri), _ptr(rd)), "xfb_match")
        return bi, bd, sd, ri, rd

    def match_ptrs(self, a_ptr, n1, b_ptr, n2, init, bi, bd, sd, ri, rd, ga=None, gb=None):
        v = lambda p: c_void_p(p) if p else None
        self._check(self.lib.xfb_match_device(self.h, v(a_ptr), n1, v(b_ptr), n2, v(ga), v(gb), int(init), v(bi), v(bd), v(sd), v(ri),
                                              v(rd)), "xfb_match_device")

    def match_frames(self, fa, fb, topk, init=INT_MAX, ptrs=None):
        """Match two frames of the last extract call on the device; host int32 arrays of `topk`."""
        if ptrs is None:
            outs = [np.zeros(topk, np.int32) for _ in range(5)]
            ptrs = [_ptr(o) for o in outs]
        else:
            outs = None
        self._check(self.lib.xfb_match_frames(self.h, fa, fb, int(init), *ptrs), "xfb_match_frames")
        return outs

    def submit(self, slot, gray_ptr, batch, frame_stride, H, W, stride, topk, nms_thr, nv_ptr, xy_ptr, sc_ptr, ds_ptr, pairs=None, init=INT_MAX,
               match_ptrs=(0, 0, 0, 0, 0)):
        """Pipelined extract (+ frame-pair matches): host pointers (pinned), returns immediately; see xfb_wait."""
        v = lambda q: c_void_p(q) if q else None
        npairs = 0 if pairs is None else int(pairs.shape[0])
        pp = None if pairs is None else _ptr(np.ascontiguousarray(pairs, np.int32))
        self._check(self.lib.xfb_submit(self.h, slot, c_void_p(gray_ptr), batch, frame_stride, H, W, stride, topk, nms_thr, v(nv_ptr), v(xy_ptr),
                                        v(sc_ptr), v(ds_ptr), pp, npairs, int(init), *[v(q) for q in match_ptrs]), "xfb_submit")

    def wait(self, slot):
        self._check(self.lib.xfb_wait(self.h, slot), "xfb_wait")

    def match_frame_pairs(self, pairs, init, out_ptrs, device=False):
        """pairs: int32 [n,2] host array; out_ptrs: 5 raw pointers (host, or device when device=True) or 0."""
        pairs = np.ascontiguousarray(pairs, np.int32)
        fn = self.lib.xfb_match_frame_pairs_device if device else self.lib.xfb_match_frame_pairs
        v = [c_void_p(p) if p else None for p in out_ptrs]
        self._check(fn(self.h, _ptr(pairs), pairs.shape[0], int(init), *v), "xfb_match_frame_pairs")

    def profile(self, enable):
        self._check(self.lib.xfb_profile_enable(self.h, int(bool(enable))), "xfb_profile_enable")

    def profile_read(self):
        """{tag name: (total ms, launches)} since the last read."""
        ms = np.zeros(40, np.float32); cnt = np.zeros(40, np.int32)
        self._check(self.lib.xfb_profile_read(self.h, _ptr(ms), _ptr(cnt)), "xfb_profile_read")
        out = {}
        for t in range(40):
            if cnt[t]:
                out[self.lib.xfb_profile_tag_name(t).decode()] = (float(ms[t]), int(cnt[t]))
        return out

    # ---- introspection ----------------------------------------------------------------------------
    def debug_read(self, name, frame=0, capacity=None):
        cap = capacity or (self._max_elems())
        buf = np.zeros(cap, np.float32)
        dims = np.zeros(4, np.int32)
        n = self._check(self.lib.xfb_debug_read(self.h, name.encode(), frame, _ptr(buf), cap, _ptr(dims)), "xfb_debug_read(%s)" % name)
        return buf[:n].reshape(int(dims[0]), int(dims[1]), int(dims[2])).copy()

    def _max_elems(self):
        return 1408 * 1408 * 4

    def debug_stats(self, name, frame=0):
        buf = np.zeros(256, np.float32)
        n = self._check(self.lib.xfb_debug_read_stats(self.h, name.encode(), frame, _ptr(buf), 256), "xfb_debug_read_stats")
        c = n // 2
        return buf[:c].copy(), buf[c:n].copy()

    def debug_post(self, feats_hwc, H1, K1h, topk, nms_thr=0.05):
        K1h = np.ascontiguousarray(K1h, np.float32)
        H, W = K1h.shape
        feats = np.ascontiguousarray(feats_hwc, np.float32); H1 = np.ascontiguousarray(H1, np.float32)
        nv = np.zeros(1, np.int32); xy = np.zeros((topk, 2), np.float32); sc = np.zeros(topk, np.float32); ds = np.zeros((topk, 64), np.float32)
        self._check(self.lib.xfb_debug_post(self.h, H, W, _ptr(feats), _ptr(H1), _ptr(K1h), topk, nms_thr, _ptr(nv), _ptr(xy), _ptr(sc),
                                            _ptr(ds)), "xfb_debug_post")
        return {"n_valid": int(nv[0]), "kpts": xy, "scores": sc, "desc": ds}

    def match_error(self, A, B):
        A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
        e = np.zeros(1, np.float32)
        self._check(self.lib.xfb_debug_match_error(self.h, _ptr(A), A.shape[0], _ptr(B), B.shape[0], _ptr(e)), "xfb_debug_match_error")
        return float(e[0])

    def force_simt(self, enable):
        self._check(self.lib.xfb_debug_force_simt(self.h, int(bool(enable))), "xfb_debug_force_simt")

    def candidates(self, frame=0):
        return self._check(self.lib.xfb_debug_candidates(self.h, frame), "xfb_debug_candidates")

    def launch_count(self):
        return int(self.lib.xfb_launch_count(self.h))
