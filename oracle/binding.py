"""ctypes binding of the CPU oracle (oracle/uz_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under uzliti_slam_b200/ may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libuz_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "uz_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "CXX=g++"] + (["-B"] if force else []))
    return _SO


class Features(C.Structure):
    _fields_ = [("descriptors", C.c_void_p), ("positions", C.c_void_p), ("valid_3d", C.c_void_p),
                ("n", C.c_int32), ("desc_bytes", C.c_int32), ("desc_stride", C.c_int32),
                ("feature_type", C.c_int32), ("sensor_frame", C.c_int32)]


class Edge(C.Structure):
    _fields_ = [("ok", C.c_int32), ("cam_from", C.c_int32), ("cam_to", C.c_int32),
                ("n_ratio_matches", C.c_int32), ("n_matches", C.c_int32), ("consensus", C.c_int32),
                ("best_iteration", C.c_int32), ("iterations_run", C.c_int32),
                ("mse", C.c_double), ("info_scale", C.c_double), ("T", C.c_double * 16)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.uzo_estimate_edge.restype = None
        _lib.uzo_estimate_svd.restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def knn2(q, t):
    """(idx[nq,2], dist[nq,2]) int32; missing neighbours are -1."""
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    nq = q.shape[0]
    nb = q.shape[1] if q.ndim == 2 else t.shape[1]
    idx = np.empty((nq, 2), np.int32)
    dist = np.empty((nq, 2), np.int32)
    lib().uzo_knn2(_p(q), nq, nb, _p(t), t.shape[0], nb, nb, _p(idx), _p(dist))
    return idx, dist


def cross_match(q, t):
    """cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match(q, t) as (idx[nq], dist[nq]); -1 = query unmatched."""
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    nq = q.shape[0]
    nb = q.shape[1]
    idx = np.empty(nq, np.int32)
    dist = np.empty(nq, np.int32)
    lib().uzo_cross_match(_p(q), nq, nb, _p(t), t.shape[0], nb, nb, _p(idx), _p(dist))
    return idx, dist


def ratio_pass(d0, d1):
    return bool(lib().uzo_ratio_pass(int(d0), int(d1)))


def sample_list(M, iterations, do_prosac=True):
    out = np.zeros((iterations, 3), np.int32)
    if M >= 3:
        lib().uzo_sample_list(int(M), int(iterations), int(bool(do_prosac)), _p(out))
    return out


def glibc_rand(n, seed=1):
    out = np.empty(n, np.int32)
    lib().uzo_glibc_rand(C.c_uint(seed), n, _p(out))
    return out


def pose_svd(P, Q):
    """P, Q: (k,3) float64 rows = points (memory == Eigen 3xk column-major).  4x4 T with T*p ~= q."""
    P = np.ascontiguousarray(P, np.float64)
    Q = np.ascontiguousarray(Q, np.float64)
    T = np.empty(16, np.float64)
    lib().uzo_pose_svd(_p(P), _p(Q), P.shape[0], _p(T))
    return T.reshape(4, 4)


def consensus3d(P, Q, T, thr):
    P = np.ascontiguousarray(P, np.float64)
    Q = np.ascontiguousarray(Q, np.float64)
    T = np.ascontiguousarray(T, np.float64)
    mask = np.zeros(P.shape[0], np.uint8)
    c = lib().uzo_consensus3d(_p(P), _p(Q), P.shape[0], _p(T), C.c_double(thr), _p(mask))
    return c, mask.astype(bool)


def estimate_svd(P, Q, thr, iterations, bp, do_prosac=True, samples=None):
    P = np.ascontiguousarray(P, np.float64)
    Q = np.ascontiguousarray(Q, np.float64)
    M = P.shape[0]
    T = np.empty(16, np.float64)
    cons = C.c_int32()
    mse = C.c_double()
    bi = C.c_int32()
    ir = C.c_int32()
    mask = np.zeros(max(M, 1), np.uint8)
    sp = None
    if samples is not None:
        samples = np.ascontiguousarray(samples, np.int32)
        sp = _p(samples)
    lib().uzo_estimate_svd(_p(P), _p(Q), M, C.c_double(thr), int(iterations), C.c_double(bp),
                           int(bool(do_prosac)), sp, _p(T), C.byref(cons), C.byref(mse), _p(mask),
                           C.byref(bi), C.byref(ir))
    return dict(T=T.reshape(4, 4), consensus=cons.value, mse=mse.value, mask=mask[:M].astype(bool),
                best_iteration=bi.value, iterations_run=ir.value)


def make_features(cams):
    """cams: list of dicts(desc uint8[n,B], pos float64[n,3], valid uint8[n], feature_type, sensor_frame).
    Returns (ctypes array, keepalive)."""
    arr = (Features * max(len(cams), 1))()
    keep = []
    for i, c in enumerate(cams):
        d = np.ascontiguousarray(c["desc"], np.uint8)
        p = np.ascontiguousarray(c["pos"], np.float64)
        v = np.ascontiguousarray(c["valid"], np.uint8)
        keep += [d, p, v]
        n = d.shape[0]
        nb = d.shape[1] if d.ndim == 2 and n > 0 else 32
        arr[i] = Features(d.ctypes.data, p.ctypes.data, v.ctypes.data, n, nb, nb,
                          int(c.get("feature_type", 2)), int(c.get("sensor_frame", 0)))
    return arr, keep


def estimate_edge(cams_from, cams_to, thr=0.1, iterations=100, bp=0.6, do_prosac=True, min_keypoints=7,
                  want_debug=True, cross_check=False):
    """estimateEdgeDirect over two keyframes (lists of camera dicts)."""
    fa, k1 = make_features(cams_from)
    ta, k2 = make_features(cams_to)
    e = Edge()
    maxm = max([c["desc"].shape[0] for c in cams_to] + [1])
    matches = np.zeros((maxm, 3), np.int32)
    mask = np.zeros(maxm, np.uint8)
    counts = np.full(iterations, -1, np.int32)
    lib().uzo_estimate_edge(fa, len(cams_from), ta, len(cams_to), C.c_double(thr), int(iterations),
                            C.c_double(bp), int(bool(do_prosac)), int(min_keypoints), C.byref(e),
                            _p(matches) if want_debug else None, _p(mask) if want_debug else None, maxm,
                            _p(counts), int(bool(cross_check)))
    M = e.n_matches
    return dict(ok=bool(e.ok), cam_from=e.cam_from, cam_to=e.cam_to, n_ratio_matches=e.n_ratio_matches,
                n_matches=M, consensus=e.consensus, best_iteration=e.best_iteration,
                iterations_run=e.iterations_run, mse=e.mse, info_scale=e.info_scale,
                T=np.array(e.T[:], np.float64).reshape(4, 4), matches=matches[:M].copy(),
                inlier_mask=mask[:M].astype(bool), counts=counts)


def gate_edge(T, ok, consensus, min_score=20.0, max_T=1.5, max_R=30.0):
    """newEdgeCallback's numeric gate -> (accept, |t|, rotation in degrees)"""
    T = np.ascontiguousarray(T, np.float64).reshape(16)
    tn, rot = C.c_double(), C.c_double()
    a = lib().uzo_gate_edge(_p(T), int(bool(ok)), int(consensus), C.c_double(min_score), C.c_double(max_T), C.c_double(max_R),
                            C.byref(tn), C.byref(rot))
    return bool(a), tn.value, rot.value


def backproject(u, v, depth, fx=525.0, fy=525.0, cx=319.5, cy=239.5, max_depth=7.0, reverse=False):
    """extract3dFeatures -> (pos float64[n,3], valid uint8[n])"""
    u = np.ascontiguousarray(u, np.int32)
    v = np.ascontiguousarray(v, np.int32)
    depth = np.ascontiguousarray(depth, np.float32)
    n = len(u)
    pos = np.zeros((max(n, 1), 3), np.float64)
    valid = np.zeros(max(n, 1), np.uint8)
    lib().uzo_backproject.restype = None
    lib().uzo_backproject(_p(u), _p(v), n, _p(depth), depth.shape[1], depth.shape[1], depth.shape[0], C.c_double(fx),
                          C.c_double(fy), C.c_double(cx), C.c_double(cy), C.c_double(max_depth), int(bool(reverse)),
                          _p(pos), _p(valid))
    return pos[:n], valid[:n]


def wire_decode(blob, capacity=4096, cols=32):
    """FeatureData::fromMsg on serialised Feature[] bytes -> (desc uint8[n,cols], pos[n,3], valid[n], uv[n,2]) or None"""
    blob = np.frombuffer(bytes(blob), np.uint8)
    desc = np.zeros((capacity, cols), np.uint8)
    pos = np.zeros((capacity, 3), np.float64)
    valid = np.zeros(capacity, np.uint8)
    uv = np.zeros((capacity, 2), np.int32)
    c = C.c_int()
    n = lib().uzo_wire_decode(_p(blob), C.c_size_t(len(blob)), capacity, _p(desc), C.byref(c), _p(pos), _p(valid), _p(uv))
    if n < 0 or (n > 0 and c.value != cols):
        return None
    return desc[:n], pos[:n], valid[:n], uv[:n]


def wire_encode(cam, uv=None):
    """FeatureData::toMsg -> bytes of the serialised Feature[] field"""
    d = np.ascontiguousarray(cam["desc"], np.uint8)
    p = np.ascontiguousarray(cam["pos"], np.float64)
    v = np.ascontiguousarray(cam["valid"], np.uint8)
    n, cols = d.shape
    out = np.zeros(4 + n * (41 + 4 * cols), np.uint8)
    uvp = None
    if uv is not None:
        uv = np.ascontiguousarray(uv, np.int32)
        uvp = _p(uv)
    lib().uzo_wire_encode.restype = C.c_long
    got = lib().uzo_wire_encode(_p(d), n, cols, cols, _p(p), _p(v), uvp, _p(out), C.c_size_t(len(out)))
    assert got == len(out)
    return out.tobytes()


class Places:
    """Sequential CPU restatement of LshSetRecognizer behind PlaceRecognizer's filters (oracle/uz_oracle.cpp, 8f-1).
    ids are arbitrary integers (the tests use store handles), stamps are nanoseconds."""

    def __init__(self, T=2.0, k=20):
        L = lib()
        L.uzo_places_new.restype = C.c_void_p
        L.uzo_places_free.restype = None
        L.uzo_places_clear.restype = None
        L.uzo_places_add.restype = None
        L.uzo_places_remove.restype = None
        L.uzo_places_votes.restype = None
        self.h = C.c_void_p(L.uzo_places_new())
        self.T, self.k = float(T), int(k)

    def __del__(self):
        try:
            lib().uzo_places_free(self.h)
        except Exception:
            pass

    def clear(self):
        lib().uzo_places_clear(self.h)

    def _pairs(self, fn, id_, stamp_ns, cams):
        arr, keep = make_features(cams)
        cap = self.k + 8
        out = np.zeros((cap, 2), np.int64)
        n = fn(self.h, C.c_longlong(int(id_)), C.c_longlong(int(stamp_ns)), arr, len(cams), C.c_double(self.T),
               self.k, _p(out), cap)
        return out[:n].copy()

    def search_and_add(self, id_, stamp_ns, cams):
        return self._pairs(lib().uzo_places_search_and_add, id_, stamp_ns, cams)

    def search(self, id_, stamp_ns, cams):
        return self._pairs(lib().uzo_places_search, id_, stamp_ns, cams)

    def add(self, id_, stamp_ns, cams):
        arr, keep = make_features(cams)
        lib().uzo_places_add(self.h, C.c_longlong(int(id_)), C.c_longlong(int(stamp_ns)), arr, len(cams))

    def remove(self, id_):
        lib().uzo_places_remove(self.h, C.c_longlong(int(id_)))

    def votes(self, cam, n_places, filtered=False):
        arr, keep = make_features([cam])
        out = np.zeros(max(n_places, 1), np.int32)
        lib().uzo_places_votes(self.h, arr, int(bool(filtered)), _p(out))
        return out[:n_places]
