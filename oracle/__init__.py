"""CPU oracles for the reconstruction-loss hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; pointcloud_b200/ never does (tests/test_abi_cpu.py::test_product_package_never_touches_the_oracle enforces it).

  * emd_forward / emd_backward   -> oracle/emd_oracle.c   (reference: loss/emd/emd_cuda.cu)
  * chamfer_forward / _backward  -> oracle/chamfer_oracle.c (pytorch3d 0.7.2, PARITY UNPINNED)
  * loss_oracle.py               -> torch-CPU restatement of pointcloud_vision/utils.py:207-309
  * build_ref.py                 -> builds the unmodified reference EMD extension into oracle/_ref/
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("emd_oracle.c", "chamfer_oracle.c", "fps_oracle.c", "Makefile")]
    have_src = all(os.path.exists(s) for s in srcs)
    stale = (not os.path.exists(_SO)) or (have_src and any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.emd_oracle_forward.restype = ctypes.c_int
        _lib.emd_oracle_backward.restype = ctypes.c_int
        _lib.chamfer_oracle_forward.restype = ctypes.c_int
        _lib.chamfer_oracle_backward.restype = ctypes.c_int
        _lib.fps_oracle.restype = ctypes.c_int
    return _lib


def _np32(a):
    """torch tensor / array -> numpy float32 view (keeps strides; last dim must be contiguous)."""
    if hasattr(a, "detach"):
        a = a.detach().cpu().float().numpy()
    a = np.asarray(a, dtype=np.float32)
    if a.strides[-1] != 4:
        a = np.ascontiguousarray(a)
    return a


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _slices(b, nthreads):
    nthreads = max(1, min(int(nthreads), b)) if b > 0 else 1
    edges = np.linspace(0, b, nthreads + 1).astype(int)
    return [(int(edges[i]), int(edges[i + 1])) for i in range(nthreads) if edges[i + 1] > edges[i]]


def _run(fn, b, nthreads):
    sl = _slices(b, nthreads)
    if len(sl) <= 1:
        for lo, hi in sl:
            fn(lo, hi)
        return
    with ThreadPoolExecutor(len(sl)) as ex:  # ctypes releases the GIL during the C call
        list(ex.map(lambda s: fn(*s), sl))


def emd_forward(xyz1, xyz2, eps, iters, nthreads=1):
    """emdFunction.forward (emd_module.py:33-61).  Returns dict(dist, assignment, price, sum_unass, race_events, iters_run)."""
    x1, x2 = _np32(xyz1), _np32(xyz2)
    b, n, _ = x1.shape
    assert x2.shape[0] == b and x2.shape[1] == n
    dist = np.zeros((b, n), np.float32)
    asg = np.full((b, n), -1, np.int32)
    price = np.zeros((b, n), np.float32)
    su = np.zeros(b, np.int64)
    races = np.zeros(b, np.int32)
    itr = np.zeros(b, np.int32)
    L = lib()

    def work(lo, hi):
        rc = L.emd_oracle_forward(
            _ptr(x1[lo:], _f32p), ctypes.c_long(x1.strides[0] // 4), ctypes.c_long(x1.strides[1] // 4),
            _ptr(x2[lo:], _f32p), ctypes.c_long(x2.strides[0] // 4), ctypes.c_long(x2.strides[1] // 4),
            ctypes.c_int(hi - lo), ctypes.c_int(n), ctypes.c_float(eps), ctypes.c_int(iters),
            _ptr(dist[lo:], _f32p), _ptr(asg[lo:], _i32p), _ptr(price[lo:], _f32p),
            _ptr(su[lo:], ctypes.POINTER(ctypes.c_longlong)), _ptr(races[lo:], _i32p), _ptr(itr[lo:], _i32p))
        assert rc == 0, rc

    _run(work, b, nthreads)
    return dict(dist=dist, assignment=asg, price=price, sum_unass=su, race_events=races, iters_run=itr)


def emd_backward(xyz1, xyz2, assignment, graddist):
    """emdFunction.backward (emd_module.py:63-72): grad wrt xyz1 only; xyz2 gets zeros."""
    x1, x2 = _np32(xyz1), _np32(xyz2)
    b, n, _ = x1.shape
    asg = np.ascontiguousarray(np.asarray(assignment, dtype=np.int32))
    gd = np.ascontiguousarray(_np32(graddist))
    g1 = np.zeros((b, n, 3), np.float32)
    rc = lib().emd_oracle_backward(
        _ptr(x1, _f32p), ctypes.c_long(x1.strides[0] // 4), ctypes.c_long(x1.strides[1] // 4),
        _ptr(x2, _f32p), ctypes.c_long(x2.strides[0] // 4), ctypes.c_long(x2.strides[1] // 4),
        ctypes.c_int(b), ctypes.c_int(n), _ptr(asg, _i32p), _ptr(gd, _f32p), _ptr(g1, _f32p))
    assert rc == 0
    return g1, np.zeros_like(g1)


def _lens(lengths, b):
    if lengths is None:
        return None
    if hasattr(lengths, "detach"):
        lengths = lengths.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(lengths, dtype=np.int64))
    assert a.shape == (b,)
    return a


def chamfer_forward(x, y, x_lengths=None, y_lengths=None, mode=0, nthreads=1):
    """Two directed K=1 searches + pytorch3d's mean/mean reduction.  Returns dict(loss, loss_x, loss_y, dist_x, idx_x, dist_y, idx_y)."""
    xa, ya = _np32(x), _np32(y)
    b, p1, d = xa.shape
    assert ya.shape[0] == b and ya.shape[2] == d
    p2 = ya.shape[1]
    xl, yl = _lens(x_lengths, b), _lens(y_lengths, b)
    dist_x = np.zeros((b, p1), np.float32); idx_x = np.zeros((b, p1), np.int32)
    dist_y = np.zeros((b, p2), np.float32); idx_y = np.zeros((b, p2), np.int32)
    cham = np.zeros((b, 2), np.float64)
    L = lib()

    def work(lo, hi):
        rc = L.chamfer_oracle_forward(
            _ptr(xa[lo:], _f32p), ctypes.c_long(xa.strides[0] // 4), ctypes.c_long(xa.strides[1] // 4),
            _ptr(xl[lo:], _i64p) if xl is not None else None,
            _ptr(ya[lo:], _f32p), ctypes.c_long(ya.strides[0] // 4), ctypes.c_long(ya.strides[1] // 4),
            _ptr(yl[lo:], _i64p) if yl is not None else None,
            ctypes.c_int(hi - lo), ctypes.c_int(p1), ctypes.c_int(p2), ctypes.c_int(d), ctypes.c_int(mode),
            _ptr(dist_x[lo:], _f32p), _ptr(idx_x[lo:], _i32p), _ptr(dist_y[lo:], _f32p), _ptr(idx_y[lo:], _i32p),
            _ptr(cham[lo:], _f64p))
        if rc != 0:
            raise ValueError(f"chamfer_oracle_forward rc={rc}")

    _run(work, b, nthreads)
    nb = max(b, 1)
    lx, ly = cham[:, 0].sum() / nb, cham[:, 1].sum() / nb
    return dict(loss=np.float32(lx + ly), loss_x=np.float32(lx), loss_y=np.float32(ly),
                dist_x=dist_x, idx_x=idx_x, dist_y=dist_y, idx_y=idx_y)


def chamfer_backward(x, y, idx_x, idx_y, g=1.0, x_lengths=None, y_lengths=None):
    xa, ya = _np32(x), _np32(y)
    b, p1, d = xa.shape
    p2 = ya.shape[1]
    xl, yl = _lens(x_lengths, b), _lens(y_lengths, b)
    ix = np.ascontiguousarray(np.asarray(idx_x, dtype=np.int32)); iy = np.ascontiguousarray(np.asarray(idx_y, dtype=np.int32))
    gx = np.zeros((b, p1, d), np.float32); gy = np.zeros((b, p2, d), np.float32)
    rc = lib().chamfer_oracle_backward(
        _ptr(xa, _f32p), ctypes.c_long(xa.strides[0] // 4), ctypes.c_long(xa.strides[1] // 4),
        _ptr(xl, _i64p) if xl is not None else None,
        _ptr(ya, _f32p), ctypes.c_long(ya.strides[0] // 4), ctypes.c_long(ya.strides[1] // 4),
        _ptr(yl, _i64p) if yl is not None else None,
        ctypes.c_int(b), ctypes.c_int(p1), ctypes.c_int(p2), ctypes.c_int(d), _ptr(ix, _i32p), _ptr(iy, _i32p),
        ctypes.c_float(g), _ptr(gx, _f32p), _ptr(gy, _f32p))
    assert rc == 0
    return gx, gy


def fps(xyz, npoint, start=None, skip_origin=False):
    """farthest_point_sample(xyz, npoint) (models/pointnet2_utils.py:89-90) -> int32 (B, npoint); see fps_oracle.c."""
    x = _np32(xyz)
    b, n, _ = x.shape
    idx = np.zeros((b, npoint), np.int32)
    st = None if start is None else np.ascontiguousarray(np.asarray(start, dtype=np.int32))
    rc = lib().fps_oracle(_ptr(x, _f32p), ctypes.c_long(x.strides[0] // 4), ctypes.c_long(x.strides[1] // 4), ctypes.c_int(b),
                          ctypes.c_int(n), ctypes.c_int(npoint), _ptr(st, _i32p) if st is not None else None,
                          ctypes.c_int(1 if skip_origin else 0), _ptr(idx, _i32p))
    assert rc == 0
    return idx


def ball_query(radius, nsample, xyz, new_xyz):
    """query_ball_point (models/pointnet2_utils.py:93-113) restated with the difference-form squared distance
    ((dx*dx + dy*dy) + dz*dz in fp32; the reference uses the matmul expansion, which differs by ~1e-7 absolute, so
    parity with it is exact except for points that close to the sphere -- tests/golden marks those centroids)."""
    x, c = _np32(xyz), _np32(new_xyz)
    b, n, _ = x.shape
    s = c.shape[1]
    r2 = np.float32(radius ** 2)
    out = np.empty((b, s, nsample), np.int32)
    for i in range(b):
        d = c[i][:, None, :3] - x[i][None, :, :3]                      # (s, n, 3) fp32
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        inside = ~(d2 > r2)
        for q in range(s):
            hits = np.nonzero(inside[q])[0]
            first = hits[0] if len(hits) else n
            row = np.full(nsample, first, np.int32)
            row[: min(len(hits), nsample)] = hits[:nsample]
            out[i, q] = row
    return out
