"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package rtpp_b200 never does (tests/test_boundary.py greps for that).

  * liboracle  = oracle/librtpp_oracle.so       — our CPU restatement (oracle/rtpp_oracle.hpp)
  * ref_*      = oracle/_ref/libref_<planner>.so — the reference's own planner sources compiled in
                 place with stand-in OpenCV/CGAL headers (built only where /root/reference exists;
                 the prebuilt .so travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(ref=True):
    """(Re)build the oracle; and oracle/_ref when the reference tree is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "librtpp_oracle.so"])
    if ref and os.path.exists("/root/reference/RTRRT/lib/RTrrt/RTRRT.cpp"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "librtpp_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _lib = C.CDLL(path)
        _lib.orc_edges_nprobes_sum.restype = C.c_longlong
        _lib.orc_rtrrt_create.restype = C.c_void_p
        _lib.orc_rtfmt_create.restype = C.c_void_p
        _lib.orc_amrrt_create.restype = C.c_void_p
        _lib.orc_rtrrt_clock_reads.restype = C.c_longlong
        _lib.orc_rtrrt_knn_mode.restype = C.c_longlong
        _lib.orc_rtfmt_start.restype = C.c_longlong
        _lib.orc_amrrt_rand_pos.restype = C.c_longlong
        _lib.orc_amrrt_diffusivity.restype = C.c_float
    return _lib


def ref_lib(name):
    """name in {'rtrrt','rtfmt','amrrt'}; returns None when oracle/_ref was not built."""
    path = os.path.join(HERE, "_ref", f"libref_{name}.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    getattr(L, f"ref_{name}_create").restype = C.c_void_p
    L.ref_rand_pos.restype = C.c_longlong
    L.ref_clock_reads.restype = C.c_longlong
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ------------------------------------------------------------------ primitives -----------------
def edges_check(S, D, xyxy, divisor, mask, nthreads=1):
    H, W = (S if S is not None else D).shape
    xyxy = _c(xyxy, np.float32).reshape(-1, 4)
    out = np.zeros(len(xyxy), np.uint8)
    lib().orc_edges_check_f32(_ptr(S), _ptr(D), W, H, _ptr(xyxy), C.c_longlong(len(xyxy)), divisor, mask, _ptr(out), nthreads)
    return out


def edges_first_hit(S, D, xyxy, nthreads=1):
    H, W = (S if S is not None else D).shape
    xyxy = _c(xyxy, np.float32).reshape(-1, 4)
    out = np.zeros(len(xyxy), np.uint8)
    ps = np.zeros((len(xyxy), 2), np.float32)
    lib().orc_edges_first_hit_f32(_ptr(S), _ptr(D), W, H, _ptr(xyxy), C.c_longlong(len(xyxy)), _ptr(out), _ptr(ps), nthreads)
    return out, ps


def edges_check_f64(S, D, xyxy, divisor, mask, nthreads=1):
    H, W = (S if S is not None else D).shape
    xyxy = _c(xyxy, np.float64).reshape(-1, 4)
    out = np.zeros(len(xyxy), np.uint8)
    lib().orc_edges_check_f64(_ptr(S), _ptr(D), W, H, _ptr(xyxy), C.c_longlong(len(xyxy)), divisor, mask, _ptr(out), nthreads)
    return out


def edges_nprobes(xyxy, divisor):
    xyxy = _c(xyxy, np.float32).reshape(-1, 4)
    out = np.zeros(len(xyxy), np.int32)
    lib().orc_edges_nprobes(_ptr(xyxy), C.c_longlong(len(xyxy)), divisor, _ptr(out))
    return out


def edges_nprobes_sum(xyxy, divisor):
    xyxy = _c(xyxy, np.float32).reshape(-1, 4)
    return int(lib().orc_edges_nprobes_sum(_ptr(xyxy), C.c_longlong(len(xyxy)), divisor))


def dist(xyxy):
    xyxy = _c(xyxy, np.float32).reshape(-1, 4)
    out = np.zeros(len(xyxy), np.float32)
    lib().orc_dist_f32(_ptr(xyxy), C.c_longlong(len(xyxy)), _ptr(out))
    return out


def rand31(seed, n, i0=0):
    out = np.zeros(n, np.int32)
    lib().orc_rand31_fill(C.c_uint64(seed), C.c_longlong(i0), C.c_longlong(n), _ptr(out))
    return out


def knn(pts, q, k, grid=None, nthreads=1):
    """grid=None: brute force; grid=(W,H,cell): the oracle's grid accelerator."""
    pts = _c(pts, np.float32).reshape(-1, 2)
    q = _c(q, np.float64).reshape(-1, 2)
    idx = np.zeros((len(q), k), np.int32)
    d2 = np.zeros((len(q), k), np.float64)
    cnt = np.zeros(len(q), np.int32)
    if grid is None:
        lib().orc_knn_brute(_ptr(pts), len(pts), _ptr(q), C.c_longlong(len(q)), k, _ptr(idx), _ptr(d2), _ptr(cnt), nthreads)
    else:
        lib().orc_knn_grid(_ptr(pts), len(pts), C.c_double(grid[0]), C.c_double(grid[1]), C.c_double(grid[2]),
                           _ptr(q), C.c_longlong(len(q)), k, _ptr(idx), _ptr(d2), _ptr(cnt), nthreads)
    return idx, d2, cnt


def knn_kdtree(pts, q, k, domain=(4096.0, 4096.0)):
    """The kd-tree form of the oracle's index (sliding midpoint, bucket 10), single-threaded: must equal knn()."""
    pts = _c(pts, np.float32).reshape(-1, 2)
    q = _c(q, np.float64).reshape(-1, 2)
    idx = np.zeros((len(q), k), np.int32)
    d2 = np.zeros((len(q), k), np.float64)
    cnt = np.zeros(len(q), np.int32)
    lib().orc_knn_kdtree(_ptr(pts), len(pts), C.c_double(domain[0]), C.c_double(domain[1]), _ptr(q), C.c_longlong(len(q)), k, _ptr(idx), _ptr(d2),
                         _ptr(cnt))
    return idx, d2, cnt


def tree_propagate(S, D, xy, parent, root, divisor, mask, cost, col):
    H, W = (S if S is not None else D).shape
    xy = _c(xy, np.float32).reshape(-1, 2)
    parent = _c(parent, np.int32)
    cost = _c(cost, np.float32).copy()
    col = _c(col, np.uint8).copy()
    r = lib().orc_tree_propagate(_ptr(S), _ptr(D), W, H, _ptr(xy), _ptr(parent), len(xy), int(root), divisor, mask, _ptr(cost), _ptr(col))
    return cost, col, r


def relax_fixpoint(S, D, xy, knn_idx, divisor, mask, cost, col, parent, max_sweeps=1000, free_only=False):
    H, W = (S if S is not None else D).shape
    xy = _c(xy, np.float32).reshape(-1, 2)
    knn_idx = _c(knn_idx, np.int32)
    cost = _c(cost, np.float32).copy()
    col = _c(col, np.uint8).copy()
    parent = _c(parent, np.int32).copy()
    s = lib().orc_relax_fixpoint(_ptr(S), _ptr(D), W, H, _ptr(xy), len(xy), _ptr(knn_idx), knn_idx.shape[1], divisor, mask,
                                 _ptr(cost), _ptr(col), _ptr(parent), max_sweeps, int(free_only))
    return cost, col, parent, s


def raster_circle(img, cx, cy, radius, value=127):
    lib().orc_raster_circle(_ptr(img), img.shape[0], img.shape[1], int(cx), int(cy), int(radius), int(value))


def raster_fillpoly(img, pts, value=127):
    pts = _c(pts, np.int32).reshape(-1, 2)
    lib().orc_raster_fillpoly(_ptr(img), img.shape[0], img.shape[1], _ptr(pts), len(pts), int(value))


# ------------------------------------------------------------------ planners -------------------
class _TreePlanner:
    """Common wrapper over the orc_<name>_* / ref_<name>_* handle functions."""

    def __init__(self, L, prefix, handle, shape):
        self.L, self.pfx, self.h, self.shape = L, prefix, C.c_void_p(handle), shape

    def _f(self, name):
        return getattr(self.L, f"{self.pfx}_{name}")

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def set_agent(self, x, y):
        return self._f("set_agent")(self.h, C.c_float(x), C.c_float(y))

    def set_goal(self, x, y):
        return self._f("set_goal")(self.h, C.c_float(x), C.c_float(y))

    def set_dobs(self, plane):
        plane = None if plane is None else _c(plane, np.uint8)
        self._f("set_dobs")(self.h, _ptr(plane))

    def update_tree(self):
        self._f("update_tree")(self.h)

    def num_nodes(self):
        return self._f("num_nodes")(self.h)

    def dump(self):
        n = self.num_nodes()
        xy = np.zeros((n, 2), np.float32)
        par = np.zeros(n, np.int32)
        cost = np.zeros(n, np.float32)
        col = np.zeros(n, np.uint8)
        self._f("dump")(self.h, _ptr(xy), _ptr(par), _ptr(cost), _ptr(col))
        return xy, par, cost, col


class OrcRTRRT(_TreePlanner):
    def __init__(self, S, divisor=5):
        S = _c(S, np.uint8)
        super().__init__(lib(), "orc_rtrrt", lib().orc_rtrrt_create(_ptr(S), S.shape[0], S.shape[1], divisor), S.shape)

    def set_clock(self, dt_ns, n0=0):
        self.L.orc_rtrrt_set_clock(self.h, C.c_longlong(dt_ns), C.c_longlong(n0))

    def knn_mode(self, mode=-1):
        """mode 1: neighbour queries from a kd-tree rebuilt after every insert (CGAL's behaviour: the reference-shaped CPU cost),
        0: the grid (default); -1: just read.  Returns the number of kd-tree builds so far.  Answers are the same in both modes."""
        return int(self.L.orc_rtrrt_knn_mode(self.h, int(mode)))

    def presample(self, r, n_max):
        r = _c(r, np.int32)
        pts = np.zeros((n_max, 2), np.float32)
        offs = np.zeros(n_max + 1, np.int64)
        n = self.L.orc_rtrrt_presample(self.h, _ptr(r), C.c_longlong(len(r)), n_max, _ptr(pts), _ptr(offs))
        return pts[:n], offs[: n + 1]

    def expand_rewiring(self, ms, pts):
        pts = _c(pts, np.float32).reshape(-1, 2)
        return self.L.orc_rtrrt_expand_rewiring(self.h, C.c_long(ms), _ptr(pts), len(pts))

    def root(self):
        xy = np.zeros(2, np.float32)
        self.L.orc_rtrrt_root(self.h, _ptr(xy))
        return xy

    def queue_sizes(self):
        a, b = C.c_int(), C.c_int()
        self.L.orc_rtrrt_queue_sizes(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def counters(self):
        out = np.zeros(5, np.int64)
        self.L.orc_rtrrt_counters(self.h, _ptr(out))
        return dict(zip(["iters", "pops_r", "pops_s", "edge_checks", "knn"], out.tolist()))

    def get_path(self, cap=2048):
        xy = np.zeros((cap, 2), np.float32)
        col, cost, cb = C.c_int(), C.c_float(), C.c_float()
        n = self.L.orc_rtrrt_get_path(self.h, _ptr(xy), cap, C.byref(col), C.byref(cost), C.byref(cb))
        return xy[: min(n, cap)], bool(col.value), cost.value, cb.value


class RefRTRRT(_TreePlanner):
    """The reference's own RTRRTplanner (oracle/_ref)."""

    def __init__(self, S):
        L = ref_lib("rtrrt")
        if L is None:
            raise FileNotFoundError("oracle/_ref/libref_rtrrt.so not built")
        S = _c(S, np.uint8)
        super().__init__(L, "ref_rtrrt", L.ref_rtrrt_create(_ptr(S), S.shape[0], S.shape[1]), S.shape)
        self._rand = None

    def set_clock(self, dt_ns, n0=0):
        self.L.ref_set_clock(C.c_longlong(dt_ns), C.c_longlong(n0))

    def set_rand(self, r):
        self._rand = _c(r, np.int32)
        self.L.ref_set_rand(_ptr(self._rand), C.c_longlong(len(self._rand)))

    def rand_pos(self):
        return int(self.L.ref_rand_pos())

    def expand_rewiring(self, ms):
        self.L.ref_rtrrt_expand_rewiring(self.h, C.c_long(ms))

    def crash(self, xyxy):
        xyxy = _c(xyxy, np.float32).reshape(-1, 4)
        out = np.zeros(len(xyxy), np.uint8)
        self.L.ref_rtrrt_crash(self.h, _ptr(xyxy), C.c_longlong(len(xyxy)), _ptr(out))
        return out

    def dist(self, xyxy):
        xyxy = _c(xyxy, np.float32).reshape(-1, 4)
        out = np.zeros(len(xyxy), np.float32)
        self.L.ref_rtrrt_dist(_ptr(xyxy), C.c_longlong(len(xyxy)), _ptr(out))
        return out

    def add_obstacle(self, num, typ, x, y, angle, radius, outline):
        o = _c(outline, np.float32).reshape(-1, 2)
        self.L.ref_rtrrt_add_obstacle(self.h, num, typ, C.c_float(x), C.c_float(y), C.c_float(angle), C.c_float(radius), _ptr(o), len(o))

    def draw_dobs(self):
        self.L.ref_rtrrt_draw_dobs(self.h)

    def get_dobs(self):
        out = np.zeros(self.shape, np.uint8)
        self.L.ref_rtrrt_get_dobs(self.h, _ptr(out))
        return out

    def root(self):
        xy = np.zeros(2, np.float32)
        self.L.ref_rtrrt_root(self.h, _ptr(xy))
        return xy

    def queue_sizes(self):
        a, b = C.c_int(), C.c_int()
        self.L.ref_rtrrt_queue_sizes(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def get_path(self, cap=2048):
        xy = np.zeros((cap, 2), np.float32)
        col, cost, cb = C.c_int(), C.c_float(), C.c_float()
        n = self.L.ref_rtrrt_get_path(self.h, _ptr(xy), cap, C.byref(col), C.byref(cost), C.byref(cb))
        return xy[: min(n, cap)], bool(col.value), cost.value, cb.value


# ------------------------------------------------------------------ fixtures -------------------
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def load_map(name):
    """Obstacle byte plane (0/255) of a bundled map from the committed bit-plane fixture."""
    z = np.load(os.path.join(GOLDEN, "maps", f"{name}.npz"))
    h, w = (int(v) for v in z["shape"])
    bits = np.unpackbits(z["bits"], axis=1, bitorder="little")[:, :w]
    return np.ascontiguousarray(bits.astype(np.uint8) * 255)


def synthetic_map(size=4096, seed=42, n_rect=400, n_disc=200):
    """C4 map: seeded union of axis-aligned rectangles 16-256 px and discs r 8-96 px."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = np.zeros((size, size), np.uint8)
    for _ in range(n_rect):
        w, h = rng.integers(16, 257, 2)
        x, y = rng.integers(0, size - 16, 2)
        m[y : y + h, x : x + w] = 255
    yy, xx = np.ogrid[:size, :size]
    for _ in range(n_disc):
        r = int(rng.integers(8, 97))
        x, y = rng.integers(0, size, 2)
        y0, y1, x0, x1 = max(0, y - r), min(size, y + r + 1), max(0, x - r), min(size, x + r + 1)
        sub = (yy[y0:y1] - y) ** 2 + (xx[:, x0:x1] - x) ** 2 <= r * r
        m[y0:y1, x0:x1][sub] = 255
    return m


class OrcRTFMT(_TreePlanner):
    def __init__(self, S):
        S = _c(S, np.uint8)
        super().__init__(lib(), "orc_rtfmt", lib().orc_rtfmt_create(_ptr(S), S.shape[0], S.shape[1]), S.shape)

    def start(self, r, sample_n):
        r = _c(r, np.int32)
        return int(self.L.orc_rtfmt_start(self.h, _ptr(r), C.c_longlong(len(r)), sample_n))

    def expand(self, iters, with_rewire=True):
        self.L.orc_rtfmt_expand(self.h, iters, int(with_rewire))

    def dump(self):
        n = self.num_nodes()
        xy = np.zeros((n, 2), np.float32)
        par = np.zeros(n, np.int32)
        cost = np.zeros(n, np.float32)
        col = np.zeros(n, np.uint8)
        st = np.zeros(n, np.uint8)
        self._f("dump")(self.h, _ptr(xy), _ptr(par), _ptr(cost), _ptr(col), _ptr(st))
        return xy, par, cost, col, st

    def z(self):
        xy = np.zeros(2, np.float32)
        self._f("z")(self.h, _ptr(xy))
        return xy

    def counters(self):
        out = np.zeros(3, np.int64)
        self.L.orc_rtfmt_counters(self.h, _ptr(out))
        return dict(zip(["expand", "edge_checks", "knn"], out.tolist()))

    def get_path(self, cap=2048):
        xy = np.zeros((cap, 2), np.float32)
        n = self._f("get_path")(self.h, _ptr(xy), cap)
        return xy[: min(n, cap)]


class RefRTFMT(OrcRTFMT):
    """The reference's own RTFMTplanner (oracle/_ref); Start() draws 2000 samples per call."""

    def __init__(self, S):
        L = ref_lib("rtfmt")
        if L is None:
            raise FileNotFoundError("oracle/_ref/libref_rtfmt.so not built")
        S = _c(S, np.uint8)
        _TreePlanner.__init__(self, L, "ref_rtfmt", L.ref_rtfmt_create(_ptr(S), S.shape[0], S.shape[1]), S.shape)

    def start(self, r, sample_n):
        assert sample_n % 2000 == 0
        self._rand = _c(r, np.int32)
        self.L.ref_set_rand(_ptr(self._rand), C.c_longlong(len(self._rand)))
        for _ in range(sample_n // 2000):
            self.L.ref_rtfmt_start(self.h)
        return int(self.L.ref_rand_pos())

    def expand(self, iters, with_rewire=True):
        self.L.ref_rtfmt_expand(self.h, iters, int(with_rewire))

    def crash(self, xyxy, ndobs=False):
        xyxy = _c(xyxy, np.float32).reshape(-1, 4)
        out = np.zeros(len(xyxy), np.uint8)
        self.L.ref_rtfmt_crash(self.h, _ptr(xyxy), C.c_longlong(len(xyxy)), int(ndobs), _ptr(out))
        return out

    def counters(self):
        return {}


class OrcAMRRT(_TreePlanner):
    def __init__(self, S):
        S = _c(S, np.uint8)
        super().__init__(lib(), "orc_amrrt", lib().orc_amrrt_create(_ptr(S), S.shape[0], S.shape[1]), S.shape)

    def set_clock(self, dt_ns, n0=0):
        self.L.orc_amrrt_set_clock(self.h, C.c_longlong(dt_ns), C.c_longlong(n0))

    def set_rand(self, r):
        r = _c(r, np.int32)
        self.L.orc_amrrt_set_rand(self.h, _ptr(r), C.c_longlong(len(r)))

    def rand_pos(self):
        return int(self.L.orc_amrrt_rand_pos(self.h))

    def set_diffusion(self, cells_ij, matrix, ratio):
        cells = _c(cells_ij, np.int32).reshape(-1, 2)
        m = _c(matrix, np.float32).reshape(len(cells), -1)
        self._f("set_diffusion")(self.h, _ptr(cells), len(cells), _ptr(m), m.shape[1], C.c_float(ratio))

    def diffusivity(self, a, b):
        f = self._f("diffusivity")
        f.restype = C.c_float
        return float(f(self.h, C.c_float(a[0]), C.c_float(a[1]), C.c_float(b[0]), C.c_float(b[1])))

    def expand(self):
        self._f("expand")(self.h)

    def set_outpath(self, nonempty):
        self._f("set_outpath")(self.h, int(bool(nonempty)))

    def c_base(self):
        f = self._f("c_base")
        f.restype = C.c_float
        return float(f(self.h))

    def root(self):
        xy = np.zeros(2, np.float32)
        self._f("root")(self.h, _ptr(xy))
        return xy

    def rewire_root(self):
        self._f("rewire_root")(self.h)

    def main_task(self):
        self._f("main_task")(self.h)

    def get_path(self, cap=2048):
        xy = np.zeros((cap, 2), np.float32)
        col, cost = C.c_int(), C.c_float()
        n = self._f("get_path")(self.h, _ptr(xy), cap, C.byref(col), C.byref(cost))
        return xy[: min(n, cap)], bool(col.value), cost.value


class RefAMRRT(OrcAMRRT):
    """The reference's own AMRRTplanner (oracle/_ref)."""

    def __init__(self, S):
        L = ref_lib("amrrt")
        if L is None:
            raise FileNotFoundError("oracle/_ref/libref_amrrt.so not built")
        S = _c(S, np.uint8)
        _TreePlanner.__init__(self, L, "ref_amrrt", L.ref_amrrt_create(_ptr(S), S.shape[0], S.shape[1]), S.shape)

    def set_clock(self, dt_ns, n0=0):
        self.L.ref_set_clock(C.c_longlong(dt_ns), C.c_longlong(n0))

    def set_rand(self, r):
        self._rand = _c(r, np.int32)
        self.L.ref_set_rand(_ptr(self._rand), C.c_longlong(len(self._rand)))

    def rand_pos(self):
        return int(self.L.ref_rand_pos())

    def crash(self, xyxy):
        xyxy = _c(xyxy, np.float32).reshape(-1, 4)
        out = np.zeros(len(xyxy), np.uint8)
        self.L.ref_amrrt_crash(self.h, _ptr(xyxy), C.c_longlong(len(xyxy)), _ptr(out))
        return out

    def crash_mt(self, xyxy, nthreads):
        xyxy = _c(xyxy, np.float32).reshape(-1, 4)
        out = np.zeros(len(xyxy), np.uint8)
        self.L.ref_amrrt_crash_mt(self.h, _ptr(xyxy), C.c_longlong(len(xyxy)), _ptr(out), int(nthreads))
        return out

    def crash_start(self, xyxy):
        xyxy = _c(xyxy, np.float32).reshape(-1, 4)
        out = np.zeros(len(xyxy), np.uint8)
        ps = np.zeros((len(xyxy), 2), np.float32)
        self.L.ref_amrrt_crash_start(self.h, _ptr(xyxy), C.c_longlong(len(xyxy)), _ptr(out), _ptr(ps))
        return out, ps


# ------------------------------------------------------------------ CDRT-RRT* --------------------
def load_cdrt_fixture(name):
    """Decomposition graph + id planes of a bundled map (tools/make_cdrt_fixtures.py, reference StartCut + cv2)."""
    z = np.load(os.path.join(GOLDEN, "cdrt", f"{name}.npz"))
    H, W = (int(v) for v in z["shape"])
    graphs = []
    for g in range(int(z["n_groups"])):
        graphs.append({k: np.ascontiguousarray(z[f"g{g}_{k}"]) for k in
                       ["cut_se", "cut_core", "cp_off", "cp", "cc_off", "cc", "pv_off", "pv", "poly_core", "pc_off", "pc", "pp_off", "pp"]})
    return {"name": name, "W": W, "H": H, "graphs": graphs, "amap": np.ascontiguousarray(z["amap"]),
            "imap": np.ascontiguousarray(z["imap"]), "resolution": float(z["resolution"])}


class CdrtCpu:
    """One CDRT-RRT* planner on the CPU: impl='oracle' (oracle/cdrt_oracle.hpp) or impl='ref' (the reference's
    CDRT.cpp compiled in place, oracle/_ref/libref_cdrt.so).  Same calls on both, so tests run them side by side.
    The reference keeps two function-local static timers (CDRT.cpp:929-930) and the harness one global clock and
    rand stream: each 'ref' instance loads a private copy of the shared object so that every planner starts fresh."""

    def __init__(self, fx, impl="oracle"):
        self.impl, self.fx = impl, fx
        self._tmp = None
        if impl == "oracle":
            self.L, self.pre = lib(), "orc_cdrt_"
        else:
            import shutil
            import tempfile

            src = os.path.join(HERE, "_ref", "libref_cdrt.so")
            if not os.path.exists(src):
                raise FileNotFoundError(src)
            fd, self._tmp = tempfile.mkstemp(suffix=".so", prefix="libref_cdrt_")
            os.close(fd)
            shutil.copyfile(src, self._tmp)
            self.L, self.pre = C.CDLL(self._tmp), "ref_cdrt_"
            self.L.ref_rand_pos.restype = C.c_longlong
            self.L.ref_clock_reads.restype = C.c_longlong
        self.f("create").restype = C.c_void_p
        self.h = C.c_void_p(self.f("create")(fx["W"], fx["H"], C.c_double(0.1), C.c_double(0.5)))
        for g in fx["graphs"]:
            ncut, npoly = len(g["cut_se"]), len(g["poly_core"])
            self.f("set_graph")(self.h, ncut, _ptr(g["cut_se"]), _ptr(g["cut_core"]), _ptr(g["cp_off"]), _ptr(g["cp"]), _ptr(g["cc_off"]),
                                _ptr(g["cc"]), npoly, _ptr(g["pv_off"]), _ptr(g["pv"]), _ptr(g["poly_core"]), _ptr(g["pc_off"]), _ptr(g["pc"]),
                                _ptr(g["pp_off"]), _ptr(g["pp"]))
        self.f("set_planes")(self.h, _ptr(fx["amap"]), _ptr(fx["imap"]))
        self._rand = None

    def f(self, name):
        return getattr(self.L, self.pre + name)

    def close(self):
        if self.h:
            self.f("destroy")(self.h)
            self.h = None
        if self._tmp:
            os.unlink(self._tmp)
            self._tmp = None

    def set_params(self, r_th=10.0, p_th=0.85, nearK=20, RewireN=100):
        self.f("set_params")(self.h, C.c_double(r_th), C.c_double(p_th), nearK, RewireN)

    def set_clock(self, dt_ns, n0=0):
        if self.impl == "oracle":
            self.L.orc_cdrt_set_clock(self.h, C.c_longlong(dt_ns), C.c_longlong(n0))
        else:
            self.L.ref_set_clock(C.c_longlong(dt_ns), C.c_longlong(n0))

    def clock_reads(self):
        if self.impl == "oracle":
            self.L.orc_cdrt_clock_reads.restype = C.c_longlong
            return self.L.orc_cdrt_clock_reads(self.h)
        return self.L.ref_clock_reads()

    def set_rand(self, r):
        self._rand = _c(r, np.int32)
        if self.impl == "oracle":
            self.L.orc_cdrt_set_rand(self.h, _ptr(self._rand), C.c_longlong(len(self._rand)))
        else:
            self.L.ref_set_rand(_ptr(self._rand), C.c_longlong(len(self._rand)))

    def rand_pos(self):
        if self.impl == "oracle":
            self.L.orc_cdrt_rand_pos.restype = C.c_longlong
            return self.L.orc_cdrt_rand_pos(self.h)
        return self.L.ref_rand_pos()

    def set_agent(self, x, y):
        self.f("set_agent")(self.h, C.c_double(x), C.c_double(y))

    def set_goal(self, x, y):
        self.f("set_goal")(self.h, C.c_double(x), C.c_double(y))

    def add_obstacle(self, num, typ, x, y, angle, radius, outline, check_len=2.0):
        o = _c(outline, np.float64).reshape(-1, 2)
        self.f("add_obstacle")(self.h, num, typ, C.c_double(x), C.c_double(y), C.c_double(angle), C.c_double(radius), _ptr(o), len(o),
                               C.c_double(check_len))

    def move_obstacle(self, num, x, y, angle=0.0):
        self.f("move_obstacle")(self.h, num, C.c_double(x), C.c_double(y), C.c_double(angle))

    def check_points(self, num, cap=8192):
        out = np.zeros((cap, 2), np.float64)
        n = self.f("check_points")(self.h, num, _ptr(out), cap)
        return out[:n]

    def start(self, resolution=None):
        self.f("start")(self.h, C.c_double(self.fx["resolution"] if resolution is None else resolution))

    def task(self):
        self.f("task")(self.h)

    def draw_dobs(self):
        out = np.zeros((self.fx["H"], self.fx["W"]), np.uint8)
        self.f("draw_dobs")(self.h, _ptr(out))
        return out

    def set_dobs(self, plane):
        self.f("set_dobs")(self.h, _ptr(None if plane is None else _c(plane, np.uint8)))

    def crash_dobs_island(self):
        self.f("crash_dobs_island")(self.h)

    def update_tree_cost(self):
        self.f("update_tree_cost")(self.h)

    def crash(self, xyxy, inv_s, inv_e):
        xyxy = _c(xyxy, np.float64).reshape(-1, 4)
        inv_s, inv_e = _c(inv_s, np.int32), _c(inv_e, np.int32)
        out = np.zeros(len(xyxy), np.uint8)
        self.f("crash")(self.h, _ptr(xyxy), _ptr(inv_s), _ptr(inv_e), C.c_longlong(len(xyxy)), _ptr(out))
        return out

    def neighbor(self, x, y, invM, cap=4096):
        out = np.zeros((cap, 2), np.float64)
        n = self.f("neighbor")(self.h, C.c_double(x), C.c_double(y), C.c_int32(int(np.array(invM & 0xFFFFFFFF, np.uint32).view(np.int32))),
                               _ptr(out), cap)
        return out[:n]

    def dump(self):
        """Nodes in Node_Map order (sorted by (x, y)): xy, parent xy, cost, collision, invM, #children."""
        n = self.f("num_nodes")(self.h)
        xy, par = np.zeros((n, 2)), np.zeros((n, 2))
        cost, col, inv, ns = np.zeros(n), np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32)
        m = self.f("dump")(self.h, _ptr(xy), _ptr(par), _ptr(cost), _ptr(col), _ptr(inv), _ptr(ns))
        assert m == n
        return xy, par, cost, col, inv, ns

    def state(self):
        d, s = np.zeros(4), np.zeros(8, np.int32)
        self.f("state")(self.h, _ptr(d), _ptr(s))
        return dict(agent=d[:2].copy(), agent_tag=d[2:].copy(), apl=int(s[0]), island=int(s[1]), queue=int(s[2]), larger=int(s[3]),
                    smaller=int(s[4]), samples_left=int(s[5]), n_trees=int(s[6]), goal_set=int(s[7]))

    def tree_islands(self, cap=4096):
        a, b = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        n = self.f("tree_islands")(self.h, _ptr(a), _ptr(b), cap)
        return a[:n].copy(), b[:n].copy()

    def samples(self, cap=1 << 16):
        xy, cut = np.zeros((cap, 2)), np.zeros(cap, np.int32)
        n = self.f("samples")(self.h, _ptr(xy), _ptr(cut), cap)
        return xy[:n].copy(), cut[:n].copy()

    def get_path(self, cap=2048):
        xy = np.zeros((cap, 2))
        col, cost = C.c_int(), C.c_double()
        n = self.f("get_path")(self.h, _ptr(xy), cap, C.byref(col), C.byref(cost))
        return xy[: min(n, cap)].copy(), bool(col.value), cost.value

    def counters(self):
        out = np.zeros(6, np.int64)
        if self.impl == "oracle":
            self.L.orc_cdrt_counters(self.h, _ptr(out))
        return dict(zip(["pops", "rewired", "sampled", "edge_checks", "knn", "pushes"], out.tolist()))

    def error(self):
        return self.L.orc_cdrt_error(self.h) if self.impl == "oracle" else 0


# ------------------------------------------------------------------ glibc transcendentals ---------
def glibc_host_fma():
    """1 when this host's libm runs its fused-multiply-add build of sinf / cosf."""
    return int(lib().orc_glibc_host_fma())


def glibc_checksum(impl, fn, first, count, nthreads=1):
    """Checksum of fn (0 sinf, 1 cosf, 2 atanf) over the binary32 bit patterns [first, first + count): impl 0 = the host's libm,
    1 / 2 = rtpp_b200/csrc/glibc_math.cuh compiled for the host without / with fused operations."""
    L = lib()
    L.orc_glibc_checksum.restype = C.c_ulonglong
    return int(L.orc_glibc_checksum(int(impl), int(fn), C.c_uint32(first & 0xFFFFFFFF), C.c_ulonglong(count), int(nthreads)))


def glibc_atan2f(impl, yx):
    yx = _c(yx, np.float32).reshape(-1, 2)
    out = np.zeros(len(yx), np.float32)
    lib().orc_glibc_atan2f(int(impl), _ptr(yx), C.c_longlong(len(yx)), _ptr(out))
    return out
