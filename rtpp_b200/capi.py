"""ctypes mirror of include/rtpp_gpu.h (one method per entry point, same names minus `rtpp_`).

Host-buffer methods take/return numpy arrays; *_dev methods take raw device pointers (ints), e.g.
torch.Tensor.data_ptr().  Fails loudly when librtpp_gpu.so is missing or no GPU is present.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PLANE_STATIC, PLANE_DYNAMIC = 1, 2


class RtppError(RuntimeError):
    pass


def lib_path():
    return os.path.join(HERE, "librtpp_gpu.so")


def build(verbose=False):
    """Compile every CUDA source for sm_100a into rtpp_b200/librtpp_gpu.so (nvcc cross-compiles)."""
    out = subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode:
        raise RtppError("building librtpp_gpu.so failed")


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path()):
            raise RtppError(f"{lib_path()} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        _lib = C.CDLL(lib_path())
        _lib.rtpp_status_string.restype = C.c_char_p
        _lib.rtpp_last_error.restype = C.c_char_p
        _lib.rtpp_kernel_launches.restype = C.c_longlong
    return _lib


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Ctx:
    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.rtpp_create(int(device), C.byref(h))
        if rc:
            raise RtppError(f"rtpp_create(device={device}) failed: {self.L.rtpp_status_string(rc).decode()} "
                            "(a CUDA device is required; there is no CPU fallback)")
        self.h = h

    def close(self):
        if self.h:
            self.L.rtpp_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise RtppError(f"{self.L.rtpp_status_string(rc).decode()}: {self.L.rtpp_last_error(self.h).decode()}")

    def call(self, name, *args):
        self._ck(getattr(self.L, "rtpp_" + name)(self.h, *args))

    # ---- lifecycle
    def set_stream(self, stream_ptr):
        self.call("set_stream", C.c_void_p(stream_ptr))

    def sync(self):
        self.call("sync")

    def kernel_launches(self):
        return int(self.L.rtpp_kernel_launches(self.h))

    def sm_count(self):
        n = C.c_int()
        self.call("device_sm_count", C.byref(n))
        return n.value

    def measure_peaks(self, l2_bytes=8 << 20, reps=5):
        """Same-run roofline denominators (rtpp_measure_peaks): L2 sector-gather / L2 stream / HBM copy GB/s, FP32 TFLOP/s."""
        out = (C.c_double * 4)()
        self.call("measure_peaks", C.c_size_t(l2_bytes), C.c_int(reps), out)
        return {"l2_sector_gather_gbs": out[0], "l2_stream_gbs": out[1], "fp32_ffma_tflops": out[2], "hbm_copy_gbs": out[3]}

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self.call("dev_alloc", C.c_size_t(nbytes), C.byref(p))
        return p.value

    def dev_free(self, ptr):
        self.call("dev_free", C.c_void_p(ptr))

    def h2d(self, dst, src_np):
        self.call("h2d", C.c_void_p(dst), _p(src_np), C.c_size_t(src_np.nbytes))

    def d2h(self, dst_np, src):
        self.call("d2h", _p(dst_np), C.c_void_p(src), C.c_size_t(dst_np.nbytes))

    # ---- maps
    def map_set_plane(self, plane, rowmajor):
        a = _c(rowmajor, np.uint8)
        self.call("map_set_plane", int(plane), _p(a), a.shape[1], a.shape[0])

    def map_clear_plane(self, plane):
        self.call("map_clear_plane", int(plane))

    def map_get_plane(self, mask):
        w, h = C.c_int(), C.c_int()
        self.call("map_shape", C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        self.call("map_get_plane", int(mask), _p(out))
        return out

    def map_draw_dynamic(self, circles=(), polys=()):
        """circles: iterable of (cx, cy, r); polys: iterable of (n_i, 2) integer vertex arrays."""
        c = _c(np.asarray(list(circles), np.int32).reshape(-1, 3), np.int32)
        polys = [np.asarray(p, np.int32).reshape(-1, 2) for p in polys]
        pxy = _c(np.concatenate(polys) if polys else np.zeros((0, 2), np.int32), np.int32)
        sizes = _c(np.array([len(p) for p in polys], np.int32), np.int32)
        self.call("map_draw_dynamic", _p(c) if len(c) else None, len(c), _p(pxy) if len(polys) else None,
                  _p(sizes) if len(polys) else None, len(polys))

    # ---- edges
    @staticmethod
    def unpack_bits(bits, n):
        return np.unpackbits(bits.view(np.uint8), bitorder="little")[:n]

    def edges_check(self, xyxy, divisor, mask):
        e = _c(xyxy, np.float32).reshape(-1, 4)
        bits = np.zeros((len(e) + 31) // 32, np.uint32)
        self.call("edges_check", _p(e), C.c_size_t(len(e)), int(divisor), int(mask), _p(bits))
        return self.unpack_bits(bits, len(e))

    def edges_check_f64(self, xyxy, divisor, mask):
        e = _c(xyxy, np.float64).reshape(-1, 4)
        bits = np.zeros((len(e) + 31) // 32, np.uint32)
        self.call("edges_check_f64", _p(e), C.c_size_t(len(e)), int(divisor), int(mask), _p(bits))
        return self.unpack_bits(bits, len(e))

    def edges_first_hit(self, xyxy):
        e = _c(xyxy, np.float32).reshape(-1, 4)
        bits = np.zeros((len(e) + 31) // 32, np.uint32)
        ps = np.zeros((len(e), 2), np.float32)
        self.call("edges_first_hit", _p(e), C.c_size_t(len(e)), _p(bits), _p(ps))
        return self.unpack_bits(bits, len(e)), ps

    def edges_check_dev(self, d_xyxy, n, divisor, mask, d_bits):
        self.call("edges_check_dev", C.c_void_p(d_xyxy), C.c_size_t(n), int(divisor), int(mask), C.c_void_p(d_bits))

    def dist(self, xyxy):
        e = _c(xyxy, np.float32).reshape(-1, 4)
        out = np.zeros(len(e), np.float32)
        self.call("dist", _p(e), C.c_size_t(len(e)), _p(out))
        return out

    # ---- nodes + k-NN
    def nodes_reset(self, cell_log2, capacity, domain_w, domain_h):
        self.call("nodes_reset", int(cell_log2), C.c_size_t(capacity), int(domain_w), int(domain_h))

    def nodes_append(self, xy):
        a = _c(xy, np.float32).reshape(-1, 2)
        first = C.c_int32()
        self.call("nodes_append", _p(a), C.c_size_t(len(a)), C.byref(first))
        return first.value

    def nodes_count(self):
        n = C.c_size_t()
        self.call("nodes_count", C.byref(n))
        return n.value

    def nodes_build(self):
        self.call("nodes_build")

    def nodes_xy_dev(self):
        p = C.c_void_p()
        self.call("nodes_xy_dev", C.byref(p))
        return p.value

    def knn(self, qxy, k):
        q = _c(qxy, np.float64).reshape(-1, 2)
        idx = np.zeros((len(q), k), np.int32)
        d2 = np.zeros((len(q), k), np.float64)
        cnt = np.zeros(len(q), np.int32)
        self.call("knn", _p(q), C.c_size_t(len(q)), int(k), _p(idx), _p(d2), _p(cnt))
        return idx, d2, cnt

    def knn_dev(self, d_q, nq, k, d_idx, d_d2=0, d_count=0):
        self.call("knn_dev", C.c_void_p(d_q), C.c_size_t(nq), int(k), C.c_void_p(d_idx), C.c_void_p(d_d2), C.c_void_p(d_count))

    def knn_nodes_dev(self, k, d_idx, d_d2=0):
        self.call("knn_nodes_dev", int(k), C.c_void_p(d_idx), C.c_void_p(d_d2))

    def knn_table_build(self, k):
        self.call("knn_table_build", int(k))

    def tick_edges_check(self, dynamic_plane, divisor, mask):
        """One tick through host buffers: dynamic byte plane in (None = unchanged), verdict bits out."""
        n = self.nodes_count()
        d = None if dynamic_plane is None else _c(dynamic_plane, np.uint8)
        ne = C.c_size_t()
        bits = np.zeros(((n - 1) + n * 32 + 31) // 32, np.uint32)
        self.call("tick_edges_check", _p(d), int(divisor), int(mask), _p(bits), C.byref(ne))
        return self.unpack_bits(bits, ne.value)

    def tick_plan_info(self, divisor=2, want_ids=True):
        """The slot order of the windowed tick (rtpp_tick_plan_info): dict(n_slots, usable, far_fraction, slot_id, slot_of_id)."""
        ns, p_id, p_of = C.c_size_t(), C.c_void_p(), C.c_void_p()
        usable, far = C.c_int(), C.c_double()
        self.call("tick_plan_info", int(divisor), C.byref(ns), C.byref(p_id), C.byref(p_of), C.byref(usable), C.byref(far))
        out = {"n_slots": ns.value, "usable": bool(usable.value), "far_fraction": far.value, "d_slot_id": p_id.value, "d_slot_of_id": p_of.value}
        if want_ids and ns.value:
            out["slot_id"] = np.zeros(ns.value, np.int32)
            self.d2h(out["slot_id"], p_id.value)
            out["slot_of_id"] = np.zeros(self.nodes_count(), np.int32)
            self.d2h(out["slot_of_id"], p_of.value)
            self.sync()
        return out

    def tick_check(self, divisor, mask):
        """One tick through host buffers in the windowed kernel's layout -> (tree_bits [n-1 verdicts], row_words [n_slots])."""
        n = self.nodes_count()
        info = self.tick_plan_info(divisor, want_ids=False)
        tree = np.zeros(max(1, (max(n - 1, 0) + 31) // 32), np.uint32)
        rows = np.zeros(max(1, info["n_slots"]), np.uint32)
        self.call("tick_check", int(divisor), int(mask), _p(tree), _p(rows))
        return self.unpack_bits(tree, max(n - 1, 0)), rows[: info["n_slots"]]

    def tick_check_dev(self, divisor, mask, d_tree_bits, d_row_bits):
        self.call("tick_check_dev", int(divisor), int(mask), C.c_void_p(d_tree_bits), C.c_void_p(d_row_bits))

    # ---- tree costs
    def tree_set(self, parent, cost, col):
        p, c, l = _c(parent, np.int32), _c(cost, np.float32), _c(col, np.uint8)
        self.call("tree_set", _p(p), _p(c), _p(l), C.c_size_t(len(p)))

    def tree_get(self, n=None):
        n = self.nodes_count() if n is None else n
        p, c, l = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        self.call("tree_get", _p(p), _p(c), _p(l), C.c_size_t(n))
        return p, c, l

    def tree_reroot(self, old_root, new_root):
        self.call("tree_reroot", int(old_root), int(new_root))

    def tree_propagate(self, root, divisor, mask, want_flags=False):
        n = self.nodes_count()
        flags = np.zeros(n, np.uint8) if want_flags else None
        reached = C.c_int()
        self.call("tree_propagate", int(root), int(divisor), int(mask), _p(flags), C.byref(reached))
        return reached.value, flags

    def tree_propagate_dev(self, root, divisor, mask):
        self.call("tree_propagate_dev", int(root), int(divisor), int(mask), None, None)

    def relax_dev(self, d_knn_idx, k, divisor, mask, free_only=False, max_sweeps=10000):
        sw = C.c_int()
        self.call("relax_dev", C.c_void_p(d_knn_idx), int(k), int(divisor), int(mask), int(free_only), int(max_sweeps), C.byref(sw))
        return sw.value


class Engine:
    """ctypes mirror of the rtpp_engine_* entry points (planner_kind 1 = RT-RRT*)."""

    def __init__(self, ctx, n_instances, node_capacity, cell_log2, domain_w, domain_h, max_samples, kind=1):
        self.ctx, self.L = ctx, ctx.L
        self.n, self.cap, self.max_samples = n_instances, node_capacity, max_samples
        h = C.c_void_p()
        ctx._ck(self.L.rtpp_engine_create(ctx.h, kind, n_instances, node_capacity, cell_log2, domain_w, domain_h, max_samples, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.rtpp_engine_destroy(self.h)
            self.h = None

    def _ck(self, rc):
        self.ctx._ck(rc)

    def set_map(self, slot, plane, rowmajor, shape=None):
        if rowmajor is None:
            self._ck(self.L.rtpp_engine_set_map(self.h, slot, plane, None, 0, 0))
        else:
            a = _c(rowmajor, np.uint8)
            self._ck(self.L.rtpp_engine_set_map(self.h, slot, plane, _p(a), a.shape[1], a.shape[0]))

    def init_instances(self, map_slot, agent_xy, goal_xy):
        s, a, g = _c(map_slot, np.int32), _c(agent_xy, np.float32).reshape(-1, 2), _c(goal_xy, np.float32).reshape(-1, 2)
        assert len(s) == len(a) == len(g) == self.n
        self._ck(self.L.rtpp_engine_init_instances(self.h, _p(s), _p(a), _p(g)))

    def set_inputs(self, agent_xy=None, goal_xy=None, samples=None, n_samples=None):
        a = None if agent_xy is None else _c(agent_xy, np.float32).reshape(self.n, 2)
        g = None if goal_xy is None else _c(goal_xy, np.float32).reshape(self.n, 2)
        sm = None if samples is None else _c(samples, np.float32).reshape(self.n, self.max_samples, 2)
        ns = None if n_samples is None else _c(n_samples, np.int32)
        self._ck(self.L.rtpp_engine_set_inputs(self.h, _p(a), _p(g), _p(sm), _p(ns)))

    def set_rand(self, rand=None, n_rand=None):
        """RT-RRT*: the rand() stream of the next tick(s); the engine draws its samples on the device.  None switches back."""
        if rand is None:
            self._ck(self.L.rtpp_engine_set_rand(self.h, None, None))
            return
        r = _c(rand, np.int32).reshape(self.n, self.max_samples)
        self._ck(self.L.rtpp_engine_set_rand(self.h, _p(r), _p(_c(n_rand, np.int32))))

    def tick(self, update_tree=True, run_ms=50, dt_ns=100000, get_path=False):
        self._ck(self.L.rtpp_engine_tick(self.h, int(update_tree), C.c_long(run_ms), C.c_longlong(dt_ns), int(get_path)))

    STAT = ["n_nodes", "root", "samples_used", "clock_reads", "iters", "pops_r", "pops_s", "qr_size", "qs_size", "status",
            "path_len", "path_col", "path_cost_bits", "c_base_bits", "rand_pos", "rand_exhausted"]

    # ---- AM-RRT* (kind 2)
    def set_diffusion(self, cells_ij, matrix, ratio):
        cells = _c(cells_ij, np.int32).reshape(-1, 2)
        m = _c(matrix, np.float32).reshape(len(cells), -1)
        self._ck(self.L.rtpp_engine_set_diffusion(self.h, _p(cells), len(cells), _p(m), m.shape[1], C.c_float(ratio)))

    def set_am_inputs(self, rand=None, n_rand=None, ellipse=None, outpath_nonempty=None):
        r = None if rand is None else _c(rand, np.int32).reshape(self.n, self.max_samples)
        nr = None if n_rand is None else _c(n_rand, np.int32)
        el = None if ellipse is None else _c(ellipse, np.float32).reshape(self.n, 6)
        op = None if outpath_nonempty is None else _c(outpath_nonempty, np.int32)
        self._ck(self.L.rtpp_engine_set_am_inputs(self.h, _p(r), _p(nr), _p(el), _p(op)))

    def stats(self):
        out = np.zeros((self.n, 16), np.int64)
        self._ck(self.L.rtpp_engine_read_stats(self.h, _p(out)))
        return out

    def stats_dict(self, inst=0):
        row = self.stats()[inst]
        d = dict(zip(self.STAT, row.tolist()))
        d["path_cost"] = float(np.array([d.pop("path_cost_bits")], np.uint32).view(np.float32)[0])
        d["c_base"] = float(np.array([d.pop("c_base_bits")], np.uint32).view(np.float32)[0])
        return d

    def read_tree(self, inst=0):
        n = C.c_int()
        self._ck(self.L.rtpp_engine_read_tree(self.h, inst, 0, None, None, None, None, C.byref(n)))
        n = n.value
        xy, par = np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
        cost, col = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        m = C.c_int()
        self._ck(self.L.rtpp_engine_read_tree(self.h, inst, n, _p(xy), _p(par), _p(cost), _p(col), C.byref(m)))
        return xy, par, cost, col

    def read_path(self, inst=0, cap=2048):
        ids = np.zeros(cap, np.int32)
        n = C.c_int()
        self._ck(self.L.rtpp_engine_read_path(self.h, inst, cap, _p(ids), C.byref(n)))
        return ids[: min(n.value, cap)]


class FmtEngine:
    """ctypes mirror of the rtpp_fmt_* entry points (RT-FMT engine)."""

    STAT = ["n_nodes", "root", "z", "n_unv", "n_open", "n_opennew", "n_toopen", "x_near", "n_b", "qo_size", "qr_size",
            "expand_calls", "status", "path_len", "path_ok"]

    def __init__(self, ctx, n_instances, node_capacity, cell_log2, domain_w, domain_h):
        self.ctx, self.L, self.n, self.cap = ctx, ctx.L, n_instances, node_capacity
        h = C.c_void_p()
        ctx._ck(self.L.rtpp_fmt_create(ctx.h, n_instances, node_capacity, cell_log2, domain_w, domain_h, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.rtpp_fmt_destroy(self.h)
            self.h = None

    def set_map(self, slot, plane, rowmajor):
        if rowmajor is None:
            self.ctx._ck(self.L.rtpp_fmt_set_map(self.h, slot, plane, None, 0, 0))
        else:
            a = _c(rowmajor, np.uint8)
            self.ctx._ck(self.L.rtpp_fmt_set_map(self.h, slot, plane, _p(a), a.shape[1], a.shape[0]))

    def start(self, map_slot, nodes_list, goal_xy):
        """nodes_list: per instance an (n_i, 2) float32 array, row 0 = the agent."""
        xy = np.zeros((self.n, self.cap, 2), np.float32)
        nn = np.zeros(self.n, np.int32)
        for i, a in enumerate(nodes_list):
            a = _c(a, np.float32).reshape(-1, 2)
            xy[i, : len(a)] = a
            nn[i] = len(a)
        s, g = _c(map_slot, np.int32), _c(goal_xy, np.float32).reshape(self.n, 2)
        self.ctx._ck(self.L.rtpp_fmt_start(self.h, _p(s), _p(nn), _p(xy), _p(g)))

    def set_inputs(self, agent_xy=None, goal_xy=None):
        a = None if agent_xy is None else _c(agent_xy, np.float32).reshape(self.n, 2)
        g = None if goal_xy is None else _c(goal_xy, np.float32).reshape(self.n, 2)
        self.ctx._ck(self.L.rtpp_fmt_set_inputs(self.h, _p(a), _p(g)))

    def tick(self, update_tree=True, iters=500, with_rewire=True, get_path=False):
        self.ctx._ck(self.L.rtpp_fmt_tick(self.h, int(update_tree), int(iters), 7 if with_rewire is True else (1 if with_rewire is False else int(with_rewire)), int(get_path)))

    def stats(self):
        out = np.zeros((self.n, 16), np.int64)
        self.ctx._ck(self.L.rtpp_fmt_read_stats(self.h, _p(out)))
        return out

    def stats_dict(self, inst=0):
        return dict(zip(self.STAT, self.stats()[inst].tolist()))

    def read_tree(self, inst=0):
        n = C.c_int()
        self.ctx._ck(self.L.rtpp_fmt_read_tree(self.h, inst, 0, None, None, None, None, C.byref(n)))
        n = n.value
        par, cost = np.zeros(n, np.int32), np.zeros(n, np.float32)
        col, st = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self.ctx._ck(self.L.rtpp_fmt_read_tree(self.h, inst, n, _p(par), _p(cost), _p(col), _p(st), C.byref(C.c_int())))
        return par, cost, col, st

    def read_path(self, inst=0, cap=2048):
        ids = np.zeros(cap, np.int32)
        n = C.c_int()
        self.ctx._ck(self.L.rtpp_fmt_read_path(self.h, inst, cap, _p(ids), C.byref(n)))
        return ids[: min(n.value, cap)]


# ---------------------------------------------------------------------------------------------
# C++ drop-in planner classes (rtpp_b200/planners, librtpp_planners.so) through their C harness (librtpp_planners_harness.so, test only)
# ---------------------------------------------------------------------------------------------
_plib = None


def planners_lib():
    global _plib
    if _plib is None:
        path = os.path.join(HERE, "librtpp_planners_harness.so")
        if not os.path.exists(path):
            raise RtppError(f"{path} is missing: run __graft_entry__.build()")
        lib()  # librtpp_gpu.so first (same directory, rpath $ORIGIN)
        _plib = C.CDLL(path)
        _plib.pl_rtrrt_create.restype = C.c_void_p
        _plib.pl_rtrrt_rand_pos.restype = C.c_longlong
        _plib.pl_rtrrt_clock_reads.restype = C.c_longlong
    return _plib


class PlannerRTRRT:
    """The C++ RTRRTplanner drop-in (same methods as the reference class) driven from Python."""

    def __init__(self, obstacle_plane, device=0):
        self.L = planners_lib()
        S = _c(obstacle_plane, np.uint8)
        self.shape = S.shape
        h = self.L.pl_rtrrt_create(_p(S), S.shape[0], S.shape[1], device)
        if not h:
            raise RtppError("RTRRTplanner could not be created (no CUDA device; there is no CPU fallback)")
        self.h = C.c_void_p(h)
        self._rand = None

    def close(self):
        if self.h:
            self.L.pl_rtrrt_destroy(self.h)
            self.h = None

    def set_clock(self, dt_ns):
        self.L.pl_rtrrt_set_clock(self.h, C.c_longlong(dt_ns))

    def set_rand(self, r):
        self._rand = _c(r, np.int32)
        self.L.pl_rtrrt_set_rand(self.h, _p(self._rand), C.c_longlong(len(self._rand)))

    def rand_pos(self):
        return int(self.L.pl_rtrrt_rand_pos(self.h))

    def set_agent(self, x, y):
        self.L.pl_rtrrt_set_agent(self.h, C.c_float(x), C.c_float(y))

    def set_goal(self, x, y):
        self.L.pl_rtrrt_set_goal(self.h, C.c_float(x), C.c_float(y))

    def set_dobs(self, plane):
        plane = None if plane is None else _c(plane, np.uint8)
        self.L.pl_rtrrt_set_dobs(self.h, _p(plane))

    def add_obstacle(self, num, typ, x, y, angle, radius, outline):
        o = _c(outline, np.float32).reshape(-1, 2)
        self.L.pl_rtrrt_add_obstacle(self.h, num, typ, C.c_float(x), C.c_float(y), C.c_float(angle), C.c_float(radius), _p(o), len(o))

    def draw_dobs(self):
        self.L.pl_rtrrt_draw_dobs(self.h)

    def get_dobs(self):
        out = np.zeros(self.shape, np.uint8)
        self.L.pl_rtrrt_get_dobs(self.h, _p(out))
        return out

    def update_tree(self):
        self.L.pl_rtrrt_update_tree(self.h)

    def expand_rewiring(self, ms):
        self.L.pl_rtrrt_expand_rewiring(self.h, C.c_long(ms))

    def main_task(self):
        self.L.pl_rtrrt_main_task(self.h)

    def crash(self, a, b):
        return bool(self.L.pl_rtrrt_crash(self.h, C.c_float(a[0]), C.c_float(a[1]), C.c_float(b[0]), C.c_float(b[1])))

    def num_nodes(self):
        return self.L.pl_rtrrt_num_nodes(self.h)

    def dump(self):
        n = self.num_nodes()
        xy, par = np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
        cost, col = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        self.L.pl_rtrrt_dump(self.h, _p(xy), _p(par), _p(cost), _p(col))
        return xy, par, cost, col

    def root(self):
        xy = np.zeros(2, np.float32)
        self.L.pl_rtrrt_root(self.h, _p(xy))
        return xy

    def queue_sizes(self):
        a, b = C.c_int(), C.c_int()
        self.L.pl_rtrrt_queue_sizes(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def get_path(self, cap=2048):
        xy = np.zeros((cap, 2), np.float32)
        col, cost = C.c_int(), C.c_float()
        n = self.L.pl_rtrrt_get_path(self.h, _p(xy), cap, C.byref(col), C.byref(cost))
        return xy[: min(n, cap)], bool(col.value), cost.value


def host_raster_circle(img, cx, cy, r, v=127):
    planners_lib().pl_raster_circle(_p(img), img.shape[0], img.shape[1], int(cx), int(cy), int(r), int(v))


def host_raster_fillpoly(img, pts, v=127):
    pts = _c(pts, np.int32).reshape(-1, 2)
    planners_lib().pl_raster_fillpoly(_p(img), img.shape[0], img.shape[1], _p(pts), len(pts), int(v))


class PlannerRTFMT:
    """The C++ RTFMTplanner drop-in driven from Python (same call sequence as the reference class)."""

    def __init__(self, obstacle_plane, device=0):
        self.L = planners_lib()
        self.L.pl_rtfmt_create.restype = C.c_void_p
        self.L.pl_rtfmt_rand_pos.restype = C.c_longlong
        S = _c(obstacle_plane, np.uint8)
        h = self.L.pl_rtfmt_create(_p(S), S.shape[0], S.shape[1], device)
        if not h:
            raise RtppError("RTFMTplanner could not be created (no CUDA device; there is no CPU fallback)")
        self.h = C.c_void_p(h)

    def close(self):
        if self.h:
            self.L.pl_rtfmt_destroy(self.h)
            self.h = None

    def set_rand(self, r):
        self._rand = _c(r, np.int32)
        self.L.pl_rtfmt_set_rand(self.h, _p(self._rand), C.c_longlong(len(self._rand)))

    def rand_pos(self):
        return int(self.L.pl_rtfmt_rand_pos(self.h))

    def set_agent(self, x, y):
        self.L.pl_rtfmt_set_agent(self.h, C.c_float(x), C.c_float(y))

    def set_goal(self, x, y):
        self.L.pl_rtfmt_set_goal(self.h, C.c_float(x), C.c_float(y))

    def set_dobs(self, plane):
        plane = None if plane is None else _c(plane, np.uint8)
        self.L.pl_rtfmt_set_dobs(self.h, _p(plane))

    def start(self, sample_n=2000):
        self.L.pl_rtfmt_start(self.h, int(sample_n))

    def update_tree(self):
        self.L.pl_rtfmt_update_tree(self.h)

    def expand(self, iters, with_rewire=True):
        self.L.pl_rtfmt_expand(self.h, int(iters), int(with_rewire))

    def num_nodes(self):
        return self.L.pl_rtfmt_num_nodes(self.h)

    def dump(self):
        n = self.num_nodes()
        xy, par = np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
        cost, col, st = np.zeros(n, np.float32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self.L.pl_rtfmt_dump(self.h, _p(xy), _p(par), _p(cost), _p(col), _p(st))
        return xy, par, cost, col, st

    def z(self):
        xy = np.zeros(2, np.float32)
        self.L.pl_rtfmt_z(self.h, _p(xy))
        return xy

    def get_path(self, cap=2048):
        xy = np.zeros((cap, 2), np.float32)
        n = self.L.pl_rtfmt_get_path(self.h, _p(xy), cap)
        return xy[: min(n, cap)]

    def set_sizes(self):
        out = np.zeros(6, np.int32)
        self.L.pl_rtfmt_set_sizes(self.h, _p(out))
        return dict(zip(["unv", "open", "opennew", "closed", "toopen", "node_map"], out.tolist()))


class PlannerAMRRT:
    """The C++ AMRRTplanner drop-in driven from Python."""

    def __init__(self, obstacle_plane, device=0):
        self.L = planners_lib()
        self.L.pl_amrrt_create.restype = C.c_void_p
        self.L.pl_amrrt_rand_pos.restype = C.c_longlong
        self.L.pl_amrrt_diffusivity.restype = C.c_float
        S = _c(obstacle_plane, np.uint8)
        h = self.L.pl_amrrt_create(_p(S), S.shape[0], S.shape[1], device)
        if not h:
            raise RtppError("AMRRTplanner could not be created (no CUDA device; there is no CPU fallback)")
        self.h = C.c_void_p(h)

    def close(self):
        if self.h:
            self.L.pl_amrrt_destroy(self.h)
            self.h = None

    def set_clock(self, dt_ns):
        self.L.pl_amrrt_set_clock(self.h, C.c_longlong(dt_ns))

    def set_rand(self, r):
        self._rand = _c(r, np.int32)
        self.L.pl_amrrt_set_rand(self.h, _p(self._rand), C.c_longlong(len(self._rand)))

    def rand_pos(self):
        return int(self.L.pl_amrrt_rand_pos(self.h))

    def set_agent(self, x, y):
        self.L.pl_amrrt_set_agent(self.h, C.c_float(x), C.c_float(y))

    def set_goal(self, x, y):
        self.L.pl_amrrt_set_goal(self.h, C.c_float(x), C.c_float(y))

    def set_dobs(self, plane):
        plane = None if plane is None else _c(plane, np.uint8)
        self.L.pl_amrrt_set_dobs(self.h, _p(plane))

    def set_diffusion(self, cells_ij, matrix, ratio):
        cells = _c(cells_ij, np.int32).reshape(-1, 2)
        m = _c(matrix, np.float32).reshape(len(cells), -1)
        self.L.pl_amrrt_set_diffusion(self.h, _p(cells), len(cells), _p(m), m.shape[1], C.c_float(ratio))

    def diffusivity(self, a, b):
        return float(self.L.pl_amrrt_diffusivity(self.h, C.c_float(a[0]), C.c_float(a[1]), C.c_float(b[0]), C.c_float(b[1])))

    def set_outpath(self, nonempty):
        self.L.pl_amrrt_set_outpath(self.h, int(bool(nonempty)))

    def main_task(self):
        self.L.pl_amrrt_main_task(self.h)

    def num_nodes(self):
        return self.L.pl_amrrt_num_nodes(self.h)

    def dump(self):
        n = self.num_nodes()
        xy, par = np.zeros((n, 2), np.float32), np.zeros(n, np.int32)
        cost, col = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        self.L.pl_amrrt_dump(self.h, _p(xy), _p(par), _p(cost), _p(col))
        return xy, par, cost, col

    def get_path(self, cap=2048):
        xy = np.zeros((cap, 2), np.float32)
        col, cost = C.c_int(), C.c_float()
        n = self.L.pl_amrrt_get_path(self.h, _p(xy), cap, C.byref(col), C.byref(cost))
        return xy[: min(n, cap)], bool(col.value), cost.value


class CdrtEngine:
    """ctypes mirror of the rtpp_cdrt_* entry points (CDRT-RRT* engine, one warp per BImap planner).
    `graph`: dict of the CSR arrays of one BIgraph (cut_se, cut_core, cp_off, cp, cc_off, cc, pv_off, pv, pc_off, pc);
    `amap`/`imap`: the two 16-bit id planes of DrawBImap; `apl` = the region (graph index) the agents live in."""

    STAT = ["n_ids", "agent_id", "island", "clock_reads", "rand_pos", "rand_exhausted", "queue", "larger", "smaller", "samples_left",
            "n_trees", "status", "pops", "rewired", "sampled", "pushes", "edge_checks", "path_len", "path_col", "path_cost_bits",
            "cyc_env_tree", "cyc_sets", "cyc_sampling", "cyc_queue"]

    def __init__(self, ctx, n_instances, graph, amap, imap, resolution, apl=0, node_capacity=8192, queue_capacity=1 << 20,
                 r_th=10.0, p_th=0.85, nearK=20, RewireN=100, rand_capacity=1 << 20):
        self.ctx, self.L, self.n = ctx, ctx.L, n_instances
        g = {k: np.ascontiguousarray(v) for k, v in graph.items()}
        isl = np.where(np.asarray(amap) == apl, np.asarray(imap), 0xFFFF).astype(np.uint16)
        isl = np.ascontiguousarray(isl)
        H, W = isl.shape
        self.npoly = len(g["pc_off"]) - 1
        h = C.c_void_p()
        ctx._ck(self.L.rtpp_cdrt_create(ctx.h, n_instances, node_capacity, queue_capacity, W, H, len(g["cut_se"]), _p(_c(g["cut_se"], np.float64)),
                                        _p(_c(g["cut_core"], np.float64)), _p(_c(g["cp_off"], np.int32)), _p(_c(g["cp"], np.int32)),
                                        _p(_c(g["cc_off"], np.int32)), _p(_c(g["cc"], np.int32)), self.npoly, _p(_c(g["pv_off"], np.int32)),
                                        _p(_c(g["pv"], np.float64)), _p(_c(g["pc_off"], np.int32)), _p(_c(g["pc"], np.int32)), _p(isl),
                                        C.c_double(resolution), C.c_double(r_th), C.c_double(p_th), nearK, RewireN, rand_capacity, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.rtpp_cdrt_destroy(self.h)
            self.h = None

    def set_agent(self, inst, x, y):
        return self.L.rtpp_cdrt_set_agent(self.h, inst, C.c_double(x), C.c_double(y))

    def set_goal(self, inst, x, y):
        return self.L.rtpp_cdrt_set_goal(self.h, inst, C.c_double(x), C.c_double(y))

    def add_obstacle(self, inst, num, typ, x, y, angle, radius, outline, check_len=2.0):
        o = _c(outline, np.float64).reshape(-1, 2)
        self.ctx._ck(self.L.rtpp_cdrt_add_obstacle(self.h, inst, num, typ, C.c_double(x), C.c_double(y), C.c_double(angle), C.c_double(radius),
                                                   _p(o), len(o), C.c_double(check_len)))

    def move_obstacle(self, inst, num, x, y, angle=0.0):
        self.ctx._ck(self.L.rtpp_cdrt_move_obstacle(self.h, inst, num, C.c_double(x), C.c_double(y), C.c_double(angle)))

    def start(self, inst, rand):
        r = _c(rand, np.int32)
        self.ctx._ck(self.L.rtpp_cdrt_start(self.h, inst, _p(r), len(r)))

    def tick(self, dt_ns, get_path=True):
        self.ctx._ck(self.L.rtpp_cdrt_tick(self.h, C.c_longlong(dt_ns), int(get_path)))

    def stats(self):
        out = np.zeros((self.n, 24), np.int64)
        self.ctx._ck(self.L.rtpp_cdrt_read_stats(self.h, _p(out)))
        return out

    def stats_dict(self, inst=0):
        return dict(zip(self.STAT, self.stats()[inst].tolist()))

    def read_cycles(self, inst=0):
        out = np.zeros(8, np.int64)
        self.ctx._ck(self.L.rtpp_cdrt_read_cycles(self.h, inst, _p(out)))
        return dict(zip(["cyc_pop", "cyc_filter", "cyc_check", "cyc_neighbours", "cyc_push", "checks", "neighbours", "rewires"], out.tolist()))

    def read_tree(self, inst=0):
        n = C.c_int()
        self.ctx._ck(self.L.rtpp_cdrt_read_tree(self.h, inst, 0, None, None, None, None, None, None, C.byref(n)))
        n = n.value
        xy, par, cost = np.zeros((n, 2)), np.zeros(n, np.int32), np.zeros(n)
        col, inv, alive = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        self.ctx._ck(self.L.rtpp_cdrt_read_tree(self.h, inst, n, _p(xy), _p(par), _p(cost), _p(col), _p(inv), _p(alive), C.byref(C.c_int())))
        return xy, par, cost, col, inv, alive

    def read_path(self, inst=0, cap=2048):
        xy = np.zeros((cap, 2))
        n = C.c_int()
        self.ctx._ck(self.L.rtpp_cdrt_read_path(self.h, inst, cap, _p(xy), C.byref(n)))
        return xy[: min(n.value, cap)].copy()

    def read_islands(self, inst=0):
        a, b = np.zeros(self.npoly, np.int32), np.zeros(self.npoly, np.int32)
        self.ctx._ck(self.L.rtpp_cdrt_read_islands(self.h, inst, _p(a), _p(b)))
        return a, b


class CdrtPlanner:
    """One instance of a CdrtEngine behind the BImap-like call surface the scripted demo loop drives
    (tests/cdrt_scenarios.py): set_params / set_clock / set_rand / add_obstacle / set_agent / set_goal / start / task /
    get_path / move_obstacle, plus dump() in Node_Map order for parity checks."""

    def __init__(self, ctx, fx, inst=0, engine=None, **kw):
        self.ctx, self.fx, self.inst, self.kw = ctx, fx, inst, kw
        self.eng, self.own = engine, engine is None
        self.params = dict(r_th=10.0, p_th=0.85, nearK=20, RewireN=100)
        self.dt, self.rand, self._pending = 20000, None, []

    def set_params(self, **p):
        self.params.update(p)

    def set_clock(self, dt_ns, n0=0):
        self.dt = dt_ns

    def set_rand(self, r):
        self.rand = _c(r, np.int32)

    def _engine(self):
        if self.eng is None:
            self.eng = CdrtEngine(self.ctx, 1, self.fx["graphs"][0], self.fx["amap"], self.fx["imap"], self.fx["resolution"],
                                  rand_capacity=len(self.rand), **self.params, **self.kw)
        return self.eng

    def add_obstacle(self, num, typ, x, y, angle, radius, outline, check_len=2.0):
        self._engine().add_obstacle(self.inst, num, typ, x, y, angle, radius, outline, check_len)

    def move_obstacle(self, num, x, y, angle=0.0):
        self.eng.move_obstacle(self.inst, num, x, y, angle)

    def set_agent(self, x, y):
        self._engine().set_agent(self.inst, x, y)

    def set_goal(self, x, y):
        self._engine().set_goal(self.inst, x, y)

    def start(self):
        self._engine().start(self.inst, self.rand)

    def task(self):
        self.eng.tick(self.dt, get_path=True)

    def get_path(self):
        s = self.eng.stats_dict(self.inst)
        cost = np.array(s["path_cost_bits"], np.int64).view(np.float64)
        return self.eng.read_path(self.inst), bool(s["path_col"]), float(cost)

    def state(self):
        return self.eng.stats_dict(self.inst)

    def dump(self):
        """Live nodes in the reference's Node_Map order (sorted by (x, y)): xy, parent xy, cost, collision, invM, #children."""
        xy, par, cost, col, inv, alive = self.eng.read_tree(self.inst)
        live = np.flatnonzero(alive)
        pxy = np.where((par >= 0)[:, None], xy[np.maximum(par, 0)], -1.0)
        nsub = np.bincount(par[live][par[live] >= 0], minlength=len(xy)).astype(np.int32)
        order = live[np.lexsort((xy[live, 1], xy[live, 0]))]
        return xy[order], pxy[order], cost[order], col[order], inv[order], nsub[order]

    def close(self):
        if self.own and self.eng is not None:
            self.eng.close()
            self.eng = None


class PlannerCDRT:
    """The C++ BImap drop-in (rtpp_b200/planners/CDRT.h: same methods as the reference class) driven from Python with
    the call surface of the scripted demo loop (tests/cdrt_scenarios.py)."""

    def __init__(self, fx, device=0):
        self.L = planners_lib()
        self.L.pl_cdrt_create.restype = C.c_void_p
        h = self.L.pl_cdrt_create(fx["W"], fx["H"], device)
        if not h:
            raise RtppError("BImap could not be created (no CUDA device; there is no CPU fallback)")
        self.h, self.fx = C.c_void_p(h), fx
        for g in fx["graphs"]:
            self.L.pl_cdrt_set_graph(self.h, len(g["cut_se"]), _p(g["cut_se"]), _p(g["cut_core"]), _p(g["cp_off"]), _p(g["cp"]), _p(g["cc_off"]),
                                     _p(g["cc"]), len(g["poly_core"]), _p(g["pv_off"]), _p(g["pv"]), _p(g["poly_core"]), _p(g["pc_off"]),
                                     _p(g["pc"]), _p(g["pp_off"]), _p(g["pp"]))
        self._rand = None

    def close(self):
        if self.h:
            self.L.pl_cdrt_destroy(self.h)
            self.h = None

    def draw_bimap(self):
        a = np.zeros((self.fx["H"], self.fx["W"]), np.uint16)
        i = np.zeros_like(a)
        self.L.pl_cdrt_draw_bimap(self.h, _p(a), _p(i))
        return a, i

    def set_params(self, r_th=10.0, p_th=0.85, nearK=20, RewireN=100):
        self.L.pl_cdrt_set_params(self.h, C.c_double(r_th), C.c_double(p_th), nearK, RewireN)

    def set_clock(self, dt_ns, n0=0):
        self.L.pl_cdrt_set_clock(self.h, C.c_longlong(dt_ns))

    def set_rand(self, r):
        self._rand = _c(r, np.int32)
        self.L.pl_cdrt_set_rand(self.h, _p(self._rand), C.c_longlong(len(self._rand)))

    def set_agent(self, x, y):
        self.L.pl_cdrt_set_agent(self.h, C.c_double(x), C.c_double(y))

    def set_goal(self, x, y):
        self.L.pl_cdrt_set_goal(self.h, C.c_double(x), C.c_double(y))

    def add_obstacle(self, num, typ, x, y, angle, radius, outline, check_len=2.0):
        o = _c(outline, np.float64).reshape(-1, 2)
        self.L.pl_cdrt_add_obstacle(self.h, num, typ, C.c_double(x), C.c_double(y), C.c_double(angle), C.c_double(radius), _p(o), len(o),
                                    C.c_double(check_len))

    def move_obstacle(self, num, x, y, angle=0.0):
        self.L.pl_cdrt_move_obstacle(self.h, num, C.c_double(x), C.c_double(y), C.c_double(angle))

    def start(self):
        self.draw_bimap()
        self.L.pl_cdrt_start(self.h, C.c_double(self.fx["resolution"]))

    def task(self):
        self.L.pl_cdrt_task(self.h)

    def get_path(self, cap=2048):
        xy = np.zeros((cap, 2))
        col, cost = C.c_int(), C.c_double()
        n = self.L.pl_cdrt_get_path(self.h, _p(xy), cap, C.byref(col), C.byref(cost))
        return xy[: min(n, cap)].copy(), bool(col.value), cost.value

    def state(self):
        out = np.zeros(24, np.int64)
        self.L.pl_cdrt_stats(self.h, _p(out))
        return dict(zip(CdrtEngine.STAT, out.tolist()))

    def dump(self):
        n = self.L.pl_cdrt_num_nodes(self.h)
        xy, par = np.zeros((n, 2)), np.zeros((n, 2))
        cost, col, inv, ns = np.zeros(n), np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32)
        assert self.L.pl_cdrt_dump(self.h, _p(xy), _p(par), _p(cost), _p(col), _p(inv), _p(ns)) == n
        return xy, par, cost, col, inv, ns
