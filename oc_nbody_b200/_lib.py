"""ctypes binding of liboc_nbody_b200.so (the C ABI declared in include/ocg.h).

There is NO CPU fallback: if the shared library is missing or no B200 is visible the calls raise.
torch tensors are used purely as device buffers (``.data_ptr()``) and for the current stream.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liboc_nbody_b200.so")
# the -DOCG_TUNING build (every sweep shape + timing-only experiments); only tools/ ask for it, via OCG_TUNING_LIB=1
LIB_PATH_TUNING = os.path.join(_HERE, "lib", "liboc_nbody_b200_tuning.so")

KERNEL_PLUMMER = 0
KERNEL_SPLINE = 1
KERNELS = {"plummer": KERNEL_PLUMMER, "spline": KERNEL_SPLINE}

# every symbol include/ocg.h declares (tests check the library exports them all)
ABI_SYMBOLS = [
    "ocg_version", "ocg_create", "ocg_destroy", "ocg_last_error", "ocg_device_info", "ocg_launch_count", "ocg_capture_epoch",
    "ocg_last_direct_kernel_ms", "ocg_set_kernel_timing", "ocg_set_source_shards", "ocg_last_direct_traffic_bytes", "ocg_recentre_f64", "ocg_cast_f64_f32", "ocg_assemble_sources",
    "ocg_field_direct", "ocg_frame_subtract", "ocg_field_build_host", "ocg_pack_planes", "ocg_grid_time_blend",
    "ocg_grid_interp", "ocg_grid_interp_multi", "ocg_grid_interp_nested", "ocg_pack_planes_indexed", "ocg_set_interp_weight_slots", "ocg_grid_interp_slot", "ocg_grid_interp_rbf", "ocg_grid_interp_rbf_nested", "ocg_self_gravity", "ocg_self_gravity_hermite", "ocg_hermite_predict", "ocg_hermite_correct", "ocg_hermite_block_evolve", "ocg_bound_com", "ocg_eject_mask", "ocg_compact_rows", "ocg_kick", "ocg_drift", "ocg_axpy", "ocg_probe_throughput", "ocg_comm_create", "ocg_comm_connect", "ocg_comm_destroy", "ocg_comm_info", "ocg_comm_status", "ocg_comm_allreduce_f64", "ocg_self_gravity_sharded", "ocg_self_gravity_hermite_sharded",
]


# include/ocg_debug.h (test / tuning hooks; not part of the drop-in ABI)
DEBUG_SYMBOLS = ["ocg_debug_set", "ocg_debug_variant_count", "ocg_debug_variant_name", "ocg_debug_variant_built",
                 "ocg_debug_rbf_phase_cycles"]
KNOBS = {"direct_variant": 0, "precise_near": 1, "mass_fold": 2, "small_cluster_path": 3, "host_chunk": 4,
         "hermite_variant": 5, "hermite_small_path": 6, "interp_variant": 7, "near_cap": 9, "pass_bytes": 10}


class OcgError(RuntimeError):
    pass


class _GridDesc(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32 * 3), ("n_cluster", ctypes.c_int32), ("node_dev", ctypes.c_void_p * 3),
                ("origin_dev", ctypes.c_void_p)]


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH_TUNING if os.environ.get("OCG_TUNING_LIB") == "1" else LIB_PATH
    if not os.path.exists(path):
        raise OcgError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(oc_nbody_b200 has no CPU fallback)" % path)
    L = ctypes.CDLL(path)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    L.ocg_version.restype = ctypes.c_int
    L.ocg_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.ocg_destroy.argtypes = [vp]
    L.ocg_last_error.restype = ctypes.c_char_p
    L.ocg_last_error.argtypes = [vp]
    L.ocg_device_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(i64)]
    L.ocg_launch_count.restype = i64
    L.ocg_launch_count.argtypes = [vp]
    L.ocg_capture_epoch.restype = i64
    L.ocg_capture_epoch.argtypes = [vp]
    L.ocg_last_direct_kernel_ms.restype = dbl
    L.ocg_last_direct_kernel_ms.argtypes = [vp]
    L.ocg_set_kernel_timing.argtypes = [vp, ctypes.c_int]
    L.ocg_set_source_shards.argtypes = [vp, i32]
    L.ocg_last_direct_traffic_bytes.restype = i64
    L.ocg_last_direct_traffic_bytes.argtypes = [vp]
    L.ocg_recentre_f64.argtypes = [vp, vp, vp, i64, ctypes.POINTER(dbl), vp, vp]
    L.ocg_cast_f64_f32.argtypes = [vp, vp, i64, vp, vp]
    L.ocg_assemble_sources.argtypes = [vp, vp, vp, vp, vp, i64, i64, dbl, i32, dbl, dbl, ctypes.POINTER(dbl), vp, vp, i64,
                                       ctypes.POINTER(i64), vp]
    L.ocg_field_direct.argtypes = [vp, vp, vp, i64, vp, i64, ctypes.c_int, dbl, vp, vp, ctypes.c_int, vp]
    L.ocg_frame_subtract.argtypes = [vp, vp, i64, i64, vp]
    L.ocg_field_build_host.argtypes = [vp, vp, vp, vp, i64, vp, i64, ctypes.POINTER(dbl), i64, ctypes.c_int, dbl, vp, vp]
    L.ocg_pack_planes.argtypes = [vp, vp, vp, i64, vp, vp]
    L.ocg_grid_time_blend.argtypes = [vp, vp, vp, dbl, i64, vp, vp, vp]
    L.ocg_grid_interp.argtypes = [vp, ctypes.POINTER(_GridDesc), vp, vp, dbl, vp, vp, vp, vp, i64, vp, vp, vp, vp]
    L.ocg_grid_interp_multi.argtypes = [vp, ctypes.POINTER(_GridDesc), vp, ctypes.POINTER(dbl), i32, vp, vp, vp, vp, i64, vp, vp,
                                        vp, vp]
    L.ocg_grid_interp_nested.argtypes = [vp, ctypes.POINTER(_GridDesc), ctypes.POINTER(_GridDesc), vp, vp, ctypes.POINTER(dbl),
                                         i32, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
    L.ocg_pack_planes_indexed.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    L.ocg_set_interp_weight_slots.argtypes = [vp, ctypes.POINTER(dbl), i32, i32, vp]
    L.ocg_grid_interp_slot.argtypes = [vp, ctypes.POINTER(_GridDesc), ctypes.POINTER(_GridDesc), vp, vp, i32, i32, vp, vp, vp, vp,
                                       i64, vp, vp, vp]
    L.ocg_self_gravity.argtypes = [vp, vp, vp, i64, vp, i32, dbl, dbl, i64, i64, vp, vp, vp]
    L.ocg_grid_interp_rbf.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
    L.ocg_grid_interp_rbf_nested.argtypes = [vp, vp, vp, vp, i64, i64, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
    L.ocg_self_gravity_hermite.argtypes = [vp, vp, vp, vp, i64, vp, i32, dbl, dbl, dbl, i64, i64, vp, vp, vp, vp]
    L.ocg_hermite_predict.argtypes = [vp, vp, vp, vp, vp, i64, dbl, dbl, vp, vp, vp]
    L.ocg_hermite_correct.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, vp, vp]
    L.ocg_hermite_block_evolve.argtypes = [vp, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, dbl, dbl, i32, ctypes.POINTER(i64),
                                           ctypes.POINTER(i64), vp]
    L.ocg_debug_rbf_phase_cycles.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.ocg_debug_set.argtypes = [vp, ctypes.c_int, i64]
    L.ocg_debug_variant_count.argtypes = [ctypes.c_int]
    L.ocg_debug_variant_name.restype = ctypes.c_char_p
    L.ocg_debug_variant_name.argtypes = [ctypes.c_int, ctypes.c_int]
    L.ocg_debug_variant_built.argtypes = [ctypes.c_int, ctypes.c_int]
    L.ocg_bound_com.argtypes = [vp, vp, vp, vp, vp, i64, vp, i32, dbl, vp, vp, vp]
    L.ocg_eject_mask.argtypes = [vp, vp, i64, dbl, dbl, vp, vp, vp]
    L.ocg_compact_rows.argtypes = [vp, vp, i32, i64, vp, vp, i64, vp, vp]
    L.ocg_kick.argtypes = [vp, vp, vp, i64, dbl, vp]
    L.ocg_drift.argtypes = [vp, vp, vp, i64, dbl, dbl, vp]
    L.ocg_axpy.argtypes = [vp, vp, vp, dbl, i64, vp]
    L.ocg_comm_create.argtypes = [vp, i32, i32, i64, vp]
    L.ocg_comm_connect.argtypes = [vp, vp]
    L.ocg_comm_destroy.argtypes = [vp]
    L.ocg_comm_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i64)]
    L.ocg_comm_status.argtypes = [vp, vp]
    L.ocg_comm_allreduce_f64.argtypes = [vp, vp, i64, vp]
    L.ocg_self_gravity_sharded.argtypes = [vp, vp, vp, i64, dbl, dbl, vp, vp, vp]
    L.ocg_self_gravity_hermite_sharded.argtypes = [vp, vp, vp, vp, i64, dbl, dbl, dbl, vp, vp, vp, vp]
    L.ocg_probe_throughput.restype = dbl
    L.ocg_probe_throughput.argtypes = [vp, ctypes.c_int]
    _lib = L
    return L


def _dptr(t):
    """Device pointer of a torch CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise OcgError("expected a CUDA tensor (device buffer), got a %s tensor" % t.device)
    if not t.is_contiguous():
        raise OcgError("device buffers must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def _hptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _vec3(v):
    return (ctypes.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


class Context:
    """One ocg_ctx: one GPU, one process, not thread-safe (include/ocg.h)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.ocg_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise OcgError("ocg_create(%d) failed (%d): %s" % (device, rc, self.lib.ocg_last_error(None).decode()))
        self.h = h
        self.device = int(device)
        sm, khz, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
        self.lib.ocg_device_info(h, ctypes.byref(sm), ctypes.byref(khz), ctypes.byref(mem))
        self.sm_count, self.sm_clock_khz, self.global_mem = sm.value, khz.value, mem.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.ocg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise OcgError("%s failed (%d): %s" % (what, rc, self.lib.ocg_last_error(self.h).decode()))

    @staticmethod
    def _stream():
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- test / tuning hooks (include/ocg_debug.h): state of THIS ctx ----
    def debug_set(self, knob, value):
        self._ck(self.lib.ocg_debug_set(self.h, KNOBS[knob], int(value)), "ocg_debug_set(%s)" % knob)

    def variant_count(self, family=0):
        return int(self.lib.ocg_debug_variant_count(int(family)))

    def variant_name(self, vid, family=0):
        return self.lib.ocg_debug_variant_name(int(family), int(vid)).decode()

    def variant_built(self, vid, family=0):
        return bool(self.lib.ocg_debug_variant_built(int(family), int(vid)))

    # ---- bookkeeping ----
    def launch_count(self):
        return int(self.lib.ocg_launch_count(self.h))

    def capture_epoch(self):
        """Changes whenever something a captured CUDA graph of this ctx's calls froze has changed (include/ocg.h)."""
        return int(self.lib.ocg_capture_epoch(self.h))

    def set_kernel_timing(self, on):
        self._ck(self.lib.ocg_set_kernel_timing(self.h, 1 if on else 0), "ocg_set_kernel_timing")

    def set_source_shards(self, n_shards):
        """This context sums 1/n_shards of the sources of a source-sharded field build (include/ocg.h)."""
        self._ck(self.lib.ocg_set_source_shards(self.h, int(n_shards)), "ocg_set_source_shards")

    def last_direct_kernel_ms(self):
        return float(self.lib.ocg_last_direct_kernel_ms(self.h))

    def last_direct_traffic_model(self):
        return int(self.lib.ocg_last_direct_traffic_bytes(self.h))

    def probe_throughput(self, which):
        v = float(self.lib.ocg_probe_throughput(self.h, int(which)))
        if v < 0:
            raise OcgError("ocg_probe_throughput failed: %s" % self.lib.ocg_last_error(self.h).decode())
        return v

    # ---- device-buffer calls (torch tensors as buffers) ----
    def recentre_f64(self, pos, mass, center, out):
        n = pos.shape[0]
        self._ck(self.lib.ocg_recentre_f64(self.h, _dptr(pos), _dptr(mass), n, _vec3(center), _dptr(out), self._stream()),
                 "ocg_recentre_f64")

    def cast_f64_f32(self, src, out):
        self._ck(self.lib.ocg_cast_f64_f32(self.h, _dptr(src), src.numel(), _dptr(out), self._stream()), "ocg_cast_f64_f32")

    SOFT_RULES = {"constant": 0, "mass_cube_root": 1, "gas_smoothing": 2}

    def assemble_sources(self, pos, mass, ids, hsml, exclude_id, rmax, rule, soft_param, soft_scale, center, out_xyzm, out_soft,
                         out_offset=0):
        """One species of the source assembly on the device (ocg_assemble_sources); returns the number of particles kept."""
        kept = ctypes.c_int64(0)
        self._ck(self.lib.ocg_assemble_sources(self.h, _dptr(pos), _dptr(mass), _dptr(ids), _dptr(hsml), pos.shape[0],
                                               int(exclude_id), float(rmax), self.SOFT_RULES[rule], float(soft_param),
                                               float(soft_scale), _vec3(center), _dptr(out_xyzm), _dptr(out_soft), int(out_offset),
                                               ctypes.byref(kept), self._stream()), "ocg_assemble_sources")
        return int(kept.value)

    def field_direct(self, src_xyzm, src_soft, tgt_xyzw, kernel, G, acc, pot=None, accumulate=False):
        n_src, n_tgt = src_xyzm.shape[0], tgt_xyzw.shape[0]
        self._ck(self.lib.ocg_field_direct(self.h, _dptr(src_xyzm), _dptr(src_soft), n_src, _dptr(tgt_xyzw), n_tgt,
                                           int(kernel), float(G), _dptr(acc), _dptr(pot), 1 if accumulate else 0,
                                           self._stream()), "ocg_field_direct")

    def frame_subtract(self, acc, center_row):
        self._ck(self.lib.ocg_frame_subtract(self.h, _dptr(acc), acc.shape[1], int(center_row), self._stream()),
                 "ocg_frame_subtract")

    def pack_planes(self, acc, pot, rec):
        self._ck(self.lib.ocg_pack_planes(self.h, _dptr(acc), _dptr(pot), acc.shape[1], _dptr(rec), self._stream()),
                 "ocg_pack_planes")

    def grid_time_blend(self, rec_a, rec_b, w_b, acc_out, pot_out=None):
        self._ck(self.lib.ocg_grid_time_blend(self.h, _dptr(rec_a), _dptr(rec_b), float(w_b), rec_a.shape[-2],
                                              _dptr(acc_out), _dptr(pot_out), self._stream()), "ocg_grid_time_blend")

    def grid_interp(self, n, nodes, origin, rec_a, rec_b, w_b, sx, sy, sz, star_cluster, acc_out, pot_out=None,
                    cell_out=None):
        d = _GridDesc()
        for k in range(3):
            d.n[k] = int(n[k])
            d.node_dev[k] = nodes[k].data_ptr()
        d.n_cluster = int(origin.shape[0])
        d.origin_dev = origin.data_ptr()
        self._ck(self.lib.ocg_grid_interp(self.h, ctypes.byref(d), _dptr(rec_a), _dptr(rec_b), float(w_b), _dptr(sx),
                                          _dptr(sy), _dptr(sz), _dptr(star_cluster), sx.shape[0], _dptr(acc_out),
                                          _dptr(pot_out), _dptr(cell_out), self._stream()), "ocg_grid_interp")

    def grid_interp_multi(self, n, nodes, origin, recs, weights, sx, sy, sz, star_cluster, acc_out, pot_out=None):
        """K3 with 1..4 record planes and their time weights (cubic B-spline blend)."""
        d = _GridDesc()
        for k in range(3):
            d.n[k] = int(n[k])
            d.node_dev[k] = nodes[k].data_ptr()
        d.n_cluster = int(origin.shape[0])
        d.origin_dev = origin.data_ptr()
        ptrs = (ctypes.c_void_p * len(recs))(*[_dptr(r).value for r in recs])
        w = (ctypes.c_double * len(recs))(*[float(x) for x in weights])
        self._ck(self.lib.ocg_grid_interp_multi(self.h, ctypes.byref(d), ptrs, w, len(recs), _dptr(sx), _dptr(sy), _dptr(sz),
                                                _dptr(star_cluster), sx.shape[0], _dptr(acc_out), _dptr(pot_out), None,
                                                self._stream()), "ocg_grid_interp_multi")

    @staticmethod
    def _grid_desc(n, nodes, origin):
        d = _GridDesc()
        for k in range(3):
            d.n[k] = int(n[k])
            d.node_dev[k] = nodes[k].data_ptr()
        d.n_cluster = int(origin.shape[0])
        d.origin_dev = origin.data_ptr()
        return d

    def grid_interp_nested(self, n, nodes, origin, recs, weights, sx, sy, sz, star_cluster, acc_out, pot_out=None,
                           fine_n=None, fine_nodes=None, recs_fine=None, tensor_out=None, level_out=None, cell_out=None):
        """K3 on the reference's two-level grid (fine_* None: single level) with optional tidal tensor [9, n_star],
        level and cell outputs."""
        dc = self._grid_desc(n, nodes, origin)
        ptrs = (ctypes.c_void_p * len(recs))(*[_dptr(r).value for r in recs])
        w = (ctypes.c_double * len(recs))(*[float(x) for x in weights])
        if fine_n is not None:
            df = self._grid_desc(fine_n, fine_nodes, origin)
            if recs_fine is None or len(recs_fine) != len(recs):
                raise OcgError("grid_interp_nested: need as many fine as coarse record planes")
            fptrs = (ctypes.c_void_p * len(recs))(*[_dptr(r).value for r in recs_fine])
            dfp = ctypes.byref(df)
        else:
            fptrs, dfp = None, None
        self._ck(self.lib.ocg_grid_interp_nested(self.h, ctypes.byref(dc), dfp, ptrs, fptrs, w, len(recs), _dptr(sx), _dptr(sy),
                                                 _dptr(sz), _dptr(star_cluster), sx.shape[0], _dptr(acc_out), _dptr(pot_out),
                                                 _dptr(tensor_out), _dptr(level_out), _dptr(cell_out), self._stream()),
                 "ocg_grid_interp_nested")

    def grid_interp_rbf(self, n, nodes, origin, field, sx, sy, sz, star_cluster, out, nclose=150, order=5, phs=3,
                        include_origin=True, tensor_out=None, status_out=None, neighbors_out=None, embedded=False):
        """K7: the reference's own spatial interpolation (kNN + polyharmonic-spline RBF, gizmo_interface.py:651-717).
        field fp64 [n_comp, n_cluster * n_node] (or [n_comp, n_node]); out [n_comp, n_star]."""
        dc = self._grid_desc(n, nodes, origin)
        self._ck(self.lib.ocg_grid_interp_rbf(self.h, ctypes.byref(dc), _dptr(field), int(field.shape[0]), int(nclose), int(order),
                                              int(phs), 1 if include_origin else 0, 1 if embedded else 0, _dptr(sx), _dptr(sy), _dptr(sz),
                                              _dptr(star_cluster), sx.shape[0], _dptr(out), _dptr(tensor_out), _dptr(status_out),
                                              _dptr(neighbors_out), self._stream()), "ocg_grid_interp_rbf")

    def grid_interp_rbf_nested(self, n, nodes, fine_n, fine_nodes, origin, coarse_row, fine_row0, field, sx, sy, sz, star_cluster, out,
                               nclose=150, order=5, phs=3, include_origin=True, tensor_out=None, status_out=None, neighbors_out=None):
        """K7 on the reference's nested grid, both levels searched (mixed-level stencils near the fine-box surface).
        field fp64 [n_comp, n_cluster * n_point] over the whole point list; coarse_row int32 [n_coarse_lattice]."""
        dc, df = self._grid_desc(n, nodes, origin), self._grid_desc(fine_n, fine_nodes, origin)
        n_point = field.shape[1] // int(origin.shape[0])
        self._ck(self.lib.ocg_grid_interp_rbf_nested(self.h, ctypes.byref(dc), ctypes.byref(df), _dptr(coarse_row), int(fine_row0),
                                                     int(n_point), _dptr(field), int(field.shape[0]), int(nclose), int(order), int(phs),
                                                     1 if include_origin else 0, _dptr(sx), _dptr(sy), _dptr(sz), _dptr(star_cluster),
                                                     sx.shape[0], _dptr(out), _dptr(tensor_out), _dptr(status_out), _dptr(neighbors_out),
                                                     self._stream()), "ocg_grid_interp_rbf_nested")

    def set_interp_weight_slots(self, weights, first_slot=0):
        """weights: sequence of up-to-4-element weight lists, written to constant-memory slots first_slot.. (stream-ordered)."""
        flat = []
        for w in weights:
            w = [float(x) for x in w]
            flat.extend(w + [0.0] * (4 - len(w)))
        arr = (ctypes.c_double * len(flat))(*flat)
        self._ck(self.lib.ocg_set_interp_weight_slots(self.h, arr, int(first_slot), len(weights), self._stream()),
                 "ocg_set_interp_weight_slots")

    def grid_interp_slot(self, n, nodes, origin, recs, w_slot, sx, sy, sz, star_cluster, acc_out, pot_out=None, fine_n=None,
                         fine_nodes=None, recs_fine=None):
        """K3 with the time-blend weights taken from constant-memory slot `w_slot` (graph-replayable)."""
        dc = self._grid_desc(n, nodes, origin)
        ptrs = (ctypes.c_void_p * len(recs))(*[_dptr(r).value for r in recs])
        if fine_n is not None:
            df = self._grid_desc(fine_n, fine_nodes, origin)
            fptrs = (ctypes.c_void_p * len(recs))(*[_dptr(r).value for r in recs_fine])
            dfp = ctypes.byref(df)
        else:
            fptrs, dfp = None, None
        self._ck(self.lib.ocg_grid_interp_slot(self.h, ctypes.byref(dc), dfp, ptrs, fptrs, int(w_slot), len(recs), _dptr(sx),
                                               _dptr(sy), _dptr(sz), _dptr(star_cluster), sx.shape[0], _dptr(acc_out),
                                               _dptr(pot_out), self._stream()), "ocg_grid_interp_slot")

    def pack_planes_indexed(self, acc, pot, index, rec):
        """rec[index[i]] = float4(acc[:, i], pot[i]) (K2 pack with a scatter)."""
        self._ck(self.lib.ocg_pack_planes_indexed(self.h, _dptr(acc), _dptr(pot), acc.shape[1], _dptr(index), _dptr(rec),
                                                  self._stream()), "ocg_pack_planes_indexed")

    def self_gravity(self, pos, mass, eps2, G, acc, pot=None, seg_offsets=None, tgt_begin=0, tgt_end=None):
        n = pos.shape[1]
        if seg_offsets is None:
            seg, n_seg = None, 1
        else:
            seg = np.ascontiguousarray(seg_offsets, dtype=np.int64)
            n_seg = len(seg) - 1
        tgt_end = n if tgt_end is None else tgt_end
        self._ck(self.lib.ocg_self_gravity(self.h, _dptr(pos), _dptr(mass), n, _hptr(seg), n_seg, float(eps2), float(G),
                                           int(tgt_begin), int(tgt_end), _dptr(acc), _dptr(pot), self._stream()),
                 "ocg_self_gravity")

    def self_gravity_hermite(self, pos, vel, mass, eps2, G, vel_to_len, acc, jerk, pot=None, seg_offsets=None,
                             tgt_begin=0, tgt_end=None):
        """K6: acceleration and jerk (and optionally the potential) of the cluster's self-gravity."""
        n = pos.shape[1]
        if seg_offsets is None:
            seg, n_seg = None, 1
        else:
            seg = np.ascontiguousarray(seg_offsets, dtype=np.int64)
            n_seg = len(seg) - 1
        tgt_end = n if tgt_end is None else tgt_end
        self._ck(self.lib.ocg_self_gravity_hermite(self.h, _dptr(pos), _dptr(vel), _dptr(mass), n, _hptr(seg), n_seg,
                                                   float(eps2), float(G), float(vel_to_len), int(tgt_begin), int(tgt_end),
                                                   _dptr(acc), _dptr(jerk), _dptr(pot), self._stream()),
                 "ocg_self_gravity_hermite")

    def hermite_predict(self, pos, vel, acc, jerk, dt, vel_to_len, pos_pred, vel_pred):
        self._ck(self.lib.ocg_hermite_predict(self.h, _dptr(pos), _dptr(vel), _dptr(acc), _dptr(jerk), pos.shape[1], float(dt),
                                              float(vel_to_len), _dptr(pos_pred), _dptr(vel_pred), self._stream()),
                 "ocg_hermite_predict")

    def hermite_correct(self, pos, vel, acc0, jerk0, pos_pred, vel_pred, acc1, jerk1, dt, vel_to_len, eta, dt_min=None):
        self._ck(self.lib.ocg_hermite_correct(self.h, _dptr(pos), _dptr(vel), _dptr(acc0), _dptr(jerk0), _dptr(pos_pred),
                                              _dptr(vel_pred), _dptr(acc1), _dptr(jerk1), pos.shape[1], float(dt),
                                              float(vel_to_len), float(eta), _dptr(dt_min), self._stream()),
                 "ocg_hermite_correct")

    def hermite_block_evolve(self, pos, vel, mass, acc, jerk, eps2, G, vel_to_len, span, eta=0.14, max_level=12):
        """ph4's individual block time steps over `span`; returns (block steps, star-steps) taken."""
        steps, star_steps = ctypes.c_int64(0), ctypes.c_int64(0)
        self._ck(self.lib.ocg_hermite_block_evolve(self.h, _dptr(pos), _dptr(vel), _dptr(mass), _dptr(acc), _dptr(jerk), pos.shape[1],
                                                   float(eps2), float(G), float(vel_to_len), float(span), float(eta), int(max_level),
                                                   ctypes.byref(steps), ctypes.byref(star_steps), self._stream()),
                 "ocg_hermite_block_evolve")
        return int(steps.value), int(star_steps.value)

    def bound_com(self, pos, vel, mass, pot, pot_to_v2, out, seg_offsets=None, bound_mask=None):
        """out [n_seg, 8] fp64 device: bound-subset COM (3), bound mass, bound count, COM velocity (3)."""
        n_seg = 1 if seg_offsets is None else seg_offsets.shape[0] - 1
        self._ck(self.lib.ocg_bound_com(self.h, _dptr(pos), _dptr(vel), _dptr(mass), _dptr(pot), pos.shape[1], _dptr(seg_offsets),
                                        n_seg, float(pot_to_v2), _dptr(out), _dptr(bound_mask), self._stream()), "ocg_bound_com")

    def eject_mask(self, pos, len_scale, cut, keep_mask, median_out=None):
        self._ck(self.lib.ocg_eject_mask(self.h, _dptr(pos), pos.shape[1], float(len_scale), float(cut), _dptr(keep_mask),
                                         _dptr(median_out), self._stream()), "ocg_eject_mask")

    def compact_rows(self, rows_in, keep_mask, rows_out, n_keep):
        """rows_in [R, n] fp64 -> rows_out [R, n_out] (or None: count only); n_keep: int64 device scalar."""
        R, n = rows_in.shape
        self._ck(self.lib.ocg_compact_rows(self.h, _dptr(rows_in), R, n, _dptr(keep_mask), _dptr(rows_out),
                                           0 if rows_out is None else rows_out.shape[1], _dptr(n_keep), self._stream()),
                 "ocg_compact_rows")

    def kick(self, vel, acc, dt):
        self._ck(self.lib.ocg_kick(self.h, _dptr(vel), _dptr(acc), vel.shape[1], float(dt), self._stream()), "ocg_kick")

    def drift(self, pos, vel, dt, vel_to_len=1.0):
        self._ck(self.lib.ocg_drift(self.h, _dptr(pos), _dptr(vel), pos.shape[1], float(dt), float(vel_to_len),
                                    self._stream()), "ocg_drift")

    def axpy(self, y, x, a):
        self._ck(self.lib.ocg_axpy(self.h, _dptr(y), _dptr(x), float(a), y.numel(), self._stream()), "ocg_axpy")

    # ---- multi-GPU exchange over peer memory (include/ocg.h: ocg_comm_*) ----
    COMM_HANDLE_BYTES = 128

    def comm_create(self, rank, nranks, window_bytes):
        """Allocate this rank's exchange window; returns the opaque handle (bytes) every other rank needs."""
        buf = ctypes.create_string_buffer(self.COMM_HANDLE_BYTES)
        self._ck(self.lib.ocg_comm_create(self.h, int(rank), int(nranks), int(window_bytes), buf), "ocg_comm_create")
        self.comm_rank, self.comm_nranks = int(rank), int(nranks)
        return buf.raw

    def comm_connect(self, handles):
        """handles: the nranks handle blobs in rank order (all-gathered by the caller)."""
        blob = b"".join(handles)
        self._ck(self.lib.ocg_comm_connect(self.h, ctypes.c_char_p(blob)), "ocg_comm_connect")
        self.comm_connected = True

    def comm_destroy(self):
        self._ck(self.lib.ocg_comm_destroy(self.h), "ocg_comm_destroy")
        self.comm_connected = False

    def comm_status(self):
        self._ck(self.lib.ocg_comm_status(self.h, self._stream()), "ocg_comm_status")

    def comm_allreduce_f64(self, buf):
        """In-place fp64 sum over the ranks, rank order (bit-identical on all ranks)."""
        self._ck(self.lib.ocg_comm_allreduce_f64(self.h, _dptr(buf), buf.numel(), self._stream()), "ocg_comm_allreduce_f64")

    def self_gravity_sharded(self, pos_local, mass_all, eps2, G, acc_local, pot_local=None):
        """K4 for this rank's block of stars, the position gather fused into the tile pack (peer memory over NVLink)."""
        self._ck(self.lib.ocg_self_gravity_sharded(self.h, _dptr(pos_local), _dptr(mass_all), mass_all.shape[0], float(eps2),
                                                   float(G), _dptr(acc_local), _dptr(pot_local), self._stream()),
                 "ocg_self_gravity_sharded")

    def self_gravity_hermite_sharded(self, pos_local, vel_local, mass_all, eps2, G, vel_to_len, acc_all, jerk_all, pot_all=None):
        """K6 for this rank's block of stars, positions + velocities gathered over peer memory inside the tile pack.  The
        outputs are full-size [3, n] / [n] arrays of which the rank's rows [a, b) are written."""
        self._ck(self.lib.ocg_self_gravity_hermite_sharded(self.h, _dptr(pos_local), _dptr(vel_local), _dptr(mass_all),
                                                           mass_all.shape[0], float(eps2), float(G), float(vel_to_len),
                                                           _dptr(acc_all), _dptr(jerk_all), _dptr(pot_all), self._stream()),
                 "ocg_self_gravity_hermite_sharded")

    # ---- host-buffer call (numpy in, numpy out; H2D/D2H inside) ----
    def field_build_host(self, src_pos, src_mass, src_soft, tgt_pos, center, center_row, kernel, G, want_pot=False):
        """_populate_grid_acceleration_ (gizmo_interface.py:512-573) in one call. Returns acc [3,n_tgt] (, pot)."""
        sp = np.ascontiguousarray(src_pos, np.float64).reshape(-1, 3)
        sm = np.ascontiguousarray(src_mass, np.float64)
        ss = None if src_soft is None else np.ascontiguousarray(src_soft, np.float64)
        tp = np.ascontiguousarray(tgt_pos, np.float64).reshape(-1, 3)
        if sm.shape[0] != sp.shape[0] or (ss is not None and ss.shape[0] != sp.shape[0]):
            raise OcgError("source position/mass/softening lengths differ: %d / %d / %s"
                           % (sp.shape[0], sm.shape[0], None if ss is None else ss.shape[0]))
        acc = np.empty((3, tp.shape[0]), np.float64)
        pot = np.empty(tp.shape[0], np.float64) if want_pot else None
        self._ck(self.lib.ocg_field_build_host(self.h, _hptr(sp), _hptr(sm), _hptr(ss), sp.shape[0], _hptr(tp), tp.shape[0],
                                               _vec3(center), int(center_row), int(kernel), float(G), _hptr(acc),
                                               _hptr(pot)), "ocg_field_build_host")
        return (acc, pot) if want_pot else acc


def private_context(device=None):
    """A ctx of its own for one code object (scratch + settings only): what a captured CUDA graph wants, so that no
    other caller can move or overwrite the buffers the graph froze."""
    return Context(int(os.environ.get("LOCAL_RANK", "0")) if device is None else device)


_default_ctx = {}


def default_context(device=None):
    """Process-wide ctx per device (LOCAL_RANK-aware)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
