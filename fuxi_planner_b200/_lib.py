"""ctypes binding of libfuxi_b200.so (include/fuxi_b200.h).  No CPU fallback: if the shared library
or a CUDA device is missing every entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FUXI_B200_SO: tuning experiments only (a differently-compiled build of the same sources, scripts/build_variants.sh)
SO_PATH = os.environ.get("FUXI_B200_SO") or os.path.join(_HERE, "libfuxi_b200.so")

FX_OK = 0
FX_COST_UNREACHABLE = -1
FX_COST_START_OOB = -2
FX_COST_OVERFLOW = -3
FX_EUCLID_WS = 2378
FX_EUCLID_WD = 3363

# every symbol include/fuxi_b200.h declares (tests/test_boundary.py checks the .so exports all of them)
SYMBOLS = ["fx_create", "fx_destroy", "fx_last_error", "fx_version", "fx_launch_count", "fx_set_search_tuning", "fx_set_search_form", "fx_canon_successors",
           "fx_project", "fx_inflate", "fx_edt", "fx_edt_rows", "fx_edt_cols", "fx_search_batch", "fx_field", "fx_field_relax", "fx_field_status",
           "fx_search_stats", "fx_search_kernel_ms", "fx_search_timings", "fx_plan_host_stages", "fx_plan_host", "fx_plan_host_f64", "fx_plan_host_csr", "fx_last_d2h_bytes",
           "fx_paths_compact", "fx_paths_jump_points", "fx_jump_points_host", "fx_map_host", "fx_halo_merge",
           "fx_grid_decode", "fx_grid_encode", "fx_grid_paste", "fx_grid_bbox", "fx_relocate_goal", "fx_path_post",
           "fx_replan_host", "fx_replan_grid_host", "fx_grid_to_image", "fx_image_to_grid",
           "fx_cloud_reserve", "fx_cloud_filter", "fx_cloud_filter_host", "fx_distance_filter", "fx_distance_filter_host",
           "fx_transform_filter", "fx_transform_filter_host"]


class ReplanIn(C.Structure):
    """fx_replan_in of include/fuxi_b200.h"""
    _fields_ = [("variant", C.c_int32), ("layout", C.c_int32), ("crop", C.c_int32), ("ifa", C.c_int32), ("hchoice", C.c_int32),
                ("shortcut", C.c_int32), ("origin_x", C.c_double), ("origin_y", C.c_double), ("reso", C.c_double),
                ("start_x", C.c_double), ("start_y", C.c_double), ("goal_x", C.c_double), ("goal_y", C.c_double),
                ("drop_px", C.c_double), ("drop_py", C.c_double), ("drop_pz", C.c_double), ("drop_radius", C.c_double)]


class ReplanOut(C.Structure):
    """fx_replan_out of include/fuxi_b200.h"""
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("paste_x", C.c_int32), ("paste_y", C.c_int32), ("start_x", C.c_int32),
                ("start_y", C.c_int32), ("goal_x", C.c_int32), ("goal_y", C.c_int32), ("goal_moved", C.c_int32),
                ("end_occu", C.c_int32), ("skipped", C.c_int32), ("raw_len", C.c_int32), ("path_len", C.c_int32),
                ("cost_i", C.c_int32), ("cost_f", C.c_double), ("origin_x", C.c_double), ("origin_y", C.c_double)]



class CloudParams(C.Structure):
    """fx_cloud_params of include/fuxi_b200.h (defaults = src/chen_filter_rgb.cpp:52-71)"""
    _fields_ = [("stride_floats", C.c_int32), ("rgb_offset", C.c_int32), ("pass_lo", C.c_float), ("pass_hi", C.c_float),
                ("leaf_x", C.c_float), ("leaf_y", C.c_float), ("leaf_z", C.c_float), ("min_neighbors", C.c_int32),
                ("radius", C.c_double)]


_lib = None


class FuxiError(RuntimeError):
    pass


def load():
    """dlopen the library (does not touch the GPU)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise FuxiError("libfuxi_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `python fuxi_planner_b200/build.py`. There is no CPU fallback." % SO_PATH)
    lib = C.CDLL(SO_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    lib.fx_create.argtypes = [i32, C.POINTER(vp)]
    lib.fx_destroy.argtypes = [vp]
    lib.fx_last_error.argtypes = [vp]
    lib.fx_last_error.restype = C.c_char_p
    lib.fx_version.argtypes = []
    lib.fx_launch_count.argtypes = [vp]
    lib.fx_launch_count.restype = i64
    lib.fx_set_search_tuning.argtypes = [vp, i32, i32]
    lib.fx_set_search_form.argtypes = [vp, i32]
    lib.fx_canon_successors.argtypes = [i32, i32]
    lib.fx_project.argtypes = [vp, vp, i64, i32, C.POINTER(f32), f32, f32, f32, f32, f32, i32, i32, vp, i32, vp]
    lib.fx_inflate.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.fx_edt.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.fx_edt_rows.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.fx_edt_cols.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.fx_search_batch.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp, i32, vp]
    lib.fx_field.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.fx_field_relax.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.fx_halo_merge.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.fx_field_status.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.fx_search_stats.argtypes = [vp, C.POINTER(i64)]
    lib.fx_search_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.fx_search_timings.argtypes = [vp, C.POINTER(C.c_float)]
    lib.fx_plan_host_stages.argtypes = [vp, C.POINTER(C.c_double)]
    lib.fx_plan_host.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp, i32]
    lib.fx_plan_host_f64.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp, i32]
    lib.fx_plan_host_csr.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32, vp, vp, vp, i32, vp, vp, i64, C.POINTER(i64)]
    lib.fx_last_d2h_bytes.argtypes = [vp]
    lib.fx_last_d2h_bytes.restype = i64
    lib.fx_paths_compact.argtypes = [vp, vp, vp, i32, i32, vp, vp, i64, vp]
    lib.fx_paths_jump_points.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32, vp, vp, i32, vp]
    lib.fx_jump_points_host.argtypes = [vp, vp, i32, i32, vp, i32, vp, i32, C.POINTER(i32)]
    lib.fx_map_host.argtypes = [vp, vp, i64, i32, C.POINTER(f32), f32, f32, f32, f32, f32, i32, i32, i32, i32, vp]
    f64p = C.POINTER(C.c_double)
    lib.fx_grid_decode.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, i32, vp]
    lib.fx_grid_encode.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.fx_grid_paste.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, i32, vp]
    lib.fx_grid_bbox.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.fx_relocate_goal.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp]
    lib.fx_path_post.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32, i32, f64p, f64p, vp, vp, vp, vp]
    lib.fx_replan_host.argtypes = [vp, vp, i32, i32, C.POINTER(ReplanIn), C.POINTER(ReplanOut), vp, vp, i32]
    lib.fx_replan_grid_host.argtypes = [vp, vp, C.c_size_t, C.POINTER(i32), C.POINTER(i32)]
    lib.fx_grid_to_image.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.fx_image_to_grid.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.fx_cloud_reserve.argtypes = [vp, i64]
    lib.fx_cloud_filter.argtypes = [vp, vp, i64, C.POINTER(CloudParams), vp, i64, vp, vp]
    lib.fx_cloud_filter_host.argtypes = [vp, vp, i64, C.POINTER(CloudParams), vp, i64, C.POINTER(i64)]
    lib.fx_distance_filter.argtypes = [vp, vp, i64, C.c_double, vp, vp, vp]
    lib.fx_distance_filter_host.argtypes = [vp, vp, i64, C.c_double, vp, C.POINTER(i64)]
    lib.fx_transform_filter.argtypes = [vp, vp, i64, i32, i32, f64p, f64p, f64p, C.c_double, C.c_double, C.c_double, vp, vp, vp]
    lib.fx_transform_filter_host.argtypes = [vp, vp, i64, i32, i32, f64p, f64p, f64p, C.c_double, C.c_double, C.c_double, vp, C.POINTER(i64)]
    for s in SYMBOLS:
        if s not in ("fx_last_error", "fx_launch_count", "fx_last_d2h_bytes"):
            getattr(lib, s).restype = i32
    _lib = lib
    return lib


class Context:
    """One fx_context bound to one CUDA device."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.fx_create(int(device), C.byref(h))
        if rc != FX_OK:
            raise FuxiError("fx_create(%d) failed (%d): %s" % (device, rc, self.lib.fx_last_error(None).decode()))
        self.handle = h
        self.device = int(device)

    def check(self, rc, what):
        if rc != FX_OK:
            raise FuxiError("%s failed (%d): %s" % (what, rc, self.lib.fx_last_error(self.handle).decode()))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.fx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_search_form(self, form):
        """"auto" (by batch size), "throughput" (forward searches only: forward-canonical paths) or "latency" (bidirectional)."""
        self.check(self.lib.fx_set_search_form(self.handle, {"auto": 0, "throughput": 1, "latency": 2}[form]), "fx_set_search_form")

    @property
    def launches(self):
        return int(self.lib.fx_launch_count(self.handle))


_default = {}


def default_context(device=0):
    ctx = _default.get(device)
    if ctx is None or ctx.handle is None:
        ctx = Context(device)
        _default[device] = ctx
    return ctx
