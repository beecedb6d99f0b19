"""fuxi_planner_b200 -- B200-native (sm_100a) implementation of FUXI's global-planning hot path:
point cloud -> 2D occupancy grid -> obstacle inflation -> batched shortest paths on the 8-connected grid.

Everything compute runs in hand-written CUDA behind the C ABI in include/fuxi_b200.h
(libfuxi_b200.so, built in-tree by fuxi_planner_b200/build.py).  There is no CPU fallback.
"""
from ._lib import (Context, FuxiError, default_context, load, SO_PATH,  # noqa: F401
                   FX_EUCLID_WS, FX_EUCLID_WD)
from .api import (PlanResult, edt, field, field_relax, field_status, inflate, map_host, plan_batch, plan_host,  # noqa: F401
                  project, search_stats, search_kernel_ms, search_timings, plan_host_stages, grid_decode, grid_encode, grid_paste, grid_bbox, relocate_goal, path_post,
                  replan_host, grid_to_image, image_to_grid, paths_compact, plan_host_csr, paths_jump_points, jump_points_host)
from . import jps1  # noqa: F401
from . import tiled  # noqa: F401
from . import cloud, planner, formats  # noqa: F401
from .cloud import cloud_affine, cloud_to_grid  # noqa: F401
from .planner import inflate_host, replan  # noqa: F401

__all__ = ["Context", "FuxiError", "default_context", "load", "SO_PATH", "PlanResult", "edt", "field", "field_relax",
           "field_status", "inflate", "map_host", "plan_batch", "plan_host", "project", "search_stats", "jps1", "tiled", "cloud", "planner", "cloud_affine", "cloud_to_grid", "inflate_host", "replan",
           "formats", "grid_to_image", "image_to_grid", "grid_decode", "grid_encode", "grid_paste", "grid_bbox", "relocate_goal", "path_post", "replan_host", "paths_compact", "plan_host_csr", "paths_jump_points", "jump_points_host"]
