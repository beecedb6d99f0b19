"""Wire / disk formats around the hot path: pure (de)serialisation, host Python like the reference.

  map PNGs        save block global_planner_st.py:365-374 (= ccst:625-634) and the pre-map loader st:176-182; the
                  pixel work runs on the device (fx_grid_to_image / fx_image_to_grid), Pillow only (de)codes the file
  pre-map merge   st:210-224: index bookkeeping here, the two slice assignments on the device (fx_grid_paste)
  OccupancyGrid   planner.occupancy_grid_to_array / array_to_occupancy_grid (fx_grid_decode / fx_grid_encode)
  PointCloud2     cloud.xyz_to_pointcloud2_fields / pointcloud2_to_xyz (plc_point2_st.py:112-138)
  Path            publish_path, st:87-100 -- plain dicts with the message's field names (no ROS in this image)
"""
import numpy as np
import torch

from . import api


def map_png_name(map_o):
    """'%.2f' % x + '%.2f' % y + '_out.png' (global_planner_st.py:373): '-16.40-4.80_out.png' is origin (-16.40, -4.80)."""
    return "%.2f" % map_o[0] + "%.2f" % map_o[1] + "_out.png"


def grid_to_image(mapu, device=0):
    """[x][y] array -> uint8 'L' image array [H][W]: 0 -> 255 (free), non-zero -> 0, then .T[::-1] (st:368-372)."""
    g = torch.from_numpy(np.ascontiguousarray((np.asarray(mapu) != 0).astype(np.uint8))).to("cuda:%d" % device)
    return api.grid_to_image(g).cpu().numpy()


def image_to_grid(img_l, threshold=200, device=0):
    """'L' image array -> uint8 [x][y]: pixel > threshold is free, everything else 1, `img[::-1].T` (st:176-182).
    threshold = 0 is the fixture convention of the saved maps (`a[::-1].T == 0`)."""
    a = torch.from_numpy(np.ascontiguousarray(np.asarray(img_l, dtype=np.uint8))).to("cuda:%d" % device)
    return api.image_to_grid(a, threshold).cpu().numpy()


def save_map_png(mapu, map_o, directory, device=0):
    """Write the planner's map dump; returns the file name (st:365-374: RGB PNG named after the map origin)."""
    import os
    from PIL import Image
    path = os.path.join(directory, map_png_name(map_o))
    Image.fromarray(grid_to_image(mapu, device)).convert("RGB").save(path)
    return path


def load_map_png(path, threshold=200, device=0):
    """Pre-known map before flight (st:176-182): 'L' conversion, x > 200 -> 0 (free) else 1, `img[::-1].T`."""
    from PIL import Image
    return image_to_grid(np.array(Image.open(path).convert("L")), threshold, device)


def merge_premap(mapu, map_o, map_t, map_pre, ori_pre, reso):
    """Merge the pre-known map with the detected one (global_planner_st.py:210-224).

    mapu / map_pre: uint8 CUDA tensors [x][y]; map_o / map_t: world position of the detected map's first cell and of its
    far corner; ori_pre: world position of the pre-map's first cell.  Returns (merged grid, new origin).  Index math
    is the reference's (`astype(int)`, `int()`: truncation); the detected map is pasted second and wins overlaps."""
    l1_pre, l2_pre = map_pre.shape
    map_c, map_r = mapu.shape
    t_pre = [ori_pre[0] + reso * l1_pre, ori_pre[1] + reso * l2_pre]
    map_o1 = [min(map_o[0], ori_pre[0]), min(map_o[1], ori_pre[1])]
    o_idx = ((np.array(map_o) - map_o1) / reso).astype(int)
    p_idx = ((np.array(ori_pre) - map_o1) / reso).astype(int)
    map_c1 = int((max(t_pre[0], map_t[0]) - map_o1[0]) / reso)
    map_r1 = int((max(t_pre[1], map_t[1]) - map_o1[1]) / reso)
    # numpy would raise on a slice that does not fit (shape mismatch); keep that contract
    if p_idx[0] + l1_pre > map_c1 or p_idx[1] + l2_pre > map_r1 or o_idx[0] + map_c > map_c1 or o_idx[1] + map_r > map_r1:
        raise ValueError("could not broadcast input array into the merged map (the reference raises here too)")
    out = torch.zeros((map_c1, map_r1), dtype=torch.uint8, device=mapu.device)
    api.grid_paste(map_pre, out, paste_at=(int(p_idx[0]), int(p_idx[1])))
    api.grid_paste(mapu, out, paste_at=(int(o_idx[0]), int(o_idx[1])))
    return out, map_o1


def path_message(points, stamp=None, frame_id="map"):
    """nav_msgs/Path as publish_path fills it (st:87-100): one PoseStamped per (x, y, z) row, frame 'map'."""
    hdr = {"frame_id": frame_id, "stamp": stamp}
    return {"header": dict(hdr),
            "poses": [{"header": dict(hdr), "pose": {"position": {"x": float(d[0]), "y": float(d[1]), "z": float(d[2])}}}
                      for d in points]}
