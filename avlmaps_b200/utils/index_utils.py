"""Index helpers with the reference's names (reference avlmaps/utils/index_utils.py)."""
from __future__ import annotations

import numpy as np

from .clip_utils import landmark_text_feats


def get_dynamic_obstacles_map_3d(clip_model, obstacles_cropped, potential_obstacle_classes, obstacle_classes, grid_feat,
                                 grid_pos, rmin, cmin, clip_feat_dim, use_multiple_templates=True, avg_mode=0, vis=False):
    """Reference index_utils.py:138-184: score the potential obstacle classes, per-voxel argmax, union of
    the voxels assigned to `obstacle_classes`, scattered to the cropped top-down grid.  `grid_feat` may be
    the numpy map (uploaded for the call) or the engine.DeviceMap already resident in HBM; the argmax is
    the fused tcgen05 pass (no (N, C) score matrix), everything after it is the reference's numpy."""
    from ..engine import DeviceMap

    all_obstacles_mask = obstacles_cropped == 0
    text_feats, _, n_tmp = landmark_text_feats(clip_model, potential_obstacle_classes, clip_feat_dim,
                                               use_multiple_templates, avg_mode, add_other=True)
    own = not isinstance(grid_feat, DeviceMap)
    dmap = DeviceMap(np.asarray(grid_feat).reshape((-1, np.asarray(grid_feat).shape[-1]))) if own else grid_feat
    try:
        if use_multiple_templates and avg_mode == 1:  # scores are averaged over templates before the argmax
            sc = dmap.scores(text_feats)
            sc = np.mean(sc.reshape((sc.shape[0], -1, n_tmp)), axis=2)
            predict = np.argmax(sc, axis=1)
        else:
            predict = dmap.argmax(text_feats)
    finally:
        if own:
            dmap.close()
    obs_inds = []
    for obs_name in obstacle_classes:
        for i, po_obs_name in enumerate(potential_obstacle_classes):
            if obs_name == po_obs_name:
                obs_inds.append(i)
    pts_mask = np.zeros_like(predict, dtype=bool)
    for id in obs_inds:
        pts_mask = np.logical_or(pts_mask, predict == id)
    new_obstacles = np.zeros_like(obstacles_cropped, dtype=bool)
    obs_pts = grid_pos[pts_mask]
    new_obstacles[obs_pts[:, 0] - rmin, obs_pts[:, 1] - cmin] = 1
    new_obstacles = np.logical_and(new_obstacles, all_obstacles_mask)
    return np.logical_not(new_obstacles)
