"""Restatement of the reference's pure-torch aggregation path.  TEST INFRASTRUCTURE ONLY.

This is the reference's own CPU-runnable implementation of the hot path
(``use_deformable_func=False``): it is the second parity oracle inside the
valid region (SURVEY.md §8 note N1) and the thing ``bench.py --impl reference``
and ``cpu_baseline`` time on the host cores.

Follows /root/reference/projects/mmdet3d_plugin/models/blocks.py
  :217-225  project_points           -> project_points()
  :228-251  feature_sampling         -> sample_levels()
  :253-264  multi_view_level_fusion  -> fuse()
  :170      .sum(dim=2) over points  -> grid_sample_path()
Pinned against tests/golden/*.npz (outputs of the unmodified blocks.py).
"""
import torch
import torch.nn.functional as F


def project_points(key_points, projection_mat, image_wh=None):
    """[bs,A,P,3] x [bs,cams,4,4] -> [bs,cams,A,P,2] (x,y), normalised by image_wh."""
    ones = torch.ones_like(key_points[..., :1])
    hom = torch.cat([key_points, ones], dim=-1)                       # blocks.py:220
    # same batched-matmul formulation as the reference so the fp32 rounding matches
    cam = torch.matmul(projection_mat[:, :, None, None], hom[:, None, ..., None]).squeeze(-1)
    xy = cam[..., :2] / torch.clamp(cam[..., 2:3], min=1e-5)          # blocks.py:222
    if image_wh is not None:
        xy = xy / image_wh[:, :, None, None]
    return xy


def sample_levels(feature_maps, points_2d):
    """feature_maps: L x [bs,cams,C,H,W]; points_2d [bs,cams,A,P,2] in (0,1).
    Returns [bs,A,cams,L,P,C] (blocks.py:239-251)."""
    bs, cams, A, P, _ = points_2d.shape
    grid = (points_2d * 2 - 1).flatten(end_dim=1)                     # [bs*cams,A,P,2]
    per_level = [
        F.grid_sample(fm.flatten(end_dim=1), grid, mode="bilinear",
                      padding_mode="zeros", align_corners=False)      # [bs*cams,C,A,P]
        for fm in feature_maps
    ]
    x = torch.stack(per_level, dim=1)                                 # [bs*cams,L,C,A,P]
    L = len(feature_maps)
    return x.reshape(bs, cams, L, -1, A, P).permute(0, 4, 1, 2, 5, 3)


def fuse(sampled, weights, num_groups):
    """sampled [bs,A,cams,L,P,C], weights [bs,A,cams,L,P,G] -> [bs,A,P,C] (blocks.py:253-264)."""
    bs, A, cams, L, P, C = sampled.shape
    g = sampled.reshape(bs, A, cams, L, P, num_groups, C // num_groups)
    g = weights[..., None] * g
    return g.sum(dim=2).sum(dim=2).reshape(bs, A, P, C)


def grid_sample_path(feature_maps, points_2d, weights):
    """The reference torch path from projected points on: -> [bs,A,C]."""
    G = weights.shape[-1]
    return fuse(sample_levels(feature_maps, points_2d), weights, G).sum(dim=2)


def grid_sample_path_from_keypoints(feature_maps, key_points, projection_mat, image_wh, weights):
    return grid_sample_path(feature_maps, project_points(key_points, projection_mat, image_wh), weights)


# ---- layout helpers between the torch path and the CUDA-op contract -------------------

def to_op_layout(points_2d, weights):
    """[bs,cams,A,P,2],[bs,A,cams,L,P,G] -> loc [bs,A,P,cams,2], w [bs,A,P,cams,L,G]
    (the permutes of blocks.py:144-158)."""
    loc = points_2d.permute(0, 2, 3, 1, 4).contiguous()
    w = weights.permute(0, 1, 4, 2, 3, 5).contiguous()
    return loc, w


def flatten_feature_maps(feature_maps):
    """L x [bs,cams,C,H,W] -> (col_feats [bs,cams*sum(HW),C], shapes [cams,L,2], starts [cams,L])
    int64 tables on the feature device (ops/__init__.py:74-103)."""
    bs, cams = feature_maps[0].shape[:2]
    cols = [fm.reshape(bs, cams, fm.shape[2], -1) for fm in feature_maps]
    col = torch.cat(cols, dim=-1).permute(0, 1, 3, 2).flatten(1, 2)
    hw = torch.tensor([[list(fm.shape[-2:]) for fm in feature_maps]] * cams,
                      dtype=torch.int64, device=col.device)
    sizes = (hw[..., 0] * hw[..., 1]).flatten()
    starts = torch.cumsum(sizes, 0) - sizes
    return col, hw, starts.reshape(cams, -1)
