"""Drop-in replacement for HiP-AD's ``projects/mmdet3d_plugin/ops`` package.

Exports exactly the names the reference's model code imports from it
(``blocks.py:21``, ``sparse_detector.py:18``, ``ego/instance_bank.py:9``, ``plan/instance_bank.py:8``):
``deformable_aggregation_function``, ``feature_maps_format``, ``DeformableAggregationFunction``
(+ the ``DeformableAggregationFunctionA800`` alias), and adds the fused inference entry point.
"""
import torch

from .deformable_aggregation import (  # noqa: F401
    DeformableAggregationFunction,
    DeformableAggregationFunctionA800,
    DeformableAggregationGroupFunction,
    FeatureMapsFormatFunction,
    KernelTimer,
    aggregation_weights,
    deformable_aggregation_group,
    format_feature_levels,
    fused_deformable_aggregation,
    sample_indices,
    share_feature_gradient,
)


def deformable_aggregation_function(feature_maps, spatial_shape, scale_start_index, sampling_location, weights):
    """Same call as ``ops/__init__.py:7-30`` of the reference, minus the GPU-name dispatch
    (one sm_100a build; the reference picks between two builds of identical sources)."""
    return DeformableAggregationFunction.apply(
        feature_maps, spatial_shape, scale_start_index, sampling_location, weights)


def _attach_host_tables(spatial_shape, scale_start_index, shape_list, start_list):
    # Host copies ride along on the tensor objects so that neither the op nor the inverse
    # format needs a device->host sync (the reference pays >= 7 syncs per inverse call).
    spatial_shape._hipad_host = shape_list
    scale_start_index._hipad_host = start_list
    spatial_shape._hipad_i32 = spatial_shape.int()
    scale_start_index._hipad_i32 = scale_start_index.int()


def feature_maps_format(feature_maps, inverse=False):
    """``ops/__init__.py:33-103`` of the reference.

    forward : list over levels of [bs, cams, C, H_l, W_l]  ->  [col_feats, spatial_shape, scale_start_index]
              col_feats [bs, cams*sum(H_l*W_l), C] (camera-major, level, row-major pixels, channels last),
              spatial_shape int64 [cams, L, 2], scale_start_index int64 [cams, L] (absolute rows).
              A nested list (camera groups of different resolution) is concatenated along cameras.
    inverse : the triple -> list over camera groups of lists over levels of [bs, cams_g, C, H_l, W_l] views.
    """
    if inverse:
        col_feats, spatial_shape, scale_start_index = feature_maps
        shapes = getattr(spatial_shape, "_hipad_host", None)
        if shapes is None:
            shapes = spatial_shape.cpu().tolist()          # reference behaviour (syncs)
        num_cams = len(shapes)
        # group consecutive cameras that share all level shapes
        groups = []
        for cam in range(num_cams):
            if groups and shapes[cam] == shapes[groups[-1][0]]:
                groups[-1].append(cam)
            else:
                groups.append([cam])
        out, row = [], 0
        for cams_g in groups:
            level_sizes = [h * w for h, w in shapes[cams_g[0]]]
            per_cam = sum(level_sizes)
            block = col_feats[:, row:row + per_cam * len(cams_g)].unflatten(1, (len(cams_g), per_cam))
            row += per_cam * len(cams_g)
            levels = []
            for (h, w), piece in zip(shapes[cams_g[0]], block.split(level_sizes, dim=2)):
                levels.append(piece.unflatten(2, (h, w)).permute(0, 1, 4, 2, 3))
            out.append(levels)
        return out

    if isinstance(feature_maps[0], (list, tuple)):
        formatted = [feature_maps_format(group) for group in feature_maps]
        col_feats = torch.cat([f[0] for f in formatted], dim=1)
        shape_list, start_list, row = [], [], 0
        for f in formatted:
            shape_list += f[1]._hipad_host
            # rows of later groups start after all rows of earlier groups (the reference leaves
            # them un-offset, ops/__init__.py:67-72, which makes later groups alias group 0)
            start_list += [[s + row for s in cam] for cam in f[2]._hipad_host]
            row += f[0].shape[1]
        spatial_shape = torch.tensor(shape_list, dtype=torch.int64, device=col_feats.device)
        scale_start_index = torch.tensor(start_list, dtype=torch.int64, device=col_feats.device)
        _attach_host_tables(spatial_shape, scale_start_index, shape_list, start_list)
        return [col_feats, spatial_shape, scale_start_index]

    bs, num_cams = feature_maps[0].shape[:2]
    level_hw = [[int(f.shape[-2]), int(f.shape[-1])] for f in feature_maps]
    f0 = feature_maps[0]
    if (f0.is_cuda and f0.dtype in (torch.float32, torch.bfloat16) and len(feature_maps) <= 8 and bs * num_cams <= 65535
            and all(f.is_cuda and f.dtype == f0.dtype and f.dim() == 5 for f in feature_maps)):
        # one transposing kernel (hipad_dfa_format_features) instead of the reference's two full copies
        col_feats = format_feature_levels(list(feature_maps))
    else:   # host tensors / exotic dtypes: the reference's own layout ops (pure data movement, no arithmetic)
        col_feats = torch.cat(
            [f.reshape(bs, num_cams, f.shape[2], -1) for f in feature_maps], dim=-1
        ).permute(0, 1, 3, 2).flatten(1, 2)
    shape_list = [[list(hw) for hw in level_hw] for _ in range(num_cams)]
    start_list, row = [], 0
    for _ in range(num_cams):
        cam_starts = []
        for h, w in level_hw:
            cam_starts.append(row)
            row += h * w
        start_list.append(cam_starts)
    spatial_shape = torch.tensor(shape_list, dtype=torch.int64, device=col_feats.device)
    scale_start_index = torch.tensor(start_list, dtype=torch.int64, device=col_feats.device)
    _attach_host_tables(spatial_shape, scale_start_index, shape_list, start_list)
    return [col_feats, spatial_shape, scale_start_index]
