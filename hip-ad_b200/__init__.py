"""hip-ad_b200 — B200-native deformable feature aggregation for HiP-AD (drop-in for
``projects/mmdet3d_plugin/ops`` and the ``DeformableFeatureAggregation`` module).

Import name: ``hipad_b200`` (a shim package next to this directory puts it on the path,
since ``hip-ad_b200`` itself is not a valid Python identifier).
"""
from .ops import (  # noqa: F401
    DeformableAggregationFunction,
    DeformableAggregationFunctionA800,
    deformable_aggregation_function,
    deformable_aggregation_group,
    feature_maps_format,
    format_feature_levels,
    fused_deformable_aggregation,
    sample_indices,
    share_feature_gradient,
)
from .blocks import (  # noqa: F401
    DeformableFeatureAggregation,
    aggregate_layer,
    SparseBox3DKeyPointsGenerator,
    SparsePoint3DKeyPointsGenerator,
)

__version__ = "0.1.0"
