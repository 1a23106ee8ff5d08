"""Host-side mirror of openpoints.models.layers.{subsample,group,upsampling} (layers/__init__.py:10-12)."""
from .subsample import furthest_point_sample, gather_operation, random_sample, fps, FurthestPointSampling
from .group import (grouping_operation, torch_grouping_operation, ball_query, QueryAndGroup, GroupAll,
                    KNNGroup, create_grouper, GroupingOperation, BallQuery)
from .upsampling import three_nn, three_interpolate, three_interpolation, ThreeNN, ThreeInterpolate
