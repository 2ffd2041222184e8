"""pointcloud_b200 -- the point-cloud reconstruction-loss hot path of JoongWonSeo/pointcloud
(Chamfer distance and auction EMD, forward + backward) re-built for NVIDIA B200 (sm_100a).

Python host layer (this package) mirrors the reference's loss interface
(pointcloud_vision/utils.py:207-309, pointcloud_vision/loss/emd/emd_module.py); all arithmetic runs in
hand-written CUDA kernels behind the C ABI of include/pcl.h (libpcl_b200.so).  No CPU fallback.
"""
from . import cfg
from .chamfer import chamfer_distance, chamfer_forward_raw
from .emd_module import emdFunction, emdModule, emd_forward_raw, set_emd_path
from .losses import (ChamferDistance, EarthMoverDistance, FilterClasses, FilteringChamferDistance,
                     SegmentingChamferDistance, chamfer_emd_loss)
from .sampling import farthest_point_sample, query_ball_point, sample_farthest_points
from .sharded import ShardedChamferEmdStep, ShardedLoss, shard_bounds

__all__ = [
    "cfg", "chamfer_distance", "chamfer_forward_raw", "emdFunction", "emdModule", "emd_forward_raw", "set_emd_path",
    "ChamferDistance", "FilteringChamferDistance", "SegmentingChamferDistance", "EarthMoverDistance",
    "FilterClasses", "chamfer_emd_loss", "ShardedLoss", "ShardedChamferEmdStep", "shard_bounds",
    "farthest_point_sample", "sample_farthest_points", "query_ball_point",
]
