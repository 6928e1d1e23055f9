"""Metric / batching helpers with the reference's names (src/utils/helper.py)."""
import torch

from .. import ops


def collate_fn(batch):
    """helper.py:22-23."""
    return tuple(zip(*batch))


def compute_ts_road_map(road_map1, road_map2):
    """helper.py:74-77: tp / (sum1 + sum2 - tp), pooled over everything passed; 0-dim tensor.
    One fused reduction pass on the GPU instead of three."""
    if not torch.is_tensor(road_map1):
        road_map1 = torch.stack(tuple(road_map1), dim=0)
    if not torch.is_tensor(road_map2):
        road_map2 = torch.stack(tuple(road_map2), dim=0)
    return ops.threat_score(road_map1, road_map2)


def convert_map_to_road_map(ego_map):
    """helper.py:17-20 (host-side label preparation; plain torch, not on the GPU hot path)."""
    mask = (ego_map[0, :, :] == 1) * (ego_map[1, :, :] == 1) * (ego_map[2, :, :] == 1)
    return ~mask
