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


def compute_ats_bounding_boxes(boxes1, boxes2, return_iou=False):
    """helper.py:33-72 (+ compute_iou, :79-83): IoU-thresholded average threat score of two box sets [N,2,4] (metres;
    rows x / y; columns fl, fr, bl, br) -> 0-dim float32 tensor.  The reference loops over the N1 x N2 pairs in Python and
    builds two shapely polygons per pair on the host; here one kernel does the hulls, the clipping and the score."""
    from .. import _lib
    b1 = boxes1.detach().to(dtype=torch.float32).contiguous()
    b2 = boxes2.detach().to(dtype=torch.float32).contiguous()
    if not (b1.is_cuda and b2.is_cuda):
        raise RuntimeError("driving-dirty_b200 runs on CUDA (sm_100a) only; compute_ats_bounding_boxes got a CPU tensor. "
                           "There is no CPU fallback on this path.")
    if b1.dim() != 3 or b2.dim() != 3 or tuple(b1.shape[1:]) != (2, 4) or tuple(b2.shape[1:]) != (2, 4):
        raise RuntimeError(f"expected boxes of shape [N,2,4], got {tuple(boxes1.shape)} and {tuple(boxes2.shape)}")
    n1, n2 = b1.shape[0], b2.shape[0]
    out = torch.empty(1, dtype=torch.float32, device=b1.device)
    iou = torch.empty(n1, n2, dtype=torch.float32, device=b1.device) if return_iou else None
    _lib.call("dd_ats_bounding_boxes", b1.data_ptr(), n1, b2.data_ptr(), n2, iou.data_ptr() if return_iou else None,
              out.data_ptr(), _lib.stream_ptr())
    return (out[0], iou) if return_iou else out[0]
