"""Multi-GPU partitioning of the RoI-pooling path: by image, no data-path
collective.

A RoI reads only its own image's pyramid (batch index, roi_align_2d.py:111,136)
and an image's feature gradient depends only on that image's RoIs (:213), so
image n goes to rank n mod world_size and every rank pools its own images'
RoIs.  torch.distributed is used only to agree on timings (max over ranks) --
the reference's multi-GPU mode is likewise one process per device
(train.py:117-121).
"""
import numpy as np
import torch
import torch.distributed as dist


def images_of_rank(n_images, world_size, rank):
    """Indices of the images owned by `rank` (round-robin)."""
    return list(range(rank, n_images, world_size))


def shard_rois(indices_and_rois, n_images, world_size, rank):
    """Rows of (R,5) RoIs that belong to this rank's images, with the batch index
    renumbered to the rank-local image order.  Returns (local_rois, global_rows)."""
    rois = np.asarray(indices_and_rois)
    mine = images_of_rank(n_images, world_size, rank)
    remap = {g: i for i, g in enumerate(mine)}
    img = rois[:, 0].astype(np.int64)
    rows = np.nonzero(np.isin(img, mine))[0]
    local = rois[rows].copy()
    local[:, 0] = [remap[int(g)] for g in img[rows]]
    return local, rows


def max_over_ranks(value, device=None):
    """max of a Python float over all ranks (identity without a process group)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
