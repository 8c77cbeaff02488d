"""Driver-side helpers the reference's training loop uses around the model (tensorf-myc/utils.py:56-62, train.py:25-37):
same names, same arithmetic, host only."""
from __future__ import annotations

import numpy as np
import torch


def N_to_reso(n_voxels, bbox):
    """utils.py:56-59: grid resolution with ~n_voxels cubic voxels inside bbox (aabb [2,3])."""
    bbox = torch.as_tensor(np.asarray(bbox.detach().cpu() if torch.is_tensor(bbox) else bbox), dtype=torch.float32)
    xyz_min, xyz_max = bbox[0], bbox[1]
    voxel_size = ((xyz_max - xyz_min).prod() / n_voxels).pow(1 / 3)
    return ((xyz_max - xyz_min) / voxel_size).long().tolist()


def cal_n_samples(reso, step_ratio=0.5):
    """utils.py:61-62."""
    return int(np.linalg.norm(reso) / step_ratio)


class SimpleSampler:
    """train.py:25-37: a fresh permutation whenever the current one cannot serve another full batch."""

    def __init__(self, total, batch, seed=None):
        self.total, self.batch = total, batch
        self.curr = total
        self.ids = None
        self.rng = np.random.default_rng(seed)

    def nextids(self):
        self.curr += self.batch
        if self.curr + self.batch > self.total:
            self.ids = torch.from_numpy(self.rng.permutation(self.total).astype(np.int64))
            self.curr = 0
        return self.ids[self.curr:self.curr + self.batch]
