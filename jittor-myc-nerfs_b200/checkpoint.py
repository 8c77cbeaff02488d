"""Checkpoint I/O with the reference's on-disk layout.

TensorBase.get_kwargs / save / load (tensorf-myc/models/tensorBase.py:229-271), NerfPlusPlus.get_kwargs
(models/nerfplusplus.py:165-171) and the driver's `jt.save` / `jt.load` calls (train.py:43-55, 148-163, 344, 360):

    ckpt = {'kwargs': get_kwargs(), 'state_dict': {name: ndarray}, **global_kwargs,
            'alphaMask.shape': (1, 1, D, H, W), 'alphaMask.mask': np.packbits(volume.reshape(-1)), 'alphaMask.aabb': aabb}

The file container is Jittor's, which cannot be exercised here (Jittor is absent from the image; assumption A12, recalled
from jittor/__init__.py `save` / `safepickle` / `safeunpickle`): every Var is replaced by its numpy array, the dict is
pickled with the highest protocol, and `sha1(payload) + b"HCAJSLHD"` is appended; the loader verifies the trailer when it is
present.  `load_checkpoint` also accepts a bare pickle, so files written by either side open on the other provided the
trailer assumption holds.  Pure host code: no arithmetic, nothing on the hot path.
"""
from __future__ import annotations

import hashlib
import pickle

import numpy as np
import torch

_MAGIC = b"HCAJSLHD"


def _to_numpy(x):
    """jt.save's dfs: containers are walked, tensors become numpy arrays."""
    if isinstance(x, dict):
        return {k: _to_numpy(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_numpy(v) for v in x)
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return x


def save_checkpoint(obj, path):
    """jt.save(obj, path) for the pickle container (.th / .pkl, the extensions train.py uses)."""
    payload = pickle.dumps(_to_numpy(obj), pickle.HIGHEST_PROTOCOL)
    with open(path, "wb") as f:
        f.write(payload + hashlib.sha1(payload).digest() + _MAGIC)


def load_checkpoint(path):
    """jt.load(path): returns the dict with numpy leaves; raises ValueError on a damaged trailer."""
    with open(path, "rb") as f:
        data = f.read()
    if data.endswith(_MAGIC):
        payload, digest = data[:-len(_MAGIC) - 20], data[-len(_MAGIC) - 20:-len(_MAGIC)]
        if hashlib.sha1(payload).digest() != digest:
            raise ValueError(f"{path}: checksum mismatch, the checkpoint is damaged")
        data = payload
    return pickle.loads(data)


def pack_alpha_volume(volume):
    """alphaMask.* entries of a checkpoint (tensorBase.py:258-262): np.packbits of the bool volume, big-endian bit order."""
    v = np.asarray(_to_numpy(volume)) > 0
    return {"alphaMask.shape": v.shape, "alphaMask.mask": np.packbits(v.reshape(-1))}


def unpack_alpha_volume(ckpt):
    """tensorBase.py:265-267."""
    shape = tuple(int(s) for s in ckpt["alphaMask.shape"])
    length = int(np.prod(shape))
    return np.unpackbits(np.asarray(ckpt["alphaMask.mask"], dtype=np.uint8))[:length].reshape(shape)


class CheckpointMixin:
    """get_kwargs / save / load of TensorBase with the reference's key names."""

    def get_kwargs(self):
        kw = {
            'aabb': self.aabb,
            'gridSize': [int(g) for g in self.gridSize.tolist()],
            'density_n_comp': self.density_n_comp,
            'appearance_n_comp': self.app_n_comp,
            'app_dim': self.app_dim,
            'density_shift': self.density_shift,
            'alphaMask_thres': self.alphaMask_thres,
            'distance_scale': self.distance_scale,
            'rayMarch_weight_thres': self.rayMarch_weight_thres,
            'fea2denseAct': self.fea2denseAct,
            'near_far': self.near_far,
            'step_ratio': self.step_ratio,
            'shadingMode': self.shadingMode,
            'pos_pe': self.pos_pe,
            'view_pe': self.view_pe,
            'fea_pe': self.fea_pe,
            'featureC': self.featureC,
        }
        if hasattr(self, "bg_net"):          # NerfPlusPlus.get_kwargs (nerfplusplus.py:165-171)
            kw.update(bg_freq=self.bg_freq, bg_view_freq=self.bg_view_freq, bg_D=self.bg_D, radii=self.radii)
        return kw

    def save(self, path, global_kwargs=None):
        ckpt = {'kwargs': self.get_kwargs(), 'state_dict': dict(self.state_dict())}
        if global_kwargs is not None:
            ckpt.update(global_kwargs)
        if self.alphaMask is not None:
            ckpt.update(pack_alpha_volume(self.alphaMask.alpha_volume))
            ckpt.update({'alphaMask.aabb': self.alphaMask.aabb})
        save_checkpoint(ckpt, path)

    def load(self, ckpt):
        from .tensorf import AlphaGridMask
        if 'alphaMask.aabb' in ckpt.keys():
            vol = torch.from_numpy(unpack_alpha_volume(ckpt).astype(np.float32))
            self.alphaMask = AlphaGridMask(self.device, ckpt['alphaMask.aabb'], vol)
        sd = {k: torch.as_tensor(np.asarray(v)) for k, v in ckpt['state_dict'].items()}
        own = self.state_dict()
        for k, v in sd.items():
            if k in own and tuple(own[k].shape) != tuple(v.shape):
                raise ValueError(f"load parameter {k} failed: expect the shape {tuple(own[k].shape)}, but got {tuple(v.shape)}")
        missing, unexpected = self.load_state_dict(sd, strict=False)
        if missing or unexpected:
            print(f"load: {len(missing)} parameters missing from the checkpoint {list(missing)[:4]}, "
                  f"{len(unexpected)} not used {list(unexpected)[:4]}")
        self._invalidate_packed()
