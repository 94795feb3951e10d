"""Seeded input cases shared by the parity tests and tests/golden/make_golden.py."""
import numpy as np
import torch

from pointnet2.scannet_b200 import synthetic as syn

LEVELS = [(8192, 1024, 0.1), (1024, 256, 0.2), (256, 64, 0.4), (64, 16, 0.8)]  # SSG: N, npoint, r


def chunk_xyz(batch, config_id=2, first=0, npoints=8192):
    x, _, _ = syn.make_batch(config_id, batch, npoints, first_chunk=first)
    return np.ascontiguousarray(x[:, :, :3]), x


def adversarial():
    """name -> (1, n, 3) float32 clouds exercising ties / near-origin skips / tiny n."""
    return {k: v[None].copy() for k, v in syn.adversarial_clouds().items()}


def t(a, device="cpu"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def gather_xyz(xyz, idx):
    """numpy: xyz (B, N, 3), idx (B, m) -> (B, m, 3)"""
    return np.take_along_axis(xyz, idx[..., None].astype(np.int64), axis=1)


def features(rng, B, C, N):
    return rng.standard_normal((B, C, N)).astype(np.float32)


def fp_weights(dist2):
    """numpy restatement of pointnet2_modules.py:395-397 on squared distances."""
    d = np.sqrt(dist2)
    r = (1.0 / (d + np.float32(1e-8))).astype(np.float32)
    return (r / r.sum(-1, keepdims=True)).astype(np.float32)
