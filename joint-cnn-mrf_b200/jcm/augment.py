"""The reference's training augmentation (augmentation.py:58-77 `augment_train`, applied at main.py:495-499) on the GPU kernels of
csrc/augment.cu: horizontal flip with the left/right joint swap, brightness, contrast, clip, rotation by up to +-20 degrees,
95 % crop + resize, and the heat-map sharpening hm**1.6 re-normalisation.  SURVEY 8(f2)."""
import math

import numpy as np
import torch

from ._lib import lib, check
from .ops import _ptr, _stream, _req, F32

CROP_SIZE = 0.95                                    # augmentation.py:40
MAX_ROTATE = math.pi / 9                            # augmentation.py:65
FLIP_PERM_10 = [3, 4, 5, 0, 1, 2, 7, 6, 8, 9]       # augmentation.py:20 (joint_names of main.py:18)


def flip_permutation(joint_names):
    """Channel permutation that swaps left/right joints ('l...' <-> 'r...') - augmentation.py:20 for any joint_names."""
    idx = {n: i for i, n in enumerate(joint_names)}
    perm = []
    for n in joint_names:
        other = ('r' + n[1:]) if n.startswith('l') else ('l' + n[1:]) if n.startswith('r') else n
        perm.append(idx.get(other, idx[n]))
    return perm


def draw_params(B, generator=None):
    """The random draws of augmentation.py:23,33,44-45,68-69 as a [B,6] CPU tensor: flip, delta, contrast, angle, rh, rw."""
    u = torch.rand(B, 6, generator=generator)
    out = torch.empty(B, 6)
    out[:, 0] = (u[:, 0] > 0.5).float()
    out[:, 1] = (u[:, 1] * 2 - 1) * (32.0 / 255.0)
    out[:, 2] = 0.8 + 0.4 * u[:, 2]
    out[:, 3] = (u[:, 3] * 2 - 1) * MAX_ROTATE
    out[:, 4] = u[:, 4] * (1 - CROP_SIZE)
    out[:, 5] = u[:, 5] * (1 - CROP_SIZE)
    return out


def augment_train(img, hm, params=None, perm=None, generator=None):
    """img [B,H,W,3] in [0,1], hm [B,h,w,C] (CUDA fp32) -> augmented (img, hm).  params: [B,6] as draw_params (drawn if None);
    perm: left/right channel permutation for the flip (default: the reference's 10-channel one, or identity-safe for other C)."""
    _req(img, F32, 'img')
    _req(hm, F32, 'hm')
    B, H, W, C = img.shape
    _, h, w, Ch = hm.shape
    if params is None:
        params = draw_params(B, generator)
    params = torch.as_tensor(params, dtype=torch.float32)
    if tuple(params.shape) != (B, 6):
        raise ValueError('params must be [B, 6]: flip, delta, contrast, angle, rh, rw')
    if perm is None:
        if Ch != 10:
            raise ValueError('perm (left/right channel permutation) is required when the heat maps do not have the 10 reference channels')
        perm = FLIP_PERM_10
    if len(perm) != Ch:
        raise ValueError('perm must have one entry per heat-map channel')
    prm = torch.zeros(B, 8, dtype=torch.float32)
    prm[:, 0:3] = params[:, 0:3]
    prm[:, 3] = torch.from_numpy(np.cos(params[:, 3].double().numpy())).float()
    prm[:, 4] = torch.from_numpy(np.sin(params[:, 3].double().numpy())).float()
    prm[:, 5:7] = params[:, 4:6]
    dev = img.device
    prm = prm.to(dev)
    permd = torch.tensor(perm, dtype=torch.int32, device=dev)
    l, st = lib(), _stream()
    mean_ws = torch.empty(B * C, dtype=F32, device=dev)
    a, b = torch.empty_like(img), torch.empty_like(img)
    check(l.jcm_augment_color(_ptr(img), _ptr(prm), _ptr(mean_ws), B, H, W, C, _ptr(a), st), 'jcm_augment_color')
    check(l.jcm_augment_rotate(_ptr(a), _ptr(prm), B, H, W, C, _ptr(b), st), 'jcm_augment_rotate')
    check(l.jcm_augment_crop_resize(_ptr(b), _ptr(prm), B, H, W, C, CROP_SIZE, _ptr(a), st), 'jcm_augment_crop_resize')
    ha, hb = torch.empty_like(hm), torch.empty_like(hm)
    check(l.jcm_augment_flip_channels(_ptr(hm), _ptr(prm), _ptr(permd), B, h, w, Ch, _ptr(ha), st), 'jcm_augment_flip_channels')
    check(l.jcm_augment_rotate(_ptr(ha), _ptr(prm), B, h, w, Ch, _ptr(hb), st), 'jcm_augment_rotate')
    check(l.jcm_augment_crop_resize(_ptr(hb), _ptr(prm), B, h, w, Ch, CROP_SIZE, _ptr(ha), st), 'jcm_augment_crop_resize')
    check(l.jcm_augment_hm_renorm(_ptr(ha), B, h, w, Ch, 1.6, 1e-5, _ptr(hb), st), 'jcm_augment_hm_renorm')
    return a, hb


def augment_test(img, hm):
    """augmentation.py:80-89: identity."""
    return img, hm
