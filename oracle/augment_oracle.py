"""CPU ORACLE of the reference's in-graph training augmentation (augmentation.py:12-77)  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, with the random draws made explicit (so the CUDA kernels can be compared on identical parameters):
  horizontal_flip  augmentation.py:12-29   flip_left_right of image and heat maps + left/right joint channel swap (:20)
  brightness       :68  tf.image.random_brightness  -> x + delta                                            [TF1]
  contrast         :69  tf.image.random_contrast    -> (x - mean_hw) * factor + mean_hw, per channel        [TF1]
  clip             :70  clip_by_value(0, 1)
  random_rotation  :32-36  tf.contrib.image.rotate(BILINEAR): projective transform about the image centre,
                           output (x, y) samples input (cos x - sin y + x_off, sin x + cos y + y_off), zero outside  [TF1]
  random_crop      :39-55  tf.image.crop_and_resize(box = [rh, rw, rh+0.95, rw+0.95]) bilinear, extrapolation 0   [TF1]
                           then hm ** 1.6, + 1e-5, normalise over H, W
PARITY PINNING STATUS: unpinned against TensorFlow (not installable here); [TF1] = library semantics restated from the TF 1.x
sources' documented behaviour."""
import math

import numpy as np
import torch

FLIP_PERM_10 = [3, 4, 5, 0, 1, 2, 7, 6, 8, 9]        # augmentation.py:20 for the 10 channels of joint_names (main.py:18)
CROP_SIZE = 0.95                                     # augmentation.py:40


def flip(img, hm, perm=FLIP_PERM_10):
    return torch.flip(img, dims=[1]), torch.flip(hm, dims=[1])[:, :, perm]


def color(img, delta, factor):
    x = img + delta
    mean = x.mean(dim=(0, 1), keepdim=True)
    return torch.clamp((x - mean) * factor + mean, 0.0, 1.0)


def _bilinear_zero(src, sy, sx):
    """src [H,W,C]; sample at float coords (sy, sx) [Ho,Wo]; each of the 4 neighbours contributes 0 when out of bounds."""
    H, W, _ = src.shape
    y0, x0 = torch.floor(sy), torch.floor(sx)
    wy, wx = (sy - y0).unsqueeze(-1), (sx - x0).unsqueeze(-1)
    out = 0
    for dy, wyv in ((0, 1 - wy), (1, wy)):
        for dx, wxv in ((0, 1 - wx), (1, wx)):
            yy, xx = (y0 + dy).long(), (x0 + dx).long()
            ok = ((yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)).unsqueeze(-1)
            v = src[yy.clamp(0, H - 1), xx.clamp(0, W - 1)]
            out = out + torch.where(ok, v, torch.zeros_like(v)) * wyv * wxv
    return out


def rotate(src, angle):
    """[TF1] tf.contrib.image.rotate(src [H,W,C], angle, 'BILINEAR') via angles_to_projective_transforms."""
    H, W, _ = src.shape
    c, s = math.cos(angle), math.sin(angle)
    x_off = ((W - 1) - (c * (W - 1) - s * (H - 1))) / 2.0
    y_off = ((H - 1) - (s * (W - 1) + c * (H - 1))) / 2.0
    ys, xs = torch.meshgrid(torch.arange(H, dtype=src.dtype), torch.arange(W, dtype=src.dtype), indexing='ij')
    sx = c * xs - s * ys + x_off
    sy = s * xs + c * ys + y_off
    return _bilinear_zero(src, sy, sx)


def crop_and_resize(src, y1, x1, y2, x2):
    """[TF1] tf.image.crop_and_resize(src, [[y1,x1,y2,x2]], crop_size = src size), bilinear, extrapolation_value 0."""
    H, W, _ = src.shape
    i = torch.arange(H, dtype=src.dtype)
    j = torch.arange(W, dtype=src.dtype)
    in_y = y1 * (H - 1) + i * ((y2 - y1) * (H - 1) / (H - 1))
    in_x = x1 * (W - 1) + j * ((x2 - x1) * (W - 1) / (W - 1))
    sy, sx = torch.meshgrid(in_y, in_x, indexing='ij')
    inside = ((sy >= 0) & (sy <= H - 1) & (sx >= 0) & (sx <= W - 1)).unsqueeze(-1)
    top, left = torch.floor(sy), torch.floor(sx)
    bot, right = torch.ceil(sy).clamp(max=H - 1), torch.ceil(sx).clamp(max=W - 1)
    wy, wx = (sy - top).unsqueeze(-1), (sx - left).unsqueeze(-1)
    g = lambda yy, xx: src[yy.long().clamp(0, H - 1), xx.long().clamp(0, W - 1)]
    t = g(top, left) + (g(top, right) - g(top, left)) * wx
    b = g(bot, left) + (g(bot, right) - g(bot, left)) * wx
    out = t + (b - t) * wy
    return torch.where(inside, out, torch.zeros_like(out))


def hm_renorm(hm):
    hm = hm ** 1.6 + 10 ** -5
    return hm / hm.sum(dim=(0, 1), keepdim=True)


def augment_each_train(img, hm, prm, perm=FLIP_PERM_10):
    """augmentation.py:64-73 for one example; prm = dict(flip, delta, contrast, angle, rh, rw)."""
    if prm['flip']:
        img, hm = flip(img, hm, perm)
    img = color(img, prm['delta'], prm['contrast'])
    img, hm = rotate(img, prm['angle']), rotate(hm, prm['angle'])
    box = (prm['rh'], prm['rw'], prm['rh'] + CROP_SIZE, prm['rw'] + CROP_SIZE)
    img, hm = crop_and_resize(img, *box), crop_and_resize(hm, *box)
    return img, hm_renorm(hm)


def draw_params(B, rng):
    """The random draws of augmentation.py:23,33,44-45,68-69 (p > 0.5 flips; angle U(-pi/9, pi/9); rh, rw U(0, 0.05))."""
    return [dict(flip=bool(rng.random() > 0.5), delta=float(rng.uniform(-32 / 255, 32 / 255)), contrast=float(rng.uniform(0.8, 1.2)),
                 angle=float(rng.uniform(-math.pi / 9, math.pi / 9)), rh=float(rng.uniform(0, 1 - CROP_SIZE)),
                 rw=float(rng.uniform(0, 1 - CROP_SIZE))) for _ in range(B)]
