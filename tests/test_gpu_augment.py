"""-m gpu: the training augmentation kernels against the CPU restatement of augmentation.py on identical random draws.
Tolerance 2e-4 absolute on images in [0,1] / relative on the re-normalised heat maps: the sampling coordinates are evaluated in
fp32 on the GPU and fp64 in the oracle (bilinear interpolation is continuous in the coordinates, so there is no tie problem)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import augment_oracle as ao

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def jcm(built_lib):
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device (no CPU fallback exists)')
    import jcm as _jcm
    _jcm.lib()
    return _jcm


@pytest.mark.parametrize('shape', [(3, 48, 72, 6, 9), (2, 480, 720, 60, 90), (4, 33, 47, 7, 5)])
def test_augment_train_matches_oracle(jcm, shape):
    B, H, W, h, w = shape
    g = torch.Generator().manual_seed(31)
    img = torch.rand(B, H, W, 3, generator=g)
    hm = torch.rand(B, h, w, 10, generator=g) ** 4
    params = jcm.augment.draw_params(B, g)
    params[0, 0] = 1.0                      # at least one flipped and one unflipped example
    params[1, 0] = 0.0
    a, b = jcm.augment.augment_train(img.cuda(), hm.cuda(), params=params)
    for n in range(B):
        prm = dict(flip=bool(params[n, 0] > 0.5), delta=float(params[n, 1]), contrast=float(params[n, 2]), angle=float(params[n, 3]),
                   rh=float(params[n, 4]), rw=float(params[n, 5]))
        ri, rh_ = ao.augment_each_train(img[n].double(), hm[n].double(), prm)
        assert float((a[n].cpu().double() - ri).abs().max()) < 2e-4
        assert float((b[n].cpu().double() - rh_).abs().max()) < 2e-4 * float(rh_.max())
        assert abs(float(b[n].sum((0, 1))[0]) - 1.0) < 1e-4


def test_augment_identity_parameters_and_flip_permutation(jcm):
    g = torch.Generator().manual_seed(32)
    img = torch.rand(2, 40, 56, 3, generator=g)
    hm = torch.rand(2, 5, 7, 8, generator=g)
    names = jcm.JOINT_NAMES[:7] + ['torso']
    perm = jcm.augment.flip_permutation(names)
    assert perm == [3, 4, 5, 0, 1, 2, 6, 7]           # lsho<->rsho, lelb<->relb, lwri<->rwri; lhip (its partner is absent) and torso stay
    assert jcm.augment.flip_permutation(jcm.JOINT_NAMES) == ao.FLIP_PERM_10
    ident = torch.tensor([[0, 0, 1, 0, 0, 0], [1, 0, 1, 0, 0, 0]], dtype=torch.float32)
    # crop of relative size 0.95 still resamples, so compare against the oracle's crop of the (flipped) input
    a, b = jcm.augment.augment_train(img.cuda(), hm.cuda(), params=ident, perm=perm)
    for n in range(2):
        src_i, src_h = (img[n], hm[n]) if n == 0 else (torch.flip(img[n], dims=[1]), torch.flip(hm[n], dims=[1])[:, :, perm])
        ri = ao.crop_and_resize(src_i.double(), 0.0, 0.0, ao.CROP_SIZE, ao.CROP_SIZE)
        assert float((a[n].cpu().double() - ri).abs().max()) < 1e-5
        rh_ = ao.hm_renorm(ao.crop_and_resize(src_h.double(), 0.0, 0.0, ao.CROP_SIZE, ao.CROP_SIZE))
        assert float((b[n].cpu().double() - rh_).abs().max()) < 1e-5 * float(rh_.max())
    with pytest.raises(ValueError):
        jcm.augment.augment_train(img.cuda(), hm.cuda(), params=ident)        # 8 channels need an explicit permutation
