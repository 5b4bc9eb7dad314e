"""CPU checks (fp64, small maps) of the three identities the tensor-core spatial model is built on (DESIGN 4.5b;
joint-cnn-mrf_b200/jcm/csrc/spatial_model.cu, "smt" section), each against the oracle's own conv_mrf (reference main.py:77-91) or
its autograd - so the algebra of the CUDA kernels is pinned where no GPU is needed:

  1. centred prior:   conv_mrf(P, L) == conv_mrf(P - c0, L) + c0 * sum(L)            (every output of the 'valid' convolution sums
                      over the whole likelihood map and the resize weights sum to one)
  2. folded resize:   d(loss)/dL through dC = R_y^T dT R_x equals the contraction of T = R_y^T dT (W columns) with the weights
                      Wd[dy][v][x'] = (1-w(x')) P[dy][x'-v+W-1] + w(x') P[dy][x'-v+W], plus c0 * sum(dT) for a centred prior
  3. prior gradient:  dP[r][c] = sum over the diagonal x - v + W - 1 = c of  blk[r][v][x] = sum_{n,y'} L[n][y'][v] dC[n][y'+r-(H-1)][x]
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import jcm_oracle as orc  # noqa: E402


def _legacy_w(n_out):
    """w(i) of the legacy bilinear resize from n_out + 1 to n_out samples (lo = i, hi = i + 1), float32 as in the kernels"""
    scale = np.float32(n_out + 1) / np.float32(n_out)
    src = np.arange(n_out, dtype=np.float32) * scale
    lo = np.floor(src)
    assert np.array_equal(lo, np.arange(n_out, dtype=np.float32))
    return torch.from_numpy((src - lo).astype(np.float64))


def _resize_matrix(n_out):
    """R [n_out, n_out + 1]: out = R @ in"""
    w = _legacy_w(n_out)
    R = torch.zeros(n_out, n_out + 1, dtype=torch.float64)
    idx = torch.arange(n_out)
    R[idx, idx] = 1 - w
    R[idx, idx + 1] = w
    return R


def _case(seed, B, H, W):
    g = torch.Generator().manual_seed(seed)
    P = 0.1386 + 0.01 * torch.rand(2 * H, 2 * W, generator=g, dtype=torch.float64)
    L = torch.rand(B, H, W, generator=g, dtype=torch.float64) + 0.14
    G = torch.randn(B, H, W, generator=g, dtype=torch.float64)      # d(loss)/d(conv_mrf output)
    return P, L, G


@pytest.mark.parametrize('B,H,W', [(2, 6, 9), (3, 12, 20), (1, 5, 16)])
def test_centred_prior_identity(B, H, W):
    P, L, _ = _case(1, B, H, W)
    ref = orc.conv_mrf(P.view(1, 2 * H, 2 * W, 1), L.unsqueeze(3))[..., 0]
    c0 = P.mean()
    cen = orc.conv_mrf((P - c0).view(1, 2 * H, 2 * W, 1), L.unsqueeze(3))[..., 0] + c0 * L.sum(dim=(1, 2)).view(B, 1, 1)
    assert float((ref - cen).abs().max() / ref.abs().max()) < 1e-13


@pytest.mark.parametrize('B,H,W', [(2, 6, 9), (3, 12, 20), (1, 5, 16)])
def test_folded_resize_and_diagonal_sum_match_autograd(B, H, W):
    P, L, G = _case(2, B, H, W)
    Pv = P.clone().requires_grad_(True)
    Lv = L.clone().requires_grad_(True)
    out = orc.conv_mrf(Pv.view(1, 2 * H, 2 * W, 1), Lv.unsqueeze(3))[..., 0]
    (out * G).sum().backward()
    dL_ref, dP_ref = Lv.grad, Pv.grad

    Ry, Rx = _resize_matrix(H), _resize_matrix(W)
    T = torch.einsum('yu,nyx->nux', Ry, G)                        # R_y^T dT   [B, H+1, W]
    dC = torch.einsum('nux,xc->nuc', T, Rx)                       # ... R_x    [B, H+1, W+1]
    w = _legacy_w(W)

    # 2. dL[n][u][v] = sum_dy sum_x' T[n][u+dy-(H-1)][x'] * ((1-w) Pc[dy][x'-v+W-1] + w Pc[dy][x'-v+W])  +  c0 * sum(dT[n])
    c0 = P.mean()
    Pc = P - c0
    xs = torch.arange(W).view(1, W)       # x'
    vs = torch.arange(W).view(W, 1)       # v
    dL = torch.zeros(B, H, W, dtype=torch.float64)
    for dy in range(2 * H):
        Wd = (1 - w).view(1, W) * Pc[dy][xs - vs + W - 1] + w.view(1, W) * Pc[dy][xs - vs + W]      # [v][x']
        for u in range(H):
            y = u + dy - (H - 1)
            if 0 <= y <= H:
                dL[:, u, :] += T[:, y, :] @ Wd.t()
    dL += c0 * G.sum(dim=(1, 2)).view(B, 1, 1)
    assert float((dL - dL_ref).abs().max() / dL_ref.abs().max()) < 1e-12

    # 3. blk[r][v][x] = sum_{n,y'} L[n][y'][v] * dC[n][y'+r-(H-1)][x];  dP[r][c] = sum_{x-v+W-1=c} blk[r][v][x]
    #    (conv_mrf convolves with the flipped map, Lf[u][v] = L[H-1-u][W-1-v]; substituting y' = H-1-a+y, v = W-1-b+x in
    #    dP[a][b] = sum dC[y][x] Lf[a-y][b-x] gives the correlation with the UNflipped map above, r = a the prior row)
    dP = torch.zeros(2 * H, 2 * W, dtype=torch.float64)
    for r in range(2 * H):
        blk = torch.zeros(W, W + 1, dtype=torch.float64)
        for yp in range(H):
            y = yp + r - (H - 1)
            if 0 <= y <= H:
                blk += torch.einsum('nv,nx->vx', L[:, yp, :], dC[:, y, :])
        for v in range(W):
            for x in range(W + 1):
                dP[r, x - v + W - 1] += blk[v, x]
    assert float((dP - dP_ref).abs().max() / dP_ref.abs().max()) < 1e-12
