"""Regenerator of the reference's `pairwise_distribution.pickle` (TEST INFRASTRUCTURE + offline tool).

The pickle shipped with the reference is corrupted (every byte >= 0x80 was replaced by U+FFFD), so the
only source of the reference's pairwise initialisation is to re-derive it from `data_FLIC.mat`.

Follows, step by step:
  * label format ............ reference `data.py:97-114,165-189` (3x3 binomial blob, /8 down-scaling, pad 5)
  * histogram + smoothing ... reference `prepare_pairwise_distribution.py:13-14,29-48` (9x9 binomial, 'same')
  * key order / naming ...... reference `prepare_pairwise_distribution.py:16,51-56`  ('<joint>_<cond>')

Two variants:
  * "shipped": no `flip_backward_poses`, torso = centre of the FLIC `torsobox` field. This is the recipe whose
    histogram supports match the surviving structure of the corrupt shipped pickle for 90/90 pairs
    (see `zero_runs_from_corrupt_pickle` / tests/test_pairwise_prior.py).
  * "scripts": the code as it stands today (`data.py:35-49` flip incl. its numpy view-aliasing behaviour,
    torso = mean of lsho, rhip, rsho, lhip, `data.py:166-168`).

Nothing here is imported by the product path.
"""
import re

import numpy as np
from scipy import signal
from scipy.io import loadmat

JOINT_IDS = ['lsho', 'lelb', 'lwri', 'rsho', 'relb', 'rwri', 'lhip', 'rhip', 'nose', 'torso']
# reference data.py:100-104 (column of `coords` for every joint; 'torso' is written into column 28)
FLIC_COL = {'lsho': 0, 'lelb': 1, 'lwri': 2, 'rsho': 3, 'relb': 4, 'rwri': 5, 'lhip': 6, 'rhip': 9, 'nose': 16,
            'torso': 28}
ORIG_H, ORIG_W = 480, 720
HM_H, HM_W = 60, 90


def _flip_backward_poses_aliasing(c):
    """reference data.py:35-49. `coords_left_joint`/`coords_right_joint` are numpy *views*, so after
    `c[:, l] = right` the later `c[:, r] = left` copies the already-overwritten column: both end up = right."""
    if c[0, FLIC_COL['lhip']] < c[0, FLIC_COL['rhip']]:
        for l, r in zip(['lwri', 'lelb', 'lhip', 'lsho'], ['rwri', 'relb', 'rhip', 'rsho']):
            l, r = FLIC_COL[l], FLIC_COL[r]
            c[:, l] = c[:, r]
    return c


def joint_positions(mat_path, variant='shipped', split='train'):
    """(row, col) float heat-map coordinates [n, 10, 2] for every example of the split (data.py:170-179)."""
    ex = loadmat(mat_path)['examples'][0]
    want = 1 if split == 'train' else 0
    out = []
    for e in ex:
        if int(e[7][0, 0]) != want:
            continue
        c = np.array(e[2], dtype=np.float64)  # 2 x 29, (x; y) in pixels
        if variant == 'scripts':
            c = _flip_backward_poses_aliasing(c)
            torso = (c[:, 0] + c[:, 9] + c[:, 3] + c[:, 6]) / 4
        elif variant == 'shipped':
            x1, y1, x2, y2 = np.array(e[6][0], dtype=np.float64)
            torso = np.array([(x1 + x2) / 2, (y1 + y2) / 2])
        else:
            raise ValueError(variant)
        c[:, 28] = torso
        pos = []
        for j in JOINT_IDS:
            x, y = c[0, FLIC_COL[j]], c[1, FLIC_COL[j]]
            row = max(min(y, ORIG_H), 0) / 8
            col = max(min(x, ORIG_W), 0) / 8
            pos.append((row, col))
        out.append(pos)
    return np.array(out)


def target_heat_maps(pos):
    """[n, 60, 90, 10] float32 labels exactly as data.py:106-114,180-189 builds them."""
    coefs = np.array([[1, 2, 1]], dtype=np.float32) / 4
    kernel = coefs.T @ coefs
    pad, temp = 5, 1
    n = pos.shape[0]
    y = np.zeros([n, HM_H, HM_W, len(JOINT_IDS)], dtype=np.float32)
    for i in range(n):
        for j in range(len(JOINT_IDS)):
            hm = np.zeros([HM_H + 2 * pad, HM_W + 2 * pad], dtype=np.float32)
            r, c = pos[i, j, 0] + pad, pos[i, j, 1] + pad
            h1, h2 = int(r - temp), int(r + temp + 1)
            w1, w2 = int(c - temp), int(c + temp + 1)
            hm[h1:h2, w1:w2] = kernel
            y[i, :, :, j] = hm[pad:pad + HM_H, pad:pad + HM_W]
    return y


def compute_pairwise_distribution(y_train, joint, cond_j):
    """prepare_pairwise_distribution.py:29-48 (float64 result, 120 x 180, sum 1 before smoothing)."""
    coefs = np.array([[1, 8, 28, 56, 70, 56, 28, 8, 1]], dtype=np.uint16) / 256
    kernel = coefs.T @ coefs
    hh, ww = y_train.shape[1], y_train.shape[2]
    pd = np.zeros([hh * 2, ww * 2])
    a, b = JOINT_IDS.index(joint), JOINT_IDS.index(cond_j)
    for i in range(y_train.shape[0]):
        img_j, img_cj = y_train[i, :, :, a], y_train[i, :, :, b]
        xj, yj = np.where(img_j == np.max(img_j))
        xcj, ycj = np.where(img_cj == np.max(img_cj))
        if len(xj) != len(xcj):  # numpy broadcasting in the reference would fail the same way unless one is 1
            if len(xj) != 1 and len(xcj) != 1:
                n = min(len(xj), len(xcj))
                xj, yj, xcj, ycj = xj[:n], yj[:n], xcj[:n], ycj[:n]
        pd[hh + (xj - xcj), ww + (yj - ycj)] += 1
    pd = pd / np.float32(np.sum(pd))
    return signal.convolve2d(pd, kernel, mode='same', boundary='fill', fillvalue=0)


def regenerate(mat_path, variant='shipped', keys=None):
    """dict '<joint>_<cond>' -> float64 (120, 180), in the reference's insertion order."""
    y = target_heat_maps(joint_positions(mat_path, variant, 'train'))
    out = {}
    for j in JOINT_IDS:
        for c in JOINT_IDS:
            if c != j:
                k = j + '_' + c
                if keys is None or k in keys:
                    out[k] = compute_pairwise_distribution(y, j, c)
    return out


def zero_runs_from_corrupt_pickle(path, min_cells=8):
    """Pin against the (corrupt) shipped pickle: for every key return the list of zero-run lengths (in cells,
    runs >= min_cells) that survive the corruption (0x00 bytes were not touched by the U+FFFD replacement)."""
    raw = open(path, 'rb').read()
    keys = [a + '_' + b for a in JOINT_IDS for b in JOINT_IDS if a != b]
    pos, start = [], 0
    for k in keys:  # protocol 4 emits <len byte> key ... value, in dict insertion order
        p = raw.find(bytes([len(k)]) + k.encode(), start)
        pos.append(p)
        if p >= 0:
            start = p + 1
    out = {}
    for i, k in enumerate(keys):
        if pos[i] < 0:
            continue
        end = pos[i + 1] if i + 1 < len(keys) and pos[i + 1] > 0 else len(raw)
        seg = raw[pos[i]:end]
        runs = [len(m.group(0)) // 8 for m in re.finditer(rb'\x00{%d,}' % (8 * min_cells), seg)]
        out[k] = runs
    return out


def zero_runs_of_array(a, min_cells=8):
    flat = (np.asarray(a).ravel() == 0)
    runs, n = [], 0
    for z in flat:
        if z:
            n += 1
        else:
            if n >= min_cells:
                runs.append(n)
            n = 0
    if n >= min_cells:
        runs.append(n)
    return runs


if __name__ == '__main__':
    import argparse
    import pickle
    ap = argparse.ArgumentParser()
    ap.add_argument('--mat', default='/root/reference/data_FLIC.mat')
    ap.add_argument('--variant', default='shipped')
    ap.add_argument('--out', required=True, help='.npz (compressed) or .pickle')
    a = ap.parse_args()
    d = regenerate(a.mat, a.variant)
    if a.out.endswith('.npz'):
        np.savez_compressed(a.out, **d)
    else:
        with open(a.out, 'wb') as f:
            pickle.dump(d, f, protocol=4)
    print('wrote', a.out, len(d), 'pairs')
