"""On-disk data formats either side of the hot path: what the reference's `data.py` and
`prepare_pairwise_distribution.py` write and `main.py:286-299` reads back.

    x_{train,test}_flic.npy   float32 [n, 480, 720, 3], pixels in [0, 1]                 (data.py:128-130,191-193)
    y_{train,test}_flic.npy   float32 [n, 60, 90, 10], one 3x3 binomial blob per joint    (data.py:165-196)
    pairwise_distribution.pickle   dict '<joint>_<cond>' -> float64 [120, 180], 90 keys   (prepare_pairwise_distribution.py)

Host-side numpy: this is file preparation done once per dataset, not part of the per-step path (which starts at
`DeviceFeed.submit`).  The reference's quirks are kept and named: annotations are clamped to the image (`data.py:173`), the
blob is placed at `int()` of the /8 coordinate on a 5-pixel padded canvas (`:181-188`), `flip_backward_poses` swaps through
numpy views so that both sides end up with the RIGHT joint's coordinates (`:40-48`), the histogram is normalised by a float32
sum and smoothed with a 9x9 binomial kernel (`prepare_pairwise_distribution.py:13-14,45-47`).
"""
import os
import pickle

import numpy as np

JOINT_IDS = ['lsho', 'lelb', 'lwri', 'rsho', 'relb', 'rwri', 'lhip', 'rhip', 'nose', 'torso']   # data.py:98 + 'torso' (:169)
# column of the FLIC `coords` array for every joint (data.py:99-104); 'torso' is written into column 28 (:168)
FLIC_COLUMN = {'lsho': 0, 'lelb': 1, 'lwri': 2, 'rsho': 3, 'relb': 4, 'rwri': 5, 'lhip': 6, 'rhip': 9, 'nose': 16, 'torso': 28}
IMAGE_H, IMAGE_W = 480, 720
MAP_H, MAP_W = 60, 90
_PAD = 5                                                            # data.py:114


def flip_backward_poses(coords):
    """data.py:35-49 on a [2, 29] (x; y) array, in place.  The reference swaps through numpy VIEWS: after
    `coords[:, left] = right_view` the assignment `coords[:, right] = left_view` copies the already overwritten column, so a
    backward-facing pose ends with both sides holding the right joint.  Reproduced, because the labels on disk were made so."""
    if coords[0, FLIC_COLUMN['lhip']] < coords[0, FLIC_COLUMN['rhip']]:
        for left, right in zip(['lwri', 'lelb', 'lhip', 'lsho'], ['rwri', 'relb', 'rhip', 'rsho']):
            coords[:, FLIC_COLUMN[left]] = coords[:, FLIC_COLUMN[right]]
    return coords


def read_flic(mat_path):
    """data_FLIC.mat -> list of dicts {coords [2,29] float64, torsobox [4], is_train, file} (data.py:95-96,106-108,125)."""
    from scipy.io import loadmat
    out = []
    for e in loadmat(mat_path)['examples'][0]:
        out.append({'coords': np.array(e[2], dtype=np.float64), 'torsobox': np.array(e[6][0], dtype=np.float64),
                    'is_train': int(e[7][0, 0]) == 1, 'file': str(e[3][0])})
    return out


def joint_positions(example, torso='mean', flip=True):
    """(row, col) of the ten joints in heat-map units, float64 [10, 2] (data.py:122-123,165-179).
    torso='mean': mean of lsho, rhip, rsho, lhip as data.py:166-168 has it today; torso='box': centre of the FLIC torso box -
    the recipe the shipped pairwise table was made with, used together with flip=False."""
    c = np.array(example['coords'], dtype=np.float64)
    if flip:
        c = flip_backward_poses(c)
    if torso == 'mean':
        c[:, 28] = (c[:, 0] + c[:, 9] + c[:, 3] + c[:, 6]) / 4
    elif torso == 'box':
        x1, y1, x2, y2 = example['torsobox']
        c[:, 28] = [(x1 + x2) / 2, (y1 + y2) / 2]
    else:
        raise ValueError("torso must be 'mean' or 'box'")
    pos = np.empty([len(JOINT_IDS), 2])
    for j, name in enumerate(JOINT_IDS):
        x, y = c[0, FLIC_COLUMN[name]], c[1, FLIC_COLUMN[name]]
        pos[j] = (max(min(y, IMAGE_H), 0) / 8, max(min(x, IMAGE_W), 0) / 8)      # clamp to the image, then /8 (:173,179)
    return pos


def heat_map_labels(positions):
    """positions [n, J, 2] (row, col) -> float32 [n, 60, 90, J]: kernel = [1,2,1]^T [1,2,1] / 16 written at rows
    int(r+5-1) .. int(r+5+2) of a canvas padded by 5, then cropped (data.py:110-114,180-189; the crop clips border blobs)."""
    positions = np.asarray(positions, dtype=np.float64)
    n, J, _ = positions.shape
    k = np.outer([1, 2, 1], [1, 2, 1]).astype(np.float32) / 16
    canvas = np.zeros([n, MAP_H + 2 * _PAD, MAP_W + 2 * _PAD, J], dtype=np.float32)
    r0 = (positions[..., 0] + _PAD - 1).astype(np.int64)          # int() truncates towards zero; arguments are >= 4 here
    c0 = (positions[..., 1] + _PAD - 1).astype(np.int64)
    ii, jj = np.meshgrid(np.arange(n), np.arange(J), indexing='ij')
    for dr in range(3):
        for dc in range(3):
            canvas[ii, r0 + dr, c0 + dc, jj] = k[dr, dc]
    return np.ascontiguousarray(canvas[:, _PAD:_PAD + MAP_H, _PAD:_PAD + MAP_W, :])


def write_labels(mat_path, out_dir, torso='mean', flip=True):
    """y_train_flic.npy / y_test_flic.npy as data.py:115-196 writes them.  Returns {'train': n, 'test': n}."""
    ex = read_flic(mat_path)
    os.makedirs(out_dir, exist_ok=True)
    counts = {}
    for split, want in (('train', True), ('test', False)):
        pos = np.array([joint_positions(e, torso, flip) for e in ex if e['is_train'] == want])
        y = heat_map_labels(pos)
        np.save(os.path.join(out_dir, 'y_%s_flic.npy' % split), y)
        counts[split] = len(y)
    return counts


def write_images(mat_path, images_dir, out_dir, limit=None):
    """x_train_flic.npy / x_test_flic.npy: every image as float32 / 255 (data.py:125-130,191-193; iclr_data_preparation off, the
    reference's recommendation, :90-93).  Decoding uses PIL (the reference used imageio - the same 8-bit RGB arrays)."""
    from PIL import Image
    ex = read_flic(mat_path)
    os.makedirs(out_dir, exist_ok=True)
    counts = {}
    for split, want in (('train', True), ('test', False)):
        files = [e['file'] for e in ex if e['is_train'] == want][:limit]
        x = np.empty([len(files), IMAGE_H, IMAGE_W, 3], dtype=np.float32)
        for i, f in enumerate(files):
            img = np.asarray(Image.open(os.path.join(images_dir, f)).convert('RGB'))
            if img.shape != (IMAGE_H, IMAGE_W, 3):
                raise ValueError('%s is %s, expected %dx%dx3 (FLIC frames)' % (f, img.shape, IMAGE_H, IMAGE_W))
            x[i] = img.astype(np.float32) / 255
        np.save(os.path.join(out_dir, 'x_%s_flic.npy' % split), x)
        counts[split] = len(files)
    return counts


def load_split(data_dir, split, mmap=True):
    """(x, y) of main.py:286-294: float32 [n,480,720,3] in [0,1] and float32 [n,60,90,10]; memory-mapped by default."""
    mode = 'r' if mmap else None
    x = np.load(os.path.join(data_dir, 'x_%s_flic.npy' % split), mmap_mode=mode)
    y = np.load(os.path.join(data_dir, 'y_%s_flic.npy' % split), mmap_mode=mode)
    if x.shape[1:] != (IMAGE_H, IMAGE_W, 3) or y.shape[1:3] != (MAP_H, MAP_W) or len(x) != len(y):
        raise ValueError('unexpected dataset shapes %s %s' % (x.shape, y.shape))
    return x, y


# ------------------------------------------------------------------------------------------------ pairwise prior
def _peaks(y_train):
    """per (image, joint): the (rows, cols) of the maxima of the label map, as np.where gives them (row-major)."""
    n, H, W, J = y_train.shape
    flat = y_train.transpose(0, 3, 1, 2).reshape(n, J, H * W)
    mx = flat.max(axis=2, keepdims=True)
    peaks = [[None] * J for _ in range(n)]
    single = (flat == mx).sum(axis=2) == 1
    arg = flat.argmax(axis=2)
    for i in range(n):
        for j in range(J):
            if single[i, j]:
                peaks[i][j] = (np.array([arg[i, j] // W]), np.array([arg[i, j] % W]))
            else:                                   # a blob clipped by the border has several equal maxima
                idx = np.nonzero(flat[i, j] == mx[i, j, 0])[0]
                peaks[i][j] = (idx // W, idx % W)
    return peaks


def pairwise_distribution(y_train, joint_ids=None):
    """prepare_pairwise_distribution.py:29-56: for every ordered pair (joint, cond) the histogram of label-peak displacements on a
    [2H, 2W] grid centred at (H, W), normalised by its float32 sum and smoothed with the 9x9 binomial kernel ('same', zero fill).
    Returns the dict in the reference's insertion order, float64 arrays.  Where the two peak sets have different sizes and neither
    is a single point (numpy broadcasting would raise in the reference; does not occur in FLIC) the common prefix is used."""
    from scipy import signal
    joint_ids = list(joint_ids or JOINT_IDS)
    n, H, W, J = y_train.shape
    if J != len(joint_ids):
        raise ValueError('y_train has %d channels, %d joint names given' % (J, len(joint_ids)))
    coefs = np.array([[1, 8, 28, 56, 70, 56, 28, 8, 1]], dtype=np.uint16) / 256
    kernel = coefs.T @ coefs
    peaks = _peaks(np.asarray(y_train))
    out = {}
    for a, joint in enumerate(joint_ids):
        for b, cond in enumerate(joint_ids):
            if a == b:
                continue
            pd = np.zeros([2 * H, 2 * W])
            for i in range(n):
                (rj, cj), (rc, cc) = peaks[i][a], peaks[i][b]
                if len(rj) != len(rc) and len(rj) != 1 and len(rc) != 1:
                    m = min(len(rj), len(rc))
                    rj, cj, rc, cc = rj[:m], cj[:m], rc[:m], cc[:m]
                pd[H + (rj - rc), W + (cj - cc)] += 1          # fancy-index '+=': repeated cells count once, as in the reference
            pd = pd / np.float32(pd.sum())
            out[joint + '_' + cond] = signal.convolve2d(pd, kernel, mode='same', boundary='fill', fillvalue=0)
    return out


def write_pairwise_distribution(table, path):
    """The reference's pickle (protocol 4 = pickle.HIGHEST_PROTOCOL of its Python 3.5/3.6, prepare_pairwise_distribution.py:69-70),
    or a compressed .npz with the same keys when the path ends in .npz (the form this package ships)."""
    for k, v in table.items():
        if np.asarray(v).dtype != np.float64 or np.asarray(v).ndim != 2:
            raise ValueError('pairwise table entries must be float64 matrices (%s)' % k)
    if path.endswith('.npz'):
        np.savez_compressed(path, **table)
    else:
        with open(path, 'wb') as f:
            pickle.dump(dict(table), f, protocol=4)
