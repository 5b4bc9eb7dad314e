"""Generates tests/golden/dataset_reference.npz by running the REFERENCE'S OWN statements (exec of the label block of
/root/reference/data.py and of compute_pairwise_distribution from /root/reference/prepare_pairwise_distribution.py) on seeded
inputs.  Run in the build container, where /root/reference exists; the fixture travels, this script's inputs do not.

    python tests/golden/make_dataset_golden.py
"""
import os
import re
import textwrap

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dataset_reference.npz')


def reference_label_block():
    """data.py:165-189: from `hmap = []` to `hmap = np.stack(hmap, axis=2)`, de-indented; plus flip_backward_poses (:35-49)."""
    src = open(os.path.join(REF, 'data.py')).read()
    body = src[src.index('            hmap = []\n'):src.index('            hmaps.append(hmap)')]
    flip = src[src.index('def flip_backward_poses'):src.index('def how_many_backward_poses')]
    return textwrap.dedent(body), flip


class _NumpyWithLibPad:
    """data.py:182 calls `np.lib.pad`, an alias of `np.pad` that NumPy 2 removed; everything else is numpy itself."""
    class lib:
        pad = staticmethod(np.pad)

    def __getattr__(self, name):
        return getattr(np, name)


def run_reference_labels(coords_list):
    body, flip = reference_label_block()
    env = {'np': _NumpyWithLibPad()}
    env['dict'] = {'lsho': 0, 'lelb': 1, 'lwri': 2, 'rsho': 3, 'relb': 4, 'rwri': 5, 'lhip': 6, 'lkne': 7, 'lank': 8, 'rhip': 9,
                   'rkne': 10, 'rank': 11, 'leye': 12, 'reye': 13, 'lear': 14, 'rear': 15, 'nose': 16, 'msho': 17, 'mhip': 18,
                   'mear': 19, 'mtorso': 20, 'mluarm': 21, 'mruarm': 22, 'mllarm': 23, 'mrlarm': 24, 'mluleg': 25, 'mruleg': 26,
                   'mllleg': 27, 'torso': 28}                                        # data.py:99-104
    exec(flip, env)
    coefs = np.array([[1, 2, 1]], dtype=np.float32) / 4                              # data.py:110-114
    env.update(joint_ids=['lsho', 'lelb', 'lwri', 'rsho', 'relb', 'rwri', 'lhip', 'rhip', 'nose'], kernel=coefs.T @ coefs,
               temp=1, pad=5, orig_h=480, orig_w=720, x_name='x_test_flic', iclr_data_preparation=False)
    out = []
    for c in coords_list:
        env['flic_coords'] = env['flip_backward_poses'](np.array(c, dtype=np.float64))   # data.py:122-123
        exec(body, env)
        out.append(env['hmap'])
    return np.array(out, dtype=np.float32)


def run_reference_prior(y_train, pairs):
    src = open(os.path.join(REF, 'prepare_pairwise_distribution.py')).read()
    fn = src[src.index('def compute_pairwise_distribution'):src.index('pairwise_distribution = {}')]
    from scipy import signal
    coefs = np.array([[1, 8, 28, 56, 70, 56, 28, 8, 1]], dtype=np.uint16) / 256      # prepare_pairwise_distribution.py:13-14
    env = {'np': np, 'signal': signal, 'y_train': y_train, 'kernel': coefs.T @ coefs,
           'joint_ids': ['lsho', 'lelb', 'lwri', 'rsho', 'relb', 'rwri', 'lhip', 'rhip', 'nose', 'torso']}
    exec(fn, env)
    return {a + '_' + b: env['compute_pairwise_distribution'](a, b) for a, b in pairs}


def main():
    rng = np.random.default_rng(20260117)
    m = 24
    coords = np.zeros([m, 2, 29])
    coords[:, 0, :] = rng.uniform(40, 680, size=[m, 29])       # x
    coords[:, 1, :] = rng.uniform(30, 450, size=[m, 29])       # y
    coords[0, :, 2] = (-15.0, 500.0)                           # annotation outside the image: clamped (data.py:173)
    coords[1, :, 16] = (720.0, 480.0)                          # exactly on the far corner: blob clipped to one cell
    coords[2, :, 5] = (3.9, 4.1)                               # near the origin: blob clipped on two sides
    coords[3, 0, 6], coords[3, 0, 9] = 100.0, 300.0            # lhip left of rhip: backward-facing, flipped (with the aliasing quirk)
    coords[4, 0, 6], coords[4, 0, 9] = 300.0, 100.0            # frontal: not flipped
    labels = run_reference_labels(coords)
    pairs = [('lwri', 'lelb'), ('nose', 'torso'), ('rsho', 'lsho'), ('nose', 'rwri')]
    prior = run_reference_prior(labels, pairs)
    np.savez_compressed(OUT, coords=coords, labels=labels, **{'prior_' + k: v for k, v in prior.items()})
    print('wrote', OUT, labels.shape, {k: float(v.sum()) for k, v in prior.items()})


if __name__ == '__main__':
    main()
