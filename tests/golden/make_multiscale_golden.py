"""Generates tests/golden/multiscale_reference.npz by running the REFERENCE'S OWN `get_different_scales` / `scale_hm_back`
(/root/reference/main.py:326-379, exec of the two function bodies) on seeded inputs.  The one dependency that is not installable
here, `skimage.transform.resize`, is bound to jcm.multiscale.resize (whose own pin is scipy.ndimage, tests/test_multiscale.py):
what this fixture pins is the crop / pad GEOMETRY and its rounding, i.e. which pixel rows and columns each scale keeps.
Run in the build container, where /root/reference exists; the fixture travels, this script's input does not.

    python tests/golden/make_multiscale_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'joint-cnn-mrf_b200'))
REF = '/root/reference/main.py'
OUT = os.path.join(HERE, 'multiscale_reference.npz')


class _NumpyWithLibPad:
    """main.py calls `np.lib.pad`, an alias of `np.pad` that NumPy 2 removed; everything else is numpy itself."""
    class lib:
        pad = staticmethod(np.pad)

    def __getattr__(self, name):
        return getattr(np, name)


def reference_functions():
    from jcm.multiscale import resize
    src = open(REF).read()
    body = src[src.index('def get_different_scales'):src.index('def get_predictions')]
    skimage = types.SimpleNamespace(transform=types.SimpleNamespace(resize=resize))
    env = {'np': _NumpyWithLibPad(), 'skimage': skimage}
    exec(body, env)
    return env['get_different_scales'], env['scale_hm_back']


def main():
    gds, shb = reference_functions()
    rng = np.random.default_rng(20260118)
    pad_array, crop_array = [1.1, 1.2, 1.3, 1.4], [0.7, 0.8, 0.9, 1.0]          # main.py:403
    out = {}
    for tag, (h, w, c) in {'a': (24, 36, 2), 'b': (15, 23, 1), 'c': (30, 45, 2)}.items():
        x = rng.random((h, w, c))
        xs = gds(x, pad_array, crop_array, h, w)
        hms = rng.random((8, h, w, c))
        back = shb(hms, pad_array, crop_array, h, w)
        out.update({tag + '_x': x.astype(np.float32), tag + '_scales': xs.astype(np.float32), tag + '_hms': hms.astype(np.float32),
                    tag + '_back': back.astype(np.float32)})
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
