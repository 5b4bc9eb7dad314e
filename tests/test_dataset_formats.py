"""the on-disk formats either side of the path (labels, images, pairwise-prior pickle) as `jcm.dataset` writes
them, against (a) outputs of the reference's own statements (tests/golden/dataset_reference.npz, made by
tests/golden/make_dataset_golden.py from /root/reference/data.py and prepare_pairwise_distribution.py), (b) the oracle's
independent restatement, (c) the reference's data_FLIC.mat and shipped prior table where the reference mount exists.  CPU only."""
import os
import pickle
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'joint-cnn-mrf_b200'), os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

REF = '/root/reference'
GOLD = os.path.join(ROOT, 'tests', 'golden', 'dataset_reference.npz')


@pytest.fixture(scope='module')
def ds(built_lib):
    from jcm import dataset
    return dataset


@pytest.fixture(scope='module')
def gold():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


def _examples(coords):
    return [{'coords': c, 'torsobox': np.zeros(4), 'is_train': True, 'file': ''} for c in coords]


def test_labels_equal_the_reference_statements_bit_for_bit(ds, gold):
    """data.py:122-123,165-189 executed verbatim vs jcm.dataset: clamped annotations, border-clipped blobs, the view-aliasing
    flip of backward-facing poses, torso = mean of shoulders and hips."""
    pos = np.array([ds.joint_positions(e) for e in _examples(gold['coords'])])
    y = ds.heat_map_labels(pos)
    assert y.dtype == np.float32 and y.shape == gold['labels'].shape == (24, 60, 90, 10)
    np.testing.assert_array_equal(y, gold['labels'])
    # the edge cases really are in the fixture
    assert gold['labels'][1, :, :, 8].sum() == np.float32(1 / 16)             # far corner: one cell of the blob survives
    assert abs(gold['labels'][5, :, :, 0].sum() - 1) < 1e-6                   # interior blob sums to 1
    flipped = ds.joint_positions(_examples(gold['coords'])[3])
    assert np.array_equal(flipped[0], flipped[3]) and np.array_equal(flipped[2], flipped[5])   # both sides = the right joint
    frontal = ds.joint_positions(_examples(gold['coords'])[4])
    assert not np.array_equal(frontal[0], frontal[3])


def test_labels_equal_the_oracle_restatement(ds):
    import pairwise_prior as pp
    rng = np.random.default_rng(3)
    pos = np.stack([rng.uniform(0, 60, size=[50, 10]), rng.uniform(0, 90, size=[50, 10])], axis=2)
    pos[0, 0] = (0.0, 0.0)
    pos[1, 1] = (60.0, 90.0)
    np.testing.assert_array_equal(ds.heat_map_labels(pos), pp.target_heat_maps(pos))


def test_pairwise_distribution_equals_the_reference_function(ds, gold):
    """prepare_pairwise_distribution.py:29-48 executed verbatim on the golden labels (which include clipped blobs with several
    equal maxima) vs jcm.dataset.pairwise_distribution; keys in the reference's insertion order."""
    table = ds.pairwise_distribution(gold['labels'])
    keys = list(table)
    assert len(keys) == 90 and keys[0] == 'lsho_lelb' and keys[-1] == 'torso_nose'
    for k in ('lwri_lelb', 'nose_torso', 'rsho_lsho', 'nose_rwri'):
        assert table[k].dtype == np.float64 and table[k].shape == (120, 180)
        np.testing.assert_allclose(table[k], gold['prior_' + k], rtol=0, atol=1e-17)


def test_pickle_and_npz_round_trip_through_the_product_loader(ds, gold, tmp_path):
    import jcm
    table = ds.pairwise_distribution(gold['labels'][:6])
    for name in ('pairwise_distribution.pickle', 'pairwise_distribution.npz'):
        path = str(tmp_path / name)
        ds.write_pairwise_distribution(table, path)
        back = jcm.get_pairwise_distr(path)                      # main.py:297-299
        assert list(back) == list(table)
        for k in table:
            np.testing.assert_array_equal(back[k], table[k])
    with open(str(tmp_path / 'pairwise_distribution.pickle'), 'rb') as f:
        assert isinstance(pickle.load(f), dict)
    with pytest.raises(ValueError):
        ds.write_pairwise_distribution({'a_b': np.zeros([120, 180], dtype=np.float32)}, str(tmp_path / 'bad.pickle'))


def test_image_and_label_files_have_the_reference_formats(ds, tmp_path):
    """x_*.npy float32 [n,480,720,3] in [0,1] (data.py:128-130), y_*.npy float32 [n,60,90,10]; read back as main.py:286-294 does."""
    from PIL import Image
    from scipy.io import savemat
    rng = np.random.default_rng(5)
    n = 3
    # data_FLIC.mat holds a MATLAB struct array `examples`; loadmat returns records indexed by position (data.py:95-96,106-108,125)
    fields = ['poselet_hit_idx', 'moviename', 'coords', 'filepath', 'imgdims', 'currframe', 'torsobox', 'istrain', 'istest']
    ex = np.zeros((1, n), dtype=[(f, 'O') for f in fields])
    imgs = []
    for i in range(n):
        img = rng.integers(0, 256, size=[480, 720, 3], dtype=np.uint8)
        imgs.append(img)
        Image.fromarray(img).save(str(tmp_path / ('f%d.png' % i)))
        for f in fields:
            ex[0, i][f] = np.zeros((1, 1))
        ex[0, i]['coords'] = np.stack([rng.uniform(50, 650, 29), rng.uniform(50, 430, 29)])
        ex[0, i]['filepath'] = np.array(['f%d.png' % i])
        ex[0, i]['torsobox'] = np.array([[100.0, 100.0, 200.0, 300.0]])
        ex[0, i]['istrain'] = np.array([[1 if i < 2 else 0]])
    mat = str(tmp_path / 'flic.mat')
    savemat(mat, {'examples': ex})
    flic = ds.read_flic(mat)
    assert flic[0]['file'] == 'f0.png' and flic[0]['coords'].shape == (2, 29) and flic[0]['torsobox'].shape == (4,)
    assert [e['is_train'] for e in flic] == [True, True, False]
    out = str(tmp_path / 'out')
    assert ds.write_images(mat, str(tmp_path), out) == {'train': 2, 'test': 1}
    assert ds.write_labels(mat, out) == {'train': 2, 'test': 1}
    x, y = ds.load_split(out, 'train')
    assert x.dtype == np.float32 and x.shape == (2, 480, 720, 3) and y.dtype == np.float32 and y.shape == (2, 60, 90, 10)
    np.testing.assert_array_equal(np.asarray(x[1]), imgs[1].astype(np.float32) / 255)
    assert float(x.max()) <= 1.0 and float(x.min()) >= 0.0
    xt, yt = ds.load_split(out, 'test')
    assert len(xt) == 1 and abs(float(yt[0, :, :, 4].sum()) - 1) < 1e-6


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'data_FLIC.mat')), reason='reference mount not present')
def test_flic_labels_and_the_shipped_prior_table(ds):
    """On the real annotation file: 3987 / 1016 examples (data.py:13), labels equal the oracle's for both recipes, and the 'box'
    recipe without the flip reproduces entries of the table this package ships (the one pinned to the reference's corrupt pickle)."""
    import pairwise_prior as pp
    flic = ds.read_flic(os.path.join(REF, 'data_FLIC.mat'))
    train = [e for e in flic if e['is_train']]
    assert len(train) == 3987 and len(flic) - len(train) == 1016
    pos = np.array([ds.joint_positions(e) for e in train])
    np.testing.assert_array_equal(pos, pp.joint_positions(os.path.join(REF, 'data_FLIC.mat'), 'scripts', 'train'))
    pos_box = np.array([ds.joint_positions(e, torso='box', flip=False) for e in train])
    np.testing.assert_array_equal(pos_box, pp.joint_positions(os.path.join(REF, 'data_FLIC.mat'), 'shipped', 'train'))
    y = ds.heat_map_labels(pos_box)
    sub = ds.pairwise_distribution(y[:, :, :, [2, 1, 9]], ['lwri', 'lelb', 'torso'])
    with np.load(os.path.join(ROOT, 'joint-cnn-mrf_b200', 'jcm', 'data', 'pairwise_distribution.npz')) as z:
        for k in ('lwri_lelb', 'lelb_torso', 'torso_lwri'):
            np.testing.assert_allclose(sub[k], z[k], rtol=0, atol=1e-15)
