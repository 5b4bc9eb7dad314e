"""Checkpoint / resume of the training state under the reference's TensorFlow variable names.

The reference saves with `tf.train.Saver(max_to_keep=50)` (main.py:604) once `epoch > n_epochs // 2` (main.py:663-666) and
restores with `saver.restore` (main.py:609-617); the checkpoint holds every global variable: the model variables, the
BatchNorm moving statistics, the optimizer slots and `n_iters` (so that the piecewise-constant learning-rate schedule of
main.py:468-470,491-492 resumes where it stopped).  TensorFlow's on-disk format cannot be read or written here (no
TensorFlow); this module keeps the variable NAMES ('conv1_fullres/weights', '.../BatchNorm/moving_mean',
'bn_sm/BatchNorm/gamma', 'energy_<joint>_<cond>', 'bias_<joint>_<cond>', '<var>/Adam', '<var>/Adam_1', 'n_iters') in a
plain numpy .npz so a converter on a machine with TensorFlow is a ten-line name-for-name copy."""
import numpy as np
import torch


def state_dict(trainer):
    """name -> numpy array for every variable of the trainer (model, BN moving statistics, pairwise parameters, optimizer slots)."""
    out = {}
    p, sm = trainer.p, trainer.sm
    for k, v in p.items():
        out[k] = v.detach().cpu().numpy()
    for k, v in sm.as_dict().items():
        out[k] = v.detach().cpu().numpy()
    # optimizer slots under TF's slot names: Adam -> '<var>/Adam' (m), '<var>/Adam_1' (v); Momentum -> '<var>/Momentum'
    slots = (('Adam', trainer.m), ('Adam_1', trainer.v)) if trainer.optimizer == 'adam' else (('Momentum', trainer.m),)
    for name, view in _flat_views(trainer).items():
        lo, shape = view
        n = int(np.prod(shape))
        for slot, buf in slots:
            out[name + '/' + slot] = buf[lo:lo + n].view(shape).detach().cpu().numpy()
    out['n_iters'] = np.asarray(trainer.t, dtype=np.int64)
    out['_meta/optimizer'] = np.asarray(trainer.optimizer)
    out['_meta/joint_names'] = np.asarray(trainer.sm.joint_names)
    return out


def _flat_views(trainer):
    """TF variable name -> (offset into the flat buffers, shape), energies / biases split per pair."""
    views = {}
    base = trainer.flat.data_ptr()
    for k, g in trainer.g.items():
        lo = (trainer.p[k].data_ptr() - base) // 4 if not k.startswith('sm/') else None
        if lo is not None:
            views[k] = (lo, tuple(g.shape))
    sm = trainer.sm
    P, H2, W2 = sm.energies.shape
    e0 = (sm.energies.data_ptr() - base) // 4
    b0 = (sm.biases.data_ptr() - base) // 4
    for i, key in enumerate(sm.keys):
        views['energy_' + key] = (e0 + i * H2 * W2, (1, H2, W2, 1))
        views['bias_' + key] = (b0 + i * (H2 // 2) * (W2 // 2), (1, H2 // 2, W2 // 2, 1))
    for k in ('gamma', 'beta'):
        views['bn_sm/BatchNorm/' + k] = ((sm.bn[k].data_ptr() - base) // 4, tuple(sm.bn[k].shape))
    return views


def save_checkpoint(path, trainer):
    """Writes <path> (.npz).  Call on rank 0 only: the replicas hold identical state - parameters and optimizer slots because
    every replica applies the same all-reduced gradient, the BatchNorm moving statistics because Trainer.apply() reduces them
    across replicas every step (bench.py reports `replicas_identical` from checksums of both)."""
    np.savez(path, **state_dict(trainer))


def load_checkpoint(path, trainer, strict=True):
    """Restores a trainer built with the same architecture (main.py:609-617): variables, moving statistics, optimizer slots and
    the update counter.  Unknown / missing names raise unless strict=False."""
    with np.load(path, allow_pickle=False) as z:
        data = {k: z[k] for k in z.files}
    if str(data.get('_meta/optimizer', trainer.optimizer)) != trainer.optimizer and strict:
        raise ValueError('checkpoint was written by the %s optimizer, the trainer uses %s' % (data['_meta/optimizer'], trainer.optimizer))
    p, sm = trainer.p, trainer.sm
    targets = dict(p)
    targets.update(sm.as_dict())
    missing = [k for k in targets if k not in data]
    if missing and strict:
        raise KeyError('checkpoint lacks variables: %s' % missing[:5])
    for k, t in targets.items():
        if k in data:
            if tuple(data[k].shape) != tuple(t.shape):
                raise ValueError('shape of %s: checkpoint %s vs model %s' % (k, data[k].shape, tuple(t.shape)))
            t.copy_(torch.from_numpy(data[k]).to(t.device))
    slots = (('Adam', trainer.m), ('Adam_1', trainer.v)) if trainer.optimizer == 'adam' else (('Momentum', trainer.m),)
    for name, (lo, shape) in _flat_views(trainer).items():
        n = int(np.prod(shape))
        for slot, buf in slots:
            key = name + '/' + slot
            if key in data:
                buf[lo:lo + n].view(shape).copy_(torch.from_numpy(data[key]).to(buf.device))
            elif strict:
                raise KeyError('checkpoint lacks optimizer slot ' + key)
    trainer.t = int(data['n_iters']) if 'n_iters' in data else trainer.t
    from .graph import params_updated
    params_updated()                 # packed operand planes are stale (in every Context)
    return trainer
