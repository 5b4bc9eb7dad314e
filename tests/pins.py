"""Test helper: pins the two DISCONTINUOUS choices of the training graph (ReLU on/off, 2x2 max-pool arg-max) of the
oracle to the ones the GPU forward made, so that gradient comparisons are well conditioned.

Why: the GPU forward agrees with the fp64 oracle to ~1e-5, but an element whose pre-activation is within that error of
zero, or a pooling window whose two largest values are within it of each other, can fall on the other side in fp32.
The forward value changes by ~1e-6; the GRADIENT of that single element is switched on/off or re-routed, which moves
e.g. one channel's bias gradient by several per cent.  That is a property of the function (TensorFlow on a GPU and on
a CPU disagree in the same way), not of the kernels under test."""
import torch


def relu_masks(tap):
    """tap: {'<layer>/relu': ReLU output (CUDA fp32)} as filled by jcm.model / Trainer.forward_backward."""
    return {k[:-5]: (v > 0).cpu() for k, v in tap.items() if k.endswith('/relu')}


def pool_select(jcm, tap, p):
    """Window element (2*dy + dx) the GPU max-pool took for every pooled layer (conv1_*, conv2_*), recomputed with the
    product's own BN-apply kernel (training-mode batch statistics): the first element equal to the pooled value."""
    sel = {}
    for key, a in tap.items():
        name = key[:-5]
        if not (name.startswith('conv1_') or name.startswith('conv2_')):
            continue
        C = a.shape[3]
        g, b = p[name + '/BatchNorm/gamma'], p[name + '/BatchNorm/beta']
        ss = jcm.ops.bn_scale_shift(a, g.clone(), b.clone(), torch.zeros(C, device=a.device), torch.ones(C, device=a.device),
                                    train=True, update_moving=False)
        full = jcm.ops.bn_apply_pool(a, ss, False, False, want_planes=False, want_f32=True).cpu()
        pooled = jcm.ops.bn_apply_pool(a, ss, True, False, want_planes=False, want_f32=True).cpu()
        B, H, W, _ = full.shape
        Ho, Wo = pooled.shape[1], pooled.shape[2]
        pad = torch.full((B, 2 * Ho, 2 * Wo, C), float('-inf'))
        pad[:, :H, :W] = full
        win = pad.reshape(B, Ho, 2, Wo, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Ho, Wo, 4, C)
        eq = win == pooled.unsqueeze(3)
        assert bool(eq.any(3).all()), 'pooled value not found in its window for ' + name
        sel[name] = eq.to(torch.uint8).argmax(3)       # first window element equal to the maximum
    return sel
