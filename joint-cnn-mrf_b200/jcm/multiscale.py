"""Multi-scale test-time inference (reference main.py:326-425).

For every test image the reference builds 8 rescaled copies on the HOST (4 zero-padded by 1.1-1.4x and 4 centre crops of 0.7-1.0x,
each resized back to 480x720 with `skimage.transform.resize`), runs the network on that batch of 8, maps the 8 heat-map sets back
to the original frame (inverse crop / pad + resize), averages them and takes the arg-max.  As in the reference this is host-side
numpy around one forward call per image; the forward call is `jcm.tower_forward` on the GPU (or any callable, see `get_predictions`).

`resize` restates `skimage.transform.resize(image, (rows, cols))` with the defaults of the scikit-image the reference ran on
(0.13.x, early 2018: order=1, mode='constant', cval=0, clip=True, preserve_range=False, no anti-aliasing):
  * sample position of output pixel o:  s = (o + 0.5) * in/out - 0.5      (pixel centres at half integers; `resize` builds this
    affine map from three corner correspondences),
  * bilinear interpolation between floor(s) and ceil(s); a neighbour outside the image contributes cval = 0 (`get_pixel2d`,
    mode 'C'), so borders fade towards 0 when shrinking a padded image,
  * clip=True: the result is clipped to [min, max] of the input, except that pixels exactly equal to cval stay cval when cval lies
    outside that range (`warp`: `preserve_cval`),
  * float input is converted to float64 (`img_as_float`).
scikit-image is not installable here, so this restatement is pinned against `scipy.ndimage.map_coordinates(order=1,
mode='grid-constant')` - an independent implementation of the same published algorithm - not against scikit-image itself
(tests/test_multiscale.py; DESIGN.md 7 states the gap).
"""
import numpy as np

PAD_ARRAY = (1.1, 1.2, 1.3, 1.4)     # main.py:403
CROP_ARRAY = (0.7, 0.8, 0.9, 1.0)


def _axis_taps(n_in, n_out):
    s = (np.arange(n_out, dtype=np.float64) + 0.5) * (float(n_in) / n_out) - 0.5
    lo = np.floor(s).astype(np.int64)
    hi = np.ceil(s).astype(np.int64)
    return lo, hi, s - lo


def resize(image, output_shape, cval=0.0, clip=True):
    """image [h, w] or [h, w, c] (float) -> float64 [rows, cols(, c)], see the module docstring."""
    img = np.asarray(image, dtype=np.float64)
    if img.ndim not in (2, 3):
        raise ValueError('resize expects [h, w] or [h, w, c]')
    if img.size and (img.min() < -1.0 or img.max() > 1.0):
        raise ValueError('Images of type float must be between -1 and 1.')        # img_as_float's check
    rows, cols = int(output_shape[0]), int(output_shape[1])
    h, w = img.shape[:2]
    rlo, rhi, dr = _axis_taps(h, rows)
    clo, chi, dc = _axis_taps(w, cols)

    def gather(r, c):
        ok = ((r >= 0) & (r < h))[:, None] & ((c >= 0) & (c < w))[None, :]
        v = img[np.clip(r, 0, h - 1)[:, None], np.clip(c, 0, w - 1)[None, :]]
        return np.where(ok[..., None] if img.ndim == 3 else ok, v, cval)

    dcb = dc[None, :, None] if img.ndim == 3 else dc[None, :]
    drb = dr[:, None, None] if img.ndim == 3 else dr[:, None]
    top = (1 - dcb) * gather(rlo, clo) + dcb * gather(rlo, chi)
    bottom = (1 - dcb) * gather(rhi, clo) + dcb * gather(rhi, chi)
    out = (1 - drb) * top + drb * bottom
    if clip and img.size:
        lo, hi = img.min(), img.max()
        preserve = not (lo <= cval <= hi)
        mask = (out == cval) if preserve else None
        out = np.clip(out, lo, hi)
        if preserve:
            out[mask] = cval
    return out


def _zoom_about_centre(img, coef, h, w):
    """One geometry primitive for both directions of the multi-scale transform: the image content is scaled by 1/coef about the
    image centre and the result brought back to [h, w] with `resize`.
      coef > 1  - zero margins of round(extent * (coef - 1) / 2) pixels are added on every side (the content shrinks),
      coef <= 1 - the centred window [round((1 - coef) / 2 * extent), ... + round(coef * extent)) is kept (the content grows).
    The rounding (Python's round, half to even) and the use of the ORIGINAL extents h, w - not the array's own - are the
    reference's (main.py:331-347,356-378): they decide which pixel rows survive, so they must match."""
    if coef > 1:
        mh, mw = round(h * (coef - 1) / 2), round(w * (coef - 1) / 2)
        margins = ((mh, mh), (mw, mw)) + ((0, 0),) * (np.ndim(img) - 2)
        piece = np.pad(img, margins, 'constant', constant_values=0)
    else:
        top, left = round((1 - coef) / 2 * h), round((1 - coef) / 2 * w)
        piece = img[top:top + round(coef * h), left:left + round(coef * w)]
    return resize(piece, (h, w))


def get_different_scales(x, pad_array=PAD_ARRAY, crop_array=CROP_ARRAY, orig_h=None, orig_w=None):
    """main.py:326-349: x [h, w, c] -> float64 [len(pad) + len(crop), h, w, c]: the padded (zoomed-out) copies first, then the
    centre crops (zoomed in)."""
    h = x.shape[0] if orig_h is None else orig_h
    w = x.shape[1] if orig_w is None else orig_w
    return np.array([_zoom_about_centre(x, c, h, w) for c in tuple(pad_array) + tuple(crop_array)])


def scale_hm_back(hms, pad_array=PAD_ARRAY, crop_array=CROP_ARRAY, orig_h=None, orig_w=None):
    """main.py:352-381: the inverse geometry on the heat maps, scale by scale: a map computed from an input zoomed by coefficient c
    is zoomed by 1/c (maps of padded inputs are centre-cropped, maps of cropped inputs are zero-padded), back to [orig_h, orig_w]."""
    h = hms[0].shape[0] if orig_h is None else orig_h
    w = hms[0].shape[1] if orig_w is None else orig_w
    coefs = tuple(pad_array) + tuple(crop_array)
    # 1 / 1.0 == 1.0 takes the crop branch with the full window, which is what the reference's pad branch with zero margins does
    return np.array([_zoom_about_centre(hms[i], 1 / c, h, w) for i, c in enumerate(coefs)])


def argmax_hm(hm):
    """main.py:389-397: hm [1, h, w, K] (or [h, w, K]) -> int [2, K] (row, col) of the first maximum per joint."""
    hm = np.squeeze(hm)
    h, w, K = hm.shape
    raw = np.argmax(np.reshape(hm, [h * w, K]), axis=0)
    rows = raw // w
    return np.stack([rows, raw - rows * w], axis=0)


def gpu_forward(params, sm, ctx):
    """The reference's `sess.run([hm_pred_pd, hm_pred_sm], feed_dict={x_in, y_in, flag_train: False})` (main.py:408) on the jcm
    kernels: float arrays in, float32 numpy heat maps out."""
    import torch
    from .graph import tower_forward
    if ctx.flag_train:
        raise ValueError('multi-scale inference runs in inference mode: pass a Context with flag_train=False')
    dev = sm.energies.device

    def forward(x_np, y_np):
        x = torch.from_numpy(np.ascontiguousarray(x_np, dtype=np.float32)).to(dev)
        y = torch.from_numpy(np.ascontiguousarray(y_np, dtype=np.float32)).to(dev)
        out = tower_forward(x, y, params, sm, ctx)
        return out['hm_pd'].cpu().numpy(), out['hm_sm'].cpu().numpy()
    return forward


def get_predictions(X, Y, forward, det_rate=None, n=1100, pad_array=PAD_ARRAY, crop_array=CROP_ARRAY):
    """main.py:383-425.  X [N,480,720,3], Y [N,60,90,K+1] numpy; forward(x8, y8) -> (hm_pd [8,h,w,K], hm_sm [8,h,w,K]) (use
    `gpu_forward`).  det_rate(hm [1,h,w,K], y [1,h,w,K+1]) -> float, optional (the reference's wrist detection rate at r = 10,
    main.py:455-456,419-420).  Returns (coords_pd [2,K,N'], coords_sm [2,K,N'], mean det-rate pd, mean det-rate sm), N' = min(N, n)."""
    X, Y = X[:n], Y[:n]
    in_h, in_w = X.shape[1], X.shape[2]
    hm_h, hm_w = Y.shape[1], Y.shape[2]
    drs_pd, drs_sm, coords_pd, coords_sm = [], [], [], []
    for x_np, y_np in zip(X, Y):
        x_scales = get_different_scales(np.asarray(x_np), pad_array, crop_array, in_h, in_w)
        y_rep = np.repeat(np.expand_dims(np.asarray(y_np), 0), x_scales.shape[0], axis=0)
        hm_pd, hm_sm = forward(x_scales, y_rep)
        hm_pd = scale_hm_back(hm_pd, pad_array, crop_array, hm_h, hm_w)
        hm_sm = scale_hm_back(hm_sm, pad_array, crop_array, hm_h, hm_w)
        hm_pd = np.expand_dims(np.average(hm_pd, axis=0), 0)
        hm_sm = np.expand_dims(np.average(hm_sm, axis=0), 0)
        coords_pd.append(argmax_hm(hm_pd))
        coords_sm.append(argmax_hm(hm_sm))
        if det_rate is not None:
            drs_pd.append(float(det_rate(hm_pd, y_rep[:1])))
            drs_sm.append(float(det_rate(hm_sm, y_rep[:1])))
    dr = (float(np.average(drs_pd)), float(np.average(drs_sm))) if drs_pd else (None, None)
    return np.stack(coords_pd, axis=2), np.stack(coords_sm, axis=2), dr[0], dr[1]
