"""Oracle: image metrics of the reference's evaluation.  TEST INFRASTRUCTURE ONLY.

SSIM as called at /root/reference/src/run-nerf.py:183-187:
``skimage.metrics.structural_similarity(rgb, rgb_gt, channel_axis=-1, data_range=1.0,
gaussian_weights=True)``.  scikit-image (0.20.0 in the reference's environment.yaml) is NOT
installed here — **parity unpinned**: the published algorithm (Wang et al. 2004 as implemented
by scikit-image: sigma = 1.5, truncate = 3.5 -> 11x11 window, scipy.ndimage.gaussian_filter with
'reflect' boundaries per channel, population covariance, K1 = 0.01, K2 = 0.03, crop of
(win-1)/2 border pixels, mean over pixels then channels) is restated with scipy.
"""
import numpy as np
from scipy.ndimage import gaussian_filter


def ssim(im1, im2, data_range=1.0, sigma=1.5, truncate=3.5, k1=0.01, k2=0.03):
    im1, im2 = np.asarray(im1, np.float64), np.asarray(im2, np.float64)
    pad = int(truncate * sigma + 0.5)
    out = []
    for c in range(im1.shape[-1]):
        x, y = im1[..., c], im2[..., c]
        f = lambda a: gaussian_filter(a, sigma, truncate=truncate, mode="reflect")  # noqa: E731
        ux, uy = f(x), f(y)
        vx, vy, vxy = f(x * x) - ux * ux, f(y * y) - uy * uy, f(x * y) - ux * uy
        c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        out.append(s[pad:-pad, pad:-pad].mean())
    return float(np.mean(out))
