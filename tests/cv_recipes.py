"""numpy restatement of the OpenCV float chain of HoVer-Net's post-process, used to pin
the arithmetic the CUDA kernels reproduce: tests/test_oracle_golden.py checks these against cv2 bit for bit, the
GPU tests check the kernels against cv2 itself.  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def normalize_minmax_f32(x):
    """cv2.normalize(x, None, 0, 1, NORM_MINMAX, CV_32F) for fp32 or fp64 input."""
    mn, mx = float(x.min()), float(x.max())
    scale = 1.0 / (mx - mn) if (mx - mn) > np.finfo(np.float64).eps else 0.0
    # cv::normalize for rtype == CV_32F: scale = (float)scale; shift = (float)dmin - (float)(smin * scale)
    a = np.float64(np.float32(scale))
    b = np.float64(np.float32(0.0) - np.float32(mn * a))
    return (x.astype(np.float64) * a + b).astype(np.float32)


def deriv_kernels_21():
    """cv2.getDerivKernels(1, 0, 21): (derivative, smoothing) integer kernels as fp64."""
    sm = np.array([1.0])
    for _ in range(20):
        sm = np.convolve(sm, [1.0, 1.0])
    d = np.array([1.0])
    for _ in range(19):
        d = np.convolve(d, [1.0, 1.0])
    d = np.convolve(d, [-1.0, 1.0])     # applied as correlation: k[t] multiplies x[i + t - 10]
    return d, sm


def _pad101(a, r, axis):
    pad = [(0, 0), (0, 0)]
    pad[axis] = (r, r)
    return np.pad(a, pad, mode="reflect")


def sobel21(x32, dx, dy):
    """cv2.Sobel(x32, CV_64F, dx, dy, ksize=21) for (dx, dy) in {(1, 0), (0, 1)}."""
    d, sm = deriv_kernels_21()
    kx, ky = (d, sm) if dx == 1 else (sm, d)
    H, W = x32.shape
    xp = _pad101(x32.astype(np.float64), 10, 1)
    row = kx[0] * xp[:, 0:W]
    for t in range(1, 21):
        row = row + kx[t] * xp[:, t:t + W]
    rp = _pad101(row, 10, 0)
    c = rp[10:10 + H]
    if dy == 0:
        acc = ky[10] * c
        for t in range(1, 11):
            acc = acc + ky[10 + t] * (rp[10 + t:10 + t + H] + rp[10 - t:10 - t + H])
    else:
        acc = np.zeros_like(c)
        for t in range(1, 11):
            acc = acc + ky[10 + t] * (rp[10 + t:10 + t + H] - rp[10 - t:10 - t + H])
    return acc


def gaussian_blur3(d64):
    """cv2.GaussianBlur(d64, (3, 3), 0) on fp64."""
    H, W = d64.shape
    xp = _pad101(d64, 1, 1)
    row = (0.25 * xp[:, 0:W] + 0.5 * xp[:, 1:W + 1]) + 0.25 * xp[:, 2:W + 2]
    rp = _pad101(row, 1, 0)
    return 0.5 * rp[1:H + 1] + 0.25 * (rp[0:H] + rp[2:H + 2])


ELLIPSE5 = np.array([[0, 0, 1, 0, 0], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [0, 0, 1, 0, 0]], np.uint8)


def resize_up2_linear(x):
    """cv2.resize(x, (0, 0), fx=2, fy=2) for fp32 HW or HWC input (OpenCV 4.13 of this image): sample positions
    (d + 0.5) / 2 - 0.5 (weights 0.75 / 0.25, clamped at the borders), horizontal pass then vertical.  One-channel
    images with both sides >= 2 take OpenCV's 2x fast path and interpolate as fma(a, x1 - x0, x0); everything
    else as x0 * (1 - a) + x1 * a with separately rounded products."""
    f32 = np.float32

    def coords(n):
        d = np.arange(2 * n)
        f = (d + 0.5) * 0.5 - 0.5
        s = np.floor(f).astype(int)
        a = (f - s).astype(f32)
        lo = s < 0; s[lo] = 0; a[lo] = 0
        hi = s >= n - 1; a[hi] = 0; s[hi] = n - 1
        return s, np.minimum(s + 1, n - 1), a

    def lerp(x0, x1, a):
        return (a.astype(np.float64) * (x1 - x0).astype(f32).astype(np.float64) + x0.astype(np.float64)).astype(f32)

    def mma(x0, x1, a):
        return (x0 * (f32(1) - a)).astype(f32) + (x1 * a).astype(f32)

    def one(S, mix):
        H, W = S.shape
        x0, x1, ax = coords(W)
        y0, y1, ay = coords(H)
        h = mix(S[:, x0], S[:, x1], np.broadcast_to(ax[None, :], (H, 2 * W)))
        return mix(h[y0, :], h[y1, :], np.broadcast_to(ay[:, None], (2 * H, 2 * W)))

    x = np.asarray(x, f32)
    if x.ndim == 2:
        return one(x, lerp if min(x.shape) >= 2 else mma)
    return np.stack([one(x[..., c], mma) for c in range(x.shape[2])], -1)
