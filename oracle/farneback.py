"""ORACLE (test infrastructure, not product code) - numpy restatement of the dense optical flow the reference computes
on the fly for its FarneBack TwoStream variant:

    opticalflow_FarneBack_extractor, /root/reference/train.py:294-332
        cv2.calcOpticalFlowFarneback(prev_gray, gray, None, pyr_scale=0.5, levels=5, winsize=11, iterations=5,
                                     poly_n=5, poly_sigma=1.1, flags=0)

The algorithm lives in a third-party dependency that is not part of the reference tree: opencv-python (requirements.txt
pins no version; this image ships 4.13.0), modules/video/src/optflowgf.cpp.  Its published algorithm (G. Farneback,
"Two-frame motion estimation based on polynomial expansion", SCIA 2003, as implemented by OpenCV) is restated here
stage by stage - pyramid, Gaussian pre-blur, bilinear resize, polynomial expansion, matrix update, box-blurred 2x2
solves - with OpenCV's choice of float / double for every intermediate.

PINNED, to a tolerance: against cv2.calcOpticalFlowFarneback itself and against the golden produced by the reference's
own extractor (tests/golden/farneback_golden.npz, tools/make_golden_farneback.py).  Bit equality with cv2 is not
attainable: its GaussianBlur / resize run through IPP / AVX2 code whose summation order (and FMA use) differs by 1-2 ulp
from any plain restatement (24-44 % of the pixels of a blurred image differ in the last bits), and its box filter is a
sliding-window update that adds float-rounded row differences where this file sums each window directly in double.
tests/test_oracle_farneback.py measures the resulting flow differences (<= 6e-6 pixel on flows of several pixels) and
bounds them at 1e-4.

The CUDA path (csrc/flow.cu) follows THIS file operation by operation (explicitly rounded multiplies and adds, the same
summation order), so `cse_farneback` is compared with it bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], np.float32)


def cv_round(v: float) -> int:
    """cvRound: round half to even (lrint)."""
    return int(np.rint(v))


def gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    """cv::getGaussianKernel(ksize, sigma, CV_32F)."""
    if sigma <= 0 and ksize <= 9:
        tab = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
               7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
               9: [v / 256 for v in (4, 13, 30, 51, 60, 51, 30, 13, 4)]}
        return np.asarray(tab[ksize], F32)
    sx = sigma if sigma > 0 else ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8
    scale2x = -0.5 / (sx * sx)
    w = [math.exp(scale2x * (i - (ksize - 1) * 0.5) * (i - (ksize - 1) * 0.5)) for i in range(ksize)]   # libm exp, like the C side
    total = 0.0
    for v in w:
        total += v
    return np.asarray([v / total for v in w], np.float64).astype(F32)


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(idx)
    idx = idx.copy()
    for _ in range(64):
        bad = (idx < 0) | (idx >= n)
        if not bad.any():
            break
        idx = np.where(idx < 0, -idx, idx)
        idx = np.where(idx >= n, 2 * n - 2 - idx, idx)
    return idx


def gaussian_blur(img: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """cv2.GaussianBlur(float32 image, (ksize, ksize), sigma, sigma), BORDER_REFLECT_101: rows first, then columns;
    s = k0 * c + sum_j k_j * (left_j + right_j)."""
    k = gaussian_kernel(ksize, sigma)
    n = ksize // 2

    def along_x(a):
        w = a.shape[1]
        cols = _reflect101(np.arange(-n, w + n), w)
        p = a[:, cols]
        s = p[:, n:n + w] * k[n]
        for j in range(1, n + 1):
            s = s + (p[:, n - j:n - j + w] + p[:, n + j:n + j + w]) * k[n + j]
        return s.astype(F32)

    h = along_x(np.ascontiguousarray(img, F32))
    return np.ascontiguousarray(along_x(np.ascontiguousarray(h.T)).T)


def _linear_taps(dst: int, src: int, scale: float, horizontal: bool):
    """Source indices and float weights of cv2.resize's bilinear taps: columns clamp the fraction at the edges, rows
    keep it and clamp the two row indices instead (imgproc/resize.cpp)."""
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(F32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(F32)).astype(F32)
    if horizontal:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), (F32(1) - f).astype(F32), f


def resize_linear(img: np.ndarray, h: int, w: int, mul: float = 1.0) -> np.ndarray:
    """cv2.resize(float32 [Hs,Ws] or [Hs,Ws,C], (w, h), INTER_LINEAR): pixel-centre mapping, float weights, horizontal
    combination first; the result is multiplied by `mul` (float)."""
    img = np.ascontiguousarray(img, F32)
    hs, ws = img.shape[:2]
    x0, x1, a0, a1 = _linear_taps(w, ws, ws / w, True)
    y0, y1, b0, b1 = _linear_taps(h, hs, hs / h, False)
    if img.ndim == 3:
        a0, a1 = a0[None, :, None], a1[None, :, None]
        b0, b1 = b0[:, None, None], b1[:, None, None]
    else:
        a0, a1 = a0[None, :], a1[None, :]
        b0, b1 = b0[:, None], b1[:, None]
    hor = img[:, x0] * a0 + img[:, x1] * a1
    out = hor[y0] * b0 + hor[y1] * b1
    return (out * F32(mul)).astype(F32)


def prepare_gaussian(n: int, sigma: float):
    """FarnebackPrepareGaussian: g, x*g, x*x*g (float) and the four entries of the inverse moment matrix it uses."""
    if sigma < 1.1920929e-07:
        sigma = n * 0.3
    xs = np.arange(-n, n + 1)
    g = np.asarray([math.exp(-x * x / (2 * sigma * sigma)) for x in range(-n, n + 1)], np.float64).astype(F32)
    s = 0.0
    for v in g:
        s += float(v)
    s = 1.0 / s
    g = (g.astype(np.float64) * s).astype(F32)
    xg = (xs * g.astype(np.float64)).astype(F32)
    xxg = (xs * xs * g.astype(np.float64)).astype(F32)
    a = b = c = d = 0.0
    for y in range(-n, n + 1):
        for x in range(-n, n + 1):
            wgt, fx, fy = g[y + n] * g[x + n], F32(x), F32(y)             # float products, as the C expression types give
            a += float(wgt)
            b += float(wgt * fx * fx)
            c += float(wgt * fx * fx * fx * fx)
            d += float(wgt * fx * fx * fy * fy)
    # G = [[a,0,0,b,b,0],[0,b,0,0,0,0],[0,0,b,0,0,0],[b,0,0,c,d,0],[b,0,0,d,c,0],[0,0,0,0,0,d]]; cofactor inverse of
    # the {1, x^2, y^2} block
    det = a * (c * c - d * d) - b * (b * c - d * b) + b * (b * d - c * b)
    ig11, ig55 = 1.0 / b, 1.0 / d
    ig03 = -(b * c - b * d) / det
    ig33 = (a * c - b * b) / det
    return g[n:], xg[n:], xxg[n:], ig11, ig03, ig33, ig55


def poly_exp(img: np.ndarray, n: int, sigma: float) -> np.ndarray:
    """FarnebackPolyExp -> [H,W,5] float32: (d/dy-ish r3, d/dx-ish r2, r5, r4, r6) in OpenCV's storage order."""
    g, xg, xxg, ig11, ig03, ig33, ig55 = prepare_gaussian(n, sigma)
    h, w = img.shape
    img = np.ascontiguousarray(img, F32)
    t0 = img * g[0]
    t1 = np.zeros_like(img)
    t2 = np.zeros_like(img)
    ys = np.arange(h)
    for k in range(1, n + 1):
        s0 = img[np.maximum(ys - k, 0)]
        s1 = img[np.minimum(ys + k, h - 1)]
        p = s0 + s1
        t0 = t0 + g[k] * p
        t1 = t1 + xg[k] * (s1 - s0)
        t2 = t2 + xxg[k] * p
    xsi = np.arange(w)
    b1 = (t0 * g[0]).astype(np.float64)
    b3 = (t1 * g[0]).astype(np.float64)
    b5 = (t2 * g[0]).astype(np.float64)
    b2 = np.zeros((h, w))
    b4 = np.zeros((h, w))
    b6 = np.zeros((h, w))
    for k in range(1, n + 1):
        xp, xm = np.minimum(xsi + k, w - 1), np.maximum(xsi - k, 0)
        tg = (t0[:, xp] + t0[:, xm]).astype(np.float64)
        b1 = b1 + tg * float(g[k])
        b4 = b4 + tg * float(xxg[k])
        b2 = b2 + ((t0[:, xp] - t0[:, xm]) * xg[k]).astype(np.float64)
        b3 = b3 + ((t1[:, xp] + t1[:, xm]) * g[k]).astype(np.float64)
        b6 = b6 + ((t1[:, xp] - t1[:, xm]) * xg[k]).astype(np.float64)
        b5 = b5 + ((t2[:, xp] + t2[:, xm]) * g[k]).astype(np.float64)
    out = np.empty((h, w, 5), F32)
    out[..., 1] = b2 * ig11
    out[..., 0] = b3 * ig11
    out[..., 3] = b1 * ig03 + b4 * ig33
    out[..., 2] = b1 * ig03 + b5 * ig33
    out[..., 4] = b6 * ig55
    return out


def update_matrices(r0: np.ndarray, r1: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """FarnebackUpdateMatrices -> [H,W,5] float32 (G11, G12, G22, h1, h2)."""
    h, w = flow.shape[:2]
    xs = np.arange(w, dtype=F32)[None, :]
    ys = np.arange(h, dtype=F32)[:, None]
    dx, dy = flow[..., 0], flow[..., 1]
    fx = (xs + dx).astype(F32)
    fy = (ys + dy).astype(F32)
    x1 = np.floor(fx).astype(np.int64)
    y1 = np.floor(fy).astype(np.int64)
    fx = (fx - x1.astype(F32)).astype(F32)
    fy = (fy - y1.astype(F32)).astype(F32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    xc, yc = np.clip(x1, 0, w - 2), np.clip(y1, 0, h - 2)
    one = F32(1)
    a00 = (one - fx) * (one - fy)
    a01 = fx * (one - fy)
    a10 = (one - fx) * fy
    a11 = fx * fy
    p00, p01, p10, p11 = r1[yc, xc], r1[yc, xc + 1], r1[yc + 1, xc], r1[yc + 1, xc + 1]

    def bil(c):
        return ((a00 * p00[..., c] + a01 * p01[..., c]) + a10 * p10[..., c]) + a11 * p11[..., c]

    r2 = np.where(inside, bil(0), F32(0)).astype(F32)
    r3 = np.where(inside, bil(1), F32(0)).astype(F32)
    r4 = np.where(inside, (r0[..., 2] + bil(2)) * F32(0.5), r0[..., 2]).astype(F32)
    r5 = np.where(inside, (r0[..., 3] + bil(3)) * F32(0.5), r0[..., 3]).astype(F32)
    r6 = np.where(inside, (r0[..., 4] + bil(4)) * F32(0.25), r0[..., 4] * F32(0.5)).astype(F32)
    r2 = (r0[..., 0] - r2) * F32(0.5)
    r3 = (r0[..., 1] - r3) * F32(0.5)
    r2 = r2 + (r4 * dy + r6 * dx)
    r3 = r3 + (r6 * dy + r5 * dx)
    nb = len(BORDER)
    xi, yi = np.arange(w), np.arange(h)
    # ((sx_left * sx_right) * sy_top) * sy_bottom, the C expression's association
    sl = np.where(xi < nb, BORDER[np.minimum(xi, nb - 1)], one).astype(F32)[None, :]
    sr = np.where(xi >= w - nb, BORDER[np.clip(w - xi - 1, 0, nb - 1)], one).astype(F32)[None, :]
    st = np.where(yi < nb, BORDER[np.minimum(yi, nb - 1)], one).astype(F32)[:, None]
    sb = np.where(yi >= h - nb, BORDER[np.clip(h - yi - 1, 0, nb - 1)], one).astype(F32)[:, None]
    scale = (((sl * sr) * st) * sb).astype(F32)
    r2, r3, r4, r5, r6 = r2 * scale, r3 * scale, r4 * scale, r5 * scale, r6 * scale
    m = np.empty((h, w, 5), F32)
    m[..., 0] = r4 * r4 + r6 * r6
    m[..., 1] = (r4 + r5) * r6
    m[..., 2] = r5 * r5 + r6 * r6
    m[..., 3] = r4 * r2 + r6 * r3
    m[..., 4] = r6 * r2 + r5 * r3
    return m


def update_flow_blur(m: np.ndarray, winsize: int) -> np.ndarray:
    """FarnebackUpdateFlow_Blur without the matrix update: winsize x winsize box sums of M (edge replicated, double,
    rows r = -m..m summed in that order, then columns), scaled by 1 / winsize^2, 2x2 solve with the 1e-3 regulariser."""
    h, w = m.shape[:2]
    half = winsize // 2
    ys, xs = np.arange(h), np.arange(w)
    v = np.zeros((h, w, 5))
    for r in range(-half, half + 1):
        v = v + m[np.clip(ys + r, 0, h - 1)].astype(np.float64)
    s = np.zeros((h, w, 5))
    for r in range(-half, half + 1):
        s = s + v[:, np.clip(xs + r, 0, w - 1)]
    s = s * (1.0 / (winsize * winsize))
    g11, g12, g22, h1, h2 = (s[..., i] for i in range(5))
    idet = 1.0 / ((g11 * g22 - g12 * g12) + 1e-3)
    flow = np.empty((h, w, 2), F32)
    flow[..., 0] = (g11 * h2 - g12 * h1) * idet
    flow[..., 1] = (g22 * h1 - g12 * h2) * idet
    return flow


def pyramid_levels(h: int, w: int, pyr_scale: float, levels: int):
    """-> [(scale, sigma, smooth_size, width, height)] coarse to fine; levels smaller than 32 pixels are dropped."""
    k, scale = 0, 1.0
    while k < levels:
        scale *= pyr_scale
        if w * scale < 32 or h * scale < 32:
            break
        k += 1
    out = []
    for lv in range(k, -1, -1):
        scale = 1.0
        for _ in range(lv):
            scale *= pyr_scale
        sigma = (1.0 / scale - 1) * 0.5
        smooth = max(cv_round(sigma * 5) | 1, 3)
        out.append((scale, sigma, smooth, cv_round(w * scale), cv_round(h * scale)))
    return out


def calc_optical_flow_farneback(prev: np.ndarray, nxt: np.ndarray, pyr_scale=0.5, levels=5, winsize=11, iterations=5,
                                poly_n=5, poly_sigma=1.1) -> np.ndarray:
    """cv2.calcOpticalFlowFarneback(prev, nxt, None, ..., flags=0) for uint8 [H,W] frames -> float32 [H,W,2]."""
    assert prev.shape == nxt.shape and prev.ndim == 2
    h, w = prev.shape
    flow = None
    for scale, sigma, smooth, lw, lh in pyramid_levels(h, w, pyr_scale, levels):
        if flow is None:
            flow = np.zeros((lh, lw, 2), F32)
        else:
            flow = resize_linear(flow, lh, lw, mul=1.0 / pyr_scale)
        r = []
        for img in (prev, nxt):
            f = gaussian_blur(img.astype(F32), smooth, sigma)
            r.append(poly_exp(resize_linear(f, lh, lw), poly_n, poly_sigma))
        m = update_matrices(r[0], r[1], flow)
        for it in range(iterations):
            flow = update_flow_blur(m, winsize)
            if it < iterations - 1:
                m = update_matrices(r[0], r[1], flow)
    return flow


def bgr2gray(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(uint8 BGR, COLOR_BGR2GRAY): 15-bit fixed point."""
    p = img.astype(np.int64)
    return ((p[..., 0] * 3735 + p[..., 1] * 19235 + p[..., 2] * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def farneback_flow(frames) -> np.ndarray:
    """opticalflow_FarneBack_extractor (train.py:294-332) after decoding -> float32 [len(frames) - 1, h', w', 2]: the
    first frame is resized (longest of its three dimensions -> 224) and then converted to gray, the later ones are
    converted first and resized second; flow between consecutive gray frames."""
    from . import resize as RZ
    factor = 224 / max(frames[0].shape)
    previous = bgr2gray(RZ.resize_linear_u8(frames[0], fx=factor, fy=factor))
    flows = []
    for frame in frames[1:]:
        current = RZ.resize_linear_u8(bgr2gray(frame), fx=factor, fy=factor)
        flows.append(calc_optical_flow_farneback(previous, current))
        previous = current
    return np.asarray(flows)


def farneback_flow_clip(frames, t: int, h: int, w: int) -> np.ndarray:
    """The FarneBack branch of get_twostream_videoclip (train.py:223-239): select_frames over the flow fields, each
    resized to (w, h) with cv2.resize's float bilinear -> float32 [T,h,w,2]."""
    from . import resize as RZ
    flows = farneback_flow(frames)
    return np.asarray([resize_linear(flows[i], h, w) for i in RZ.select_frame_indices(len(flows), t)], F32)
