"""Seeded synthetic inputs.  Pure input generation: numpy + cv2's CPU blur.

These are the frames, masks, boxes and descriptor databases every test and bench line uses;
nothing here is on the measured path.
"""
import numpy as np


def frame(w=752, h=480, seed=1):
    """Uniform noise -> Gaussian sigma 1.5 -> min-max to [0,255] -> 60 filled grey rectangles.
    Textured enough that every pyramid level holds more FAST corners than its quota."""
    import cv2
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 256, (h, w)).astype(np.float32)
    a = cv2.GaussianBlur(a, (0, 0), 1.5)
    a = (a - a.min()) / max(float(a.max() - a.min()), 1e-6) * 255.0
    img = a.astype(np.uint8)
    for _ in range(60):
        rw, rh = rs.randint(8, 81, 2)
        x0 = rs.randint(0, max(1, w - rw)); y0 = rs.randint(0, max(1, h - rh))
        img[y0:y0 + rh, x0:x0 + rw] = rs.randint(0, 256)
    return np.ascontiguousarray(img)


def natural_frame(w=752, h=480, seed=1, slope=1.6):
    """A second image family for the parity tests: noise with a 1/f^slope amplitude spectrum (the statistics of
    natural scenes -- large smooth regions, few strong corners, levels that do not reach their quota) plus a few
    dark/bright polygons' worth of sharp edges."""
    rs = np.random.RandomState(seed)
    fy = np.fft.fftfreq(h)[:, None]; fx = np.fft.rfftfreq(w)[None, :]
    amp = 1.0 / np.maximum(np.hypot(fx, fy), 1.0 / max(w, h)) ** slope
    spec = amp * np.exp(2j * np.pi * rs.rand(h, fx.shape[1]))
    a = np.fft.irfft2(spec, s=(h, w))
    lo, hi = np.percentile(a, (1, 99))
    img = np.clip((a - lo) / max(hi - lo, 1e-9) * 255.0, 0, 255).astype(np.uint8)
    for _ in range(12):
        rw, rh = rs.randint(20, 120, 2)
        x0 = rs.randint(0, max(1, w - rw)); y0 = rs.randint(0, max(1, h - rh))
        img[y0:y0 + rh, x0:x0 + rw] = np.clip(img[y0:y0 + rh, x0:x0 + rw].astype(np.int32) + rs.randint(-90, 91), 0, 255)
    return np.ascontiguousarray(img)


def shifted(img, dx=3, dy=2, sigma=2.0, seed=2):
    """Second frame of a pair: integer translation (REFLECT_101 fill) plus Gaussian noise."""
    h, w = img.shape
    ys = np.arange(h) - dy; xs = np.arange(w) - dx

    def refl(i, n):
        i = np.abs(i)
        return np.where(i >= n, 2 * n - 2 - i, i)
    out = img[refl(ys, h)][:, refl(xs, w)].astype(np.float32)
    out += np.random.RandomState(seed).normal(0.0, sigma, out.shape).astype(np.float32)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def frame_stream(n, w=752, h=480, seed0=1000):
    """n frames: even indices are fresh textures, odd ones the shifted+noisy partner, so
    consecutive-frame matching has true correspondences (config C2)."""
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        if i % 2 == 0:
            out[i] = frame(w, h, seed0 + i)
        else:
            out[i] = shifted(out[i - 1], seed=seed0 + i)
    return out


def mask_and_boxes(w, h, seed=3000, n=6):
    """C3: n random rectangles zeroed in a 0/255 mask, and the same rectangles as detector boxes
    (x1,y1,x2,y2,confidence,class_id) with dynamic class ids."""
    rs = np.random.RandomState(seed)
    m = np.full((h, w), 255, np.uint8)
    boxes = np.zeros((n, 6), np.float32)
    for i in range(n):
        bw, bh = rs.randint(w // 16, w // 4), rs.randint(h // 16, h // 4)
        x0 = rs.randint(0, w - bw); y0 = rs.randint(0, h - bh)
        m[y0:y0 + bh, x0:x0 + bw] = 0
        boxes[i] = (x0, y0, x0 + bw - 1, y0 + bh - 1, 0.9, (0, 1, 2, 3, 5, 7)[i % 6])
    return m, boxes


def descriptor_db(rows=1_000_000, seed=5):
    return np.random.RandomState(seed).randint(0, 256, (rows, 32), dtype=np.uint8)


def loop_queries(db, nq=2000, seed=6, max_flips=40, lo=0, hi=None):
    """Queries = database rows from [lo,hi) with 0..max_flips random bit flips; every 16th query is
    an exact duplicate of its predecessor's source row so (distance, index) ties occur."""
    rs = np.random.RandomState(seed)
    hi = len(db) if hi is None else hi
    src = rs.randint(lo, hi, nq)
    src[15::16] = src[14::16][:len(src[15::16])]
    q = db[src].copy()
    for i in range(nq):
        k = 0 if i % 16 >= 14 else rs.randint(0, max_flips + 1)
        if k:
            bits = rs.choice(256, k, replace=False)
            np.bitwise_xor.at(q[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    return q, src
