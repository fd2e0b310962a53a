"""Seeded synthetic gray frames (SURVEY.md section 8d "Synthetic inputs").

The reference ships no imagery or generator, so the build defines one: 8-bit gray frames made of
signed Gaussian blobs plus uniform noise, seeded with MT19937(1234 + frame index).  On the
reference this produces >= 1e4 NMS peaks at VGA, so top-4096 selection is saturated.  Frame
*pairs* for the matcher are two crops of one larger blob scene, offset by a known translation.

Pure numpy; used by tests, bench.py and the golden-vector generator alike.
"""
import numpy as np

BASE_SEED = 1234


def _render(rng, H, W, n_blobs, canvas=None):
    img = np.full((H, W), 128.0, dtype=np.float32) if canvas is None else canvas
    cx = rng.uniform(0, W, n_blobs)
    cy = rng.uniform(0, H, n_blobs)
    sig = rng.uniform(2.0, 12.0, n_blobs)
    amp = rng.uniform(40.0, 200.0, n_blobs) * rng.choice([-1.0, 1.0], n_blobs)
    for i in range(n_blobs):
        r = int(np.ceil(3.5 * sig[i]))
        x0, x1 = max(0, int(cx[i]) - r), min(W, int(cx[i]) + r + 1)
        y0, y1 = max(0, int(cy[i]) - r), min(H, int(cy[i]) + r + 1)
        if x0 >= x1 or y0 >= y1:
            continue
        xs = np.arange(x0, x1, dtype=np.float32) - np.float32(cx[i])
        ys = np.arange(y0, y1, dtype=np.float32) - np.float32(cy[i])
        g = np.exp(-(ys[:, None] ** 2 + xs[None, :] ** 2) / np.float32(2.0 * sig[i] ** 2))
        img[y0:y1, x0:x1] += np.float32(amp[i]) * g
    return img


def synthetic_frame(idx, H=480, W=640, n_blobs=None):
    """uint8 [H, W] frame number `idx`."""
    rng = np.random.RandomState(BASE_SEED + int(idx))
    if n_blobs is None:
        n_blobs = max(8, int(round(400 * (H * W) / (480.0 * 640.0))))
    img = _render(rng, H, W, n_blobs)
    img += rng.uniform(-8.0, 8.0, (H, W)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synthetic_frames(start, count, H=480, W=640):
    """uint8 [count, H, W]."""
    return np.stack([synthetic_frame(start + i, H, W) for i in range(count)], axis=0)


def synthetic_pair(idx, H=480, W=640, shift=(12, 7)):
    """Two uint8 [H, W] crops of one scene; frame B shows the scene shifted by `shift` = (dx, dy)
    pixels (a point at (x, y) in A appears at (x - dx, y - dy) in B), with independent noise."""
    dx, dy = shift
    rng = np.random.RandomState(BASE_SEED + 100003 + int(idx))
    Hc, Wc = H + abs(dy), W + abs(dx)
    n_blobs = max(8, int(round(400 * (Hc * Wc) / (480.0 * 640.0))))
    scene = _render(rng, Hc, Wc, n_blobs)
    ax, ay = (0, 0) if dx >= 0 else (-dx, 0)
    if dy < 0:
        ay = -dy
    bx, by = ax + dx, ay + dy
    out = []
    for (ox, oy) in ((ax, ay), (bx, by)):
        crop = scene[oy:oy + H, ox:ox + W].copy()
        crop += rng.uniform(-8.0, 8.0, (H, W)).astype(np.float32)
        out.append(np.clip(np.rint(crop), 0, 255).astype(np.uint8))
    return out[0], out[1]
