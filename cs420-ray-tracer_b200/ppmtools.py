"""PPM helpers for tests and bench: reader, P3 text and the reference's comparison rule."""
import numpy as np


def read_ppm(path):
    """Returns (W, H, maxval, rgb[H, W, 3] uint8/uint16 in FILE row order, i.e. top row first)."""
    with open(path, "rb") as f:
        data = f.read()
    # header tokens, '#' comments allowed between them (scripts/compare_ppm.py:10-27)
    pos = 0
    toks = []
    while len(toks) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] not in (b"\n", b""):
                pos += 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        toks.append(data[pos:end])
        pos = end
    magic, W, H, maxval = toks[0].decode(), int(toks[1]), int(toks[2]), int(toks[3])
    if magic == "P3":
        vals = np.array(data[pos:].split(), dtype=np.int64)
    elif magic == "P6":
        pos += 1
        vals = np.frombuffer(data[pos:], dtype=np.uint8 if maxval < 256 else ">u2").astype(np.int64)
    else:
        raise ValueError("Unsupported PPM format: %s" % magic)
    if vals.size != W * H * 3:
        raise ValueError("pixel count mismatch: %d vs %d" % (vals.size, W * H * 3))
    return W, H, maxval, vals.reshape(H, W, 3)


def ppm_text(rgb_bottom_first):
    """P3 text exactly as src/main.cpp:69-91 writes it; input is [H, W, 3] with row 0 = bottom."""
    a = np.asarray(rgb_bottom_first)
    H, W, _ = a.shape
    flipped = a[::-1].reshape(-1, 3)
    body = "\n".join("%d %d %d" % (r, g, b) for r, g, b in flipped.tolist())
    return "P3\n%d %d\n255\n%s\n" % (W, H, body)


def compare_rgb(a, b, tolerance_percent=0.5, maxval=255):
    """The rule of scripts/compare_ppm.py:50-93 on arrays: counts CHANNEL SAMPLES whose absolute
    difference exceeds int(maxval*tol/100) (0.5 -> 1 LSB); match iff < 0.1 % of samples do.
    Returns (match, diff_percent, max_abs_diff)."""
    a = np.asarray(a, dtype=np.int64).ravel()
    b = np.asarray(b, dtype=np.int64).ravel()
    if a.size != b.size:
        return False, 100.0, -1
    tol = int(maxval * tolerance_percent / 100.0)
    d = np.abs(a - b)
    pct = float((d > tol).sum()) / a.size * 100.0
    return pct < 0.1, pct, int(d.max()) if d.size else 0
