"""Writes the golden PNG fixtures of tests/golden/png/ with two independent encoders (Pillow's own zlib
encoder with adaptive filters, OpenCV's bundled libpng) plus manifest.json with the shape, dtype and SHA-256
of the pixels those libraries DEcode from them.  Run in the build container (Pillow 12.2, OpenCV 4.13); the files travel with the repo so the tests do
not need either library.  usage: python tools/make_png_golden.py"""
import hashlib
import io
import json
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "png"
OUT.mkdir(parents=True, exist_ok=True)
rng = np.random.default_rng(7)


def image(h, w, ch, kind):
    y, x = np.mgrid[0:h, 0:w]
    if kind == "gradient":
        base = np.stack([(3 * x + y) % 256, (x + 2 * y) % 256, ((x * y) >> 4) % 256, 255 - (x % 64)], -1)
    elif kind == "noise":
        base = rng.integers(0, 256, (h, w, 4))
    else:  # photo-like: smooth + small noise + flat rectangles
        base = np.stack([128 + 100 * np.sin(x / 17.0) * np.cos(y / 23.0), 128 + 90 * np.cos(x / 11.0 + y / 31.0),
                         (x + y) / 2 % 256, np.full((h, w), 255)], -1) + rng.integers(-3, 4, (h, w, 4))
        base[h // 4: h // 2, w // 3: 2 * w // 3] = (10, 200, 30, 255)
    a = np.clip(base, 0, 255).astype(np.uint8)
    return a[..., 0] if ch == 1 else a[..., :ch] if ch != 2 else a[..., [0, 3]]


def main():
    from PIL import Image
    import cv2

    for f in OUT.glob("*"):
        f.unlink()
    manifest = {}

    def expect(name, px):
        px = np.ascontiguousarray(px)
        manifest[name] = {"shape": list(px.shape), "dtype": str(px.dtype), "sha256": hashlib.sha256(px.tobytes()).hexdigest()}

    n = 0
    for h, w in ((1, 1), (3, 5), (17, 33), (48, 64), (90, 257)):
        for ch, mode in ((1, "L"), (2, "LA"), (3, "RGB"), (4, "RGBA")):
            for kind in ("gradient", "noise", "photo"):
                if h * w > 4000 and kind != "photo":
                    continue
                a = image(h, w, ch, kind)
                name = f"{kind}_{mode}_{h}x{w}"
                buf = io.BytesIO()
                Image.fromarray(a, mode).save(buf, format="PNG", compress_level=(n % 9) + 1)
                (OUT / f"pil_{name}.png").write_bytes(buf.getvalue())
                expect(f"pil_{name}.png", np.asarray(Image.open(io.BytesIO(buf.getvalue()))))
                if ch in (1, 3, 4):
                    ok, enc = cv2.imencode(".png", a if ch == 1 else a[..., ::-1] if ch == 3 else a[..., [2, 1, 0, 3]],
                                           [cv2.IMWRITE_PNG_COMPRESSION, n % 10, cv2.IMWRITE_PNG_STRATEGY, n % 5])
                    assert ok
                    (OUT / f"cv_{name}.png").write_bytes(enc.tobytes())
                    dec = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)
                    dec = dec if ch == 1 else dec[..., ::-1] if ch == 3 else dec[..., [2, 1, 0, 3]]
                    expect(f"cv_{name}.png", dec)
                n += 1
    # 16-bit samples (bpp 2 and 6/8) from OpenCV
    for ch in (1, 3):
        a16 = rng.integers(0, 65536, (31, 45) if ch == 1 else (31, 45, 3), dtype=np.uint16)
        ok, enc = cv2.imencode(".png", a16)
        (OUT / f"cv_noise16_{ch}ch_31x45.png").write_bytes(enc.tobytes())
        dec = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)
        expect(f"cv_noise16_{ch}ch_31x45.png", dec if ch == 1 else dec[..., ::-1])  # (OpenCV decodes to BGR)
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=0, sort_keys=True))
    print(len(list(OUT.glob("*.png"))), "PNG files,", sum(f.stat().st_size for f in OUT.iterdir()) // 1024, "KiB in", OUT)


if __name__ == "__main__":
    main()
