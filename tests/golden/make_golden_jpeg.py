"""Golden JPEG fixture: a few small files written by Pillow's encoder and the pixels Pillow's decoder (libjpeg-turbo, the
decoder behind the reference's Image.open, dataset/local_dataset.py:283-286) returns for them. Pins oracle/jpeg.py without
Pillow at test time.   python tests/golden/make_golden_jpeg.py"""
import io
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(7)
out = {}
n = 0
for (h, w, sub, q, kw) in [(24, 40, 2, 75, {}), (37, 53, 2, 90, {}), (16, 24, 1, 60, {}), (17, 9, 0, 85, {}),
                           (32, 48, 2, 80, {"restart_marker_blocks": 2})]:
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3)).astype(np.uint8)
    img = np.asarray(Image.fromarray(base).resize((w, h), Image.BICUBIC)).astype(np.float32)
    img = np.clip(img + rng.normal(0, 10, img.shape), 0, 255).astype(np.uint8)
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", quality=q, subsampling=sub, **kw)
    data = b.getvalue()
    out[f"file{n}"] = np.frombuffer(data, np.uint8)
    out[f"rgb{n}"] = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    n += 1
out["n"] = np.array(n)
np.savez_compressed(os.path.join(HERE, "jpeg.npz"), **out)
print("wrote", n, "files")
