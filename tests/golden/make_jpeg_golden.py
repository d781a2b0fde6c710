"""Generates tests/golden/jpeg_golden.npz: JPEG streams (baseline, the subsamplings / sizes / restart intervals the decoder
supports) and their decode by libjpeg-turbo with default settings (Pillow: JDCT_ISLOW, fancy upsampling) — the library the
reference's Tools/JPEGLoader.h calls. Run in the build container (Pillow + OpenCV present):  python tests/golden/make_jpeg_golden.py"""
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import cv2
    from PIL import Image

    from elasticfusion_b200 import synth

    K = synth.Intrinsics(160, 120, 132.0, 132.0, 80.0, 60.0)
    rgb = synth.render(synth.trajectory(1)[0], K)[0]
    out = {}
    k = 0

    def add(data, note):
        nonlocal k
        ref = np.array(Image.open(io.BytesIO(data)).convert("RGB"))
        out[f"jpg_{k}"] = np.frombuffer(data, np.uint8)
        out[f"rgb_{k}"] = ref
        out[f"note_{k}"] = np.array(note)
        k += 1

    for (w, h), sub, q in (((160, 120), 2, 90), ((160, 120), 0, 75), ((157, 113), 1, 60), ((157, 113), 2, 95), ((33, 17), 2, 50),
                           ((17, 1), 2, 90), ((1, 1), 1, 90)):
        b = io.BytesIO()
        Image.fromarray(np.ascontiguousarray(rgb[:h, :w])).save(b, "JPEG", quality=q, subsampling=sub)
        add(b.getvalue(), f"pillow {w}x{h} subsampling={sub} q={q}")
    ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(rgb[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 4])
    add(enc.tobytes(), "opencv BGR 160x120 4:2:0 q=90 restart interval 4 (what the reference's Logger writes, plus DRI)")
    b = io.BytesIO()
    Image.fromarray(rgb).convert("L").save(b, "JPEG", quality=85)
    add(b.getvalue(), "pillow greyscale 160x120 q=85")
    out["count"] = np.array(k)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "jpeg_golden.npz"), **out)
    print("wrote", k, "streams")


if __name__ == "__main__":
    main()
