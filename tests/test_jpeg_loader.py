"""include/efusion/Tools/JPEGLoader.h (own baseline JPEG decoder behind the reference's JPEGLoader interface,
Tools/JPEGLoader.h:32-91) against libjpeg-turbo's default decode: byte-identical on the committed golden streams
(tests/golden/jpeg_golden.npz, written by make_jpeg_golden.py) and, when Pillow is importable, on freshly encoded ones;
plus a .klg whose colour payloads are JPEG and depth payloads zlib (what the reference's Logger records) through RawLogReader."""
import io
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def jpeg_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("jpeg") / "jpeg_check")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{ROOT}/include/efusion", os.path.join(ROOT, "tests", "cpp", "jpeg_check.cpp"), "-o", exe])
    return exe


def _decode(exe, data, tmp_path):
    p = str(tmp_path / "x.jpg")
    open(p, "wb").write(data)
    r = subprocess.run([exe, p], capture_output=True)
    if r.returncode:
        return None, r.stderr.decode()
    hdr, _, raw = r.stdout.partition(b"\n")
    w, h = map(int, hdr.split())
    return np.frombuffer(raw, np.uint8).reshape(h, w, 3), None


def test_golden_streams_decode_bit_exactly(jpeg_check, tmp_path):
    z = np.load(os.path.join(ROOT, "tests", "golden", "jpeg_golden.npz"))
    n = int(z["count"])
    assert n >= 8
    for k in range(n):
        got, err = _decode(jpeg_check, z[f"jpg_{k}"].tobytes(), tmp_path)
        assert got is not None, (str(z[f"note_{k}"]), err)
        assert np.array_equal(got, z[f"rgb_{k}"]), str(z[f"note_{k}"])


def test_fresh_streams_match_pillow(jpeg_check, tmp_path, small_frames):
    Image = pytest.importorskip("PIL.Image")
    rgb = small_frames[1][0]
    for (w, h), sub, q in (((160, 120), 2, 30), ((159, 119), 2, 100), ((96, 40), 1, 85), ((41, 77), 0, 70)):
        b = io.BytesIO()
        Image.fromarray(np.ascontiguousarray(rgb[:h, :w])).save(b, "JPEG", quality=q, subsampling=sub)
        ref = np.array(Image.open(io.BytesIO(b.getvalue())).convert("RGB"))
        got, err = _decode(jpeg_check, b.getvalue(), tmp_path)
        assert got is not None, err
        assert np.array_equal(got, ref), (w, h, sub, q)
    b = io.BytesIO()
    Image.fromarray(rgb).save(b, "JPEG", quality=85, progressive=True)
    got, err = _decode(jpeg_check, b.getvalue(), tmp_path)
    assert got is None and "progressive" in err  # refused with an actionable message, not decoded wrongly


def _fnv(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_klg_with_jpeg_and_zlib_payloads(tmp_path, small_K, small_frames):
    """The layout the reference's Logger writes (RawLogReader.cpp:80-97): zlib depth + JPEG colour. The reader must deliver the
    libjpeg decode with channels 0 and 2 swapped (JPEGLoader.h:73-82), then flipColors on top when asked."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "jpeg_golden.npz"))
    jpgs = [(z[f"jpg_{k}"].tobytes(), z[f"rgb_{k}"]) for k in range(int(z["count"])) if z[f"rgb_{k}"].shape == (small_K.height, small_K.width, 3)]
    assert len(jpgs) >= 3
    frames = [(jpgs[i % len(jpgs)], small_frames[i % len(small_frames)][1]) for i in range(5)]
    klg = str(tmp_path / "jz.klg")
    with open(klg, "wb") as f:
        f.write(struct.pack("<i", len(frames)))
        for i, ((jb, _), depth) in enumerate(frames):
            db = zlib.compress(depth.astype("<u2").tobytes())
            f.write(struct.pack("<qii", 1000 + i * 33333, len(db), len(jb)))
            f.write(db)
            f.write(jb)
    exe = str(tmp_path / "log_reader_check")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{ROOT}/include/efusion", f"-I{ROOT}/include",
                           os.path.join(ROOT, "tests", "cpp", "log_reader_check.cpp"), "-o", exe, "-lz"])
    for extra in ([], ["flip"], ["peek"]):
        out = subprocess.check_output([exe, klg, str(small_K.width), str(small_K.height)] + extra, text=True).strip().split("\n")
        body = [l for l in out[1:] if l[0].isdigit()]
        assert len(body) == len(frames) - 1
        for k, line in enumerate(body):
            _, ts, hr, hd = line.split()
            rgb = frames[k][0][1][..., ::-1]  # decoder output with channels 0 and 2 swapped
            if "flip" in extra:
                rgb = rgb[..., ::-1]
            assert int(hr) == _fnv(np.ascontiguousarray(rgb).tobytes()), (extra, k)
            assert int(hd) == _fnv(frames[k][1].astype("<u2").tobytes()), (extra, k)
