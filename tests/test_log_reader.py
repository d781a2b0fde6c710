"""include/efusion/Tools/RawLogReader.h against the .klg layout of the reference (Tools/RawLogReader.cpp:22-141): raw and
zlib-compressed depth payloads, the dropped last frame, colour flip, peek-ahead, rewind / fastForward / getBack."""
import os
import struct
import subprocess
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fnv(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _write(path, frames, compress):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(frames)))
        for i, (rgb, depth) in enumerate(frames):
            db = depth.astype("<u2").tobytes()
            if compress and i % 2:
                db = zlib.compress(db)
            ib = rgb.astype(np.uint8).tobytes() if i != 3 else b""  # frame 3 has no image payload -> zeros
            f.write(struct.pack("<qii", 1000 + i * 33333, len(db), len(ib)))
            f.write(db)
            f.write(ib)


def test_raw_log_reader(tmp_path, small_K, small_frames):
    exe = str(tmp_path / "log_reader_check")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{ROOT}/include/efusion", f"-I{ROOT}/include",
                           os.path.join(ROOT, "tests", "cpp", "log_reader_check.cpp"), "-o", exe, "-lz"])
    frames = [(f[0], f[1]) for f in small_frames]
    w, h = small_K.width, small_K.height
    for compress in (False, True):
        klg = str(tmp_path / f"seq{int(compress)}.klg")
        _write(klg, frames, compress)
        for extra in ([], ["peek"], ["flip"], ["peek", "flip"]):
            out = subprocess.check_output([exe, klg, str(w), str(h)] + extra, text=True).strip().split("\n")
            assert out[0] == f"FRAMES {len(frames)}"
            body = [l for l in out[1:] if l[0].isdigit()]
            assert len(body) == len(frames) - 1  # hasMore() never delivers the last frame (RawLogReader.cpp:139-141)
            for k, line in enumerate(body):
                cf, ts, hr, hd = line.split()
                rgb = frames[k][0] if k != 3 else np.zeros_like(frames[k][0])
                if "flip" in extra:
                    rgb = rgb[..., ::-1]
                assert int(cf) == k + 1 and int(ts) == 1000 + k * 33333
                assert int(hr) == _fnv(np.ascontiguousarray(rgb).tobytes()), (compress, extra, k)
                assert int(hd) == _fnv(frames[k][1].astype("<u2").tobytes()), (compress, extra, k)
            ff = [l for l in out if l.startswith("FF ")][0].split()
            bk = [l for l in out if l.startswith("BACK ")][0].split()
            # fastForward(2) skips frames 0 and 1; getNext delivers frame 2; getBack re-delivers it
            assert int(ff[1]) == 3 and int(ff[2]) == 1000 + 2 * 33333 and int(ff[3]) == _fnv(frames[2][1].astype("<u2").tobytes())
            assert int(bk[2]) == int(ff[2]) and int(bk[3]) == int(ff[3])
