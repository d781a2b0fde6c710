"""The source-level drop-in: the reference README's Core-API program (tests/cpp/core_api_example.cpp, written against
`#include <ElasticFusion.h>`) must compile against include/efusion/ + libefusion.so, and on a GPU produce the same pose
and surfel count as the C ABI driven from Python."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "core_api_example.cpp")
CUDA_LIB = "/usr/local/cuda/lib64"


def _compile(out, extra=()):
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", f"-I{ROOT}/include/efusion", f"-I{ROOT}/include", *extra, SRC, "-o", out,
           f"-L{ROOT}/elasticfusion_b200", "-lefusion", f"-Wl,-rpath,{ROOT}/elasticfusion_b200", f"-L{CUDA_LIB}", "-lcudart"]
    subprocess.check_call(cmd)


def test_core_api_example_compiles(tmp_path):
    _compile(str(tmp_path / "example"))


def test_core_api_example_compiles_with_reference_sophus(tmp_path):
    """With the reference's vendored Sophus + Eigen on the include path the pose type is Sophus::SE3d, as in the reference."""
    ref = "/root/reference/third-party"
    if not os.path.isdir(os.path.join(ref, "Sophus")):
        pytest.skip("reference tree not present (GPU box)")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-DEIGEN_MAX_ALIGN_BYTES=0", f"-I{ref}/Sophus", f"-I{ref}/Eigen",
                           f"-I{ROOT}/include/efusion", f"-I{ROOT}/include", SRC])


@pytest.mark.gpu
def test_core_api_example_runs_and_matches_c_abi(tmp_path, small_K, small_frames):
    from elasticfusion_b200 import capi, synth

    exe = str(tmp_path / "example")
    _compile(exe)
    klg = str(tmp_path / "seq.klg")
    synth.write_klg(klg, [(f[0], f[1]) for f in small_frames])
    K = small_K
    out = subprocess.check_output([exe, klg, str(K.width), str(K.height), str(K.fx), str(K.fy), str(K.cx), str(K.cy)], text=True)
    # the same program handing processFrame the next frame as well (look-ahead) must print exactly the same result
    out_la = subprocess.check_output([exe, klg, str(K.width), str(K.height), str(K.fx), str(K.fy), str(K.cx), str(K.cy), "lookahead"], text=True)
    assert out_la == out
    pose = np.array([float(x) for x in out.split("POSE")[1].split("\n")[0].split()]).reshape(4, 4)
    count = int(out.split("COUNT")[1].split()[0])
    tick = int(out.split("TICK")[1].split()[0])
    ctx = capi.Context(capi.default_config(K.width, K.height, K.fx, K.fy, K.cx, K.cy, capacity=500000, time_delta=2147483647 // 2))
    try:
        for i, (rgb, d, _) in enumerate(small_frames):
            ctx.process_frame(rgb, d, i * 33333)
        assert np.abs(ctx.get_pose() - pose).max() < 1e-8
        assert ctx.map_count() == count and ctx.get_tick() == tick == len(small_frames) + 1
    finally:
        ctx.close()
    assert os.path.exists("/tmp/ef_b200_example.freiburg")
    lines = open("/tmp/ef_b200_example.freiburg").read().strip().split("\n")
    assert len(lines) == len(small_frames) and len(lines[0].split()) == 8
