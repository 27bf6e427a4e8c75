"""The C++ host mirror of the plug-in's row shuttle (avif-format_b200/host) against a mock Photoshop host and a mock
libheif, on a B200: same entry points as the reference (CreateHeifImage* / ReadHeifImage*), multi-row advanceState
blocks, exceptions mapped like the reference's -- results compared with the CPU oracle inside the C++ test."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = tmp_path / "host_shuttle_test"
    pkg = os.path.join(ROOT, "avif-format_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(pkg, "host"), "-I", os.path.join(ROOT, "oracle", "shim"),
           os.path.join(ROOT, "tests", "native", "host_shuttle_test.cpp"), os.path.join(pkg, "host", "GpuRowShuttle.cpp"),
           os.path.join(ROOT, "oracle", "shim", "mock_heif.cpp"),
           os.path.join(pkg, "lib", "libavifgpu.so"), os.path.join(ROOT, "oracle", "liboracle.so"),
           "-Wl,-rpath," + os.path.join(pkg, "lib"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lpthread", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return str(exe)


def test_host_shuttle_compiles_and_links(tmp_path):
    build(tmp_path)


@pytest.mark.gpu
def test_host_shuttle_matches_oracle(tmp_path):
    out = subprocess.run([build(tmp_path)], capture_output=True, text=True)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL PASSED" in out.stdout and "FAIL" not in out.stdout
