"""The C++ row shuttle (avif-format_b200/host/GpuRowShuttle.cpp) on a machine WITHOUT a GPU: the same native test program
as tests/test_gpu_host_shuttle.py -- mock Photoshop host, mock libheif, every CreateHeifImage* / ReadHeifImage* entry point,
multi-row advanceState blocks, staging budget, row-transform seam, colour-profile guard, error mapping -- linked against
tests/native/fake_avifgpu_on_oracle.cpp, which stands in for the GPU-touching C-ABI calls and converts with the CPU oracle.
What this pins is the shuttle's own logic (which rows go where, in which blocks, with which strides, and what it throws);
the conversions themselves are the GPU tests' business."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = tmp_path / "host_shuttle_test_cpu"
    pkg = os.path.join(ROOT, "avif-format_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(pkg, "host"), "-I", os.path.join(ROOT, "oracle", "shim"),
           os.path.join(ROOT, "tests", "native", "host_shuttle_test.cpp"), os.path.join(ROOT, "tests", "native", "fake_avifgpu_on_oracle.cpp"),
           os.path.join(pkg, "host", "GpuRowShuttle.cpp"), os.path.join(ROOT, "oracle", "shim", "mock_heif.cpp"),
           os.path.join(pkg, "lib", "libavifgpu.so"), os.path.join(ROOT, "oracle", "liboracle.so"),
           "-Wl,-rpath," + os.path.join(pkg, "lib"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lpthread", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return str(exe)


def test_host_shuttle_logic_on_the_cpu(tmp_path):
    out = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=600)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL PASSED" in out.stdout and "FAIL" not in out.stdout
