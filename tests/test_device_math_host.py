"""Host-side gate for the device libm (avif-format_b200/csrc/device_math.cuh): the very header the CUDA kernels
compile is built for the CPU and compared with the system libm -- powf at the exponents the path uses (and a few
generic ones, negative bases, NaN / inf / subnormal inputs), expf and logf -- over a dense sweep of all floats.
0 mismatches is required: the GPU's float outputs are only as exact as this replica."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker_binary(tmp_path_factory):
    exe = tmp_path_factory.mktemp("libm") / "libm_replica_check"
    subprocess.run(["g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-I", os.path.join(ROOT, "avif-format_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "libm_replica_check.cpp"), "-o", str(exe), "-lpthread"], check=True)
    return str(exe)


def test_device_libm_source_matches_system_libm(checker_binary):
    # stride 97: every 97th float of each sign, ~22 M comparisons per function and exponent
    out = subprocess.run([checker_binary, "97", str(min(os.cpu_count() or 1, 16))], capture_output=True, text=True)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 26
    for line in lines:
        assert "mismatched=0 " in line + " ", line


def test_tables_are_the_installed_libms(tmp_path):
    """libm_tables.inc is generated from the installed libm: regenerating must reproduce the committed file."""
    committed = open(os.path.join(ROOT, "avif-format_b200", "csrc", "libm_tables.inc")).read()
    script = os.path.join(ROOT, "avif-format_b200", "tools", "gen_libm_tables.py")
    text = open(script).read().replace('os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "libm_tables.inc")',
                                       repr(str(tmp_path / "libm_tables.inc")))
    patched = tmp_path / "gen.py"
    patched.write_text(text)
    try:
        subprocess.run(["python", str(patched)], check=True, capture_output=True)
    except (subprocess.CalledProcessError, StopIteration):
        pytest.skip("libm.so.6 of this platform does not expose the expected tables")
    regenerated = open(tmp_path / "libm_tables.inc").read()
    strip = lambda s: "\n".join(s.splitlines()[1:])  # noqa: E731  (first line names the source path)
    assert strip(regenerated) == strip(committed)
