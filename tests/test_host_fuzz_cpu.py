"""Memory safety of the host-side C++ stages (Blosc/LZ4 codec, tensor-text parser, pile-up, candidate extraction, CRC-32C):
tests/tools/host_fuzz.cpp built with AddressSanitizer + UBSan from the stage sources and run on valid and damaged inputs with
a fixed seed.  No GPU, no CUDA; the product library is not involved (the harness compiles the same .cpp files directly)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "clairvoyante_b200", "csrc")
STAGES = ["blosc_frame.cpp", "text_feed.cpp", "pileup.cpp", "candidates.cpp", "crc32c.cpp", "sam_view.cpp", "vcf_text.cpp"]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("host_fuzz") / "host_fuzz")
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fno-omit-frame-pointer", os.path.join(ROOT, "tests", "tools", "host_fuzz.cpp")]
    cmd += [os.path.join(CSRC, f) for f in STAGES] + ["-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stderr[-4000:]
    return exe


@pytest.mark.parametrize("seed", [1, 20261017])
def test_host_stages_under_sanitizers(harness, seed):
    r = subprocess.run([harness, str(seed), "120"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert "iterations clean" in r.stdout
