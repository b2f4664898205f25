"""Builds and runs the C++ test of the host-side BatchReader mirror against the CUDA library + oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "tests", "cpp", "test_batch_reader.cc"),
       os.path.join(ROOT, "cnosdb_b200", "csrc", "host", "batch_reader.cc"),
       os.path.join(ROOT, "cnosdb_b200", "csrc", "host", "tsm_writer.cc"),
       os.path.join(ROOT, "cnosdb_b200", "csrc", "host_util.cc")]


def build(tmp_path):
    exe = str(tmp_path / "test_batch_reader")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe] + SRC +
                          ["-L" + os.path.join(ROOT, "cnosdb_b200"), "-ltskv_gpu", "-L" + os.path.join(ROOT, "oracle"),
                           "-ltskv_oracle", "-Wl,-rpath," + os.path.join(ROOT, "cnosdb_b200"),
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-pthread"])
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    assert os.path.exists(build(tmp_path))


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path, engine):
    out = subprocess.run([build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "test_batch_reader ok" in out.stdout
