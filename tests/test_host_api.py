"""Compiles tests/cpp/host_api_test.cpp (the C++ mirror of the Go API over the C ABI) and runs it:
against the CPU oracle here (host logic, error strings, F1-F4 plumbing), against libsonar.so on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp")


def build_and_run(tmp_path, libdir, libname):
    exe = str(tmp_path / "host_api_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, SRC, f"-L{libdir}", f"-l:{libname}",
                           f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    return r


def test_host_api_against_oracle(tmp_path):
    r = build_and_run(tmp_path, os.path.join(ROOT, "oracle"), "libsonar_oracle.so")
    assert r.returncode == 0 and "OK backend=cpu-oracle" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_api_against_cuda_library(tmp_path):
    r = build_and_run(tmp_path, os.path.join(ROOT, "sonido-sonar_b200"), "libsonar.so")
    assert r.returncode == 0 and "OK backend=cuda-sm100a" in r.stdout, r.stdout + r.stderr


def test_xcorr_screen_fft_passes_on_the_cpu(tmp_path):
    """The FP64 Stockham passes and the packed two-sequence correlation of csrc/xcorr_fft.cuh are __host__ __device__:
    tests/cpp/xcorr_fft_selftest.cu replays them on the CPU against a naive DFT and direct correlation sums."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "xcorr_fft_selftest")
    subprocess.check_call([nvcc, "-O2", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "xcorr_fft_selftest.cu")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
