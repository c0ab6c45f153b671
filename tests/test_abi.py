"""The C-ABI boundary: every symbol include/sonar.h declares is exported by both implementations,
struct layouts in the ctypes binding match the C compiler's, and the product library refuses to run
without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sonar.h")
PRODUCT = os.path.join(ROOT, "sonido-sonar_b200", "libsonar.so")
ORACLE = os.path.join(ROOT, "oracle", "libsonar_oracle.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sonar_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(capi):
    assert declared_symbols() == sorted(capi.EXPORTS)


@pytest.mark.parametrize("path", [PRODUCT, ORACLE])
def test_every_declared_symbol_is_exported(path):
    if not os.path.exists(path):
        pytest.fail(f"{path} not built: run python -c 'import __graft_entry__ as g; g.build()'")
    lib = C.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.sonar_abi_version.restype = C.c_int
    assert lib.sonar_abi_version() == 2
    lib.sonar_backend.restype = C.c_char_p
    assert lib.sonar_backend() == (b"cuda-sm100a" if path == PRODUCT else b"cpu-oracle")


def test_struct_layouts_match_the_c_compiler(capi, tmp_path):
    structs = {
        "sonar_fp_params": capi.FpParams, "sonar_fp_sizes_t": capi.FpSizes, "sonar_fp_out": capi.FpOut,
        "sonar_fp_dev_layout_t": capi.FpDevLayout, "sonar_xcorr_summary": capi.XcorrSummary,
        "sonar_xcorr_shard_peak": capi.XcorrShardPeak, "sonar_xcorr_shard_metrics": capi.XcorrShardMetrics,
        "sonar_align_result": capi.AlignResult, "sonar_dtw_out": capi.DtwOut, "sonar_cmp_features": capi.CmpFeatures,
        "sonar_cmp_weights": capi.CmpWeights, "sonar_cmp_result": capi.CmpResult,
        "sonar_kernel_time": capi.KernelTime, "sonar_pair_out": capi.PairOut,
    }
    body = "\n".join(f'  printf("{n} %zu\\n", sizeof({n}));' for n in structs)
    src = tmp_path / "sz.c"
    src.write_text(f'#include <stdio.h>\n#include "{HEADER}"\nint main(void) {{\n{body}\n  return 0;\n}}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-o", str(exe), str(src)])  # the header must be plain C
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n, t in structs.items():
        assert int(out[n]) == C.sizeof(t), n


def test_host_only_entry_points_work_without_a_context(capi):
    """Size arithmetic / window tables / shard merges are pure host code in the product library."""
    lib = capi.SonarLib(PRODUCT, init=False)
    p = lib.default_params()
    s = lib.fp_sizes(p, 1323000)
    assert (s.n_frames, s.n_bins, s.n_flux, s.n_pitch_frames) == (5164, 513, 5163, 2582)
    w = lib.window("hann", 1024)
    assert w[0] == 0.0 and abs((w * w).sum() / 1024 - 1.0) < 1e-12
    pk = [capi.XcorrShardPeak(0.5, 10), capi.XcorrShardPeak(0.75, 40), capi.XcorrShardPeak(0.75, 30)]
    assert lib.xcorr_merge_peaks(pk) == 30  # larger |c| wins, ties -> smaller index


def test_no_cpu_fallback_without_a_device(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.SonarError) as e:
        capi.SonarLib(PRODUCT)
    assert e.value.code == capi.ERR_CUDA and "no CPU fallback" in e.value.msg


def test_product_path_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may touch oracle/."""
    pkg_dir = os.path.join(ROOT, "sonido-sonar_b200")
    offenders = []
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".go")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"libsonar_oracle|sonar_oracle\.cpp|oracle/", text) and f != "capi.py":
                    offenders.append(f)
    assert not offenders, offenders
    # capi.py only mentions the oracle in its docstring; it never loads it by itself
    text = open(os.path.join(pkg_dir, "capi.py")).read()
    assert "PRODUCT_LIB = os.path.join(_HERE, \"libsonar.so\")" in text
