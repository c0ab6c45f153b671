"""The Go shim under go/ cannot be compiled here (no Go toolchain), so its calls into the C ABI are checked textually:
every C.sonar_* function it calls must be declared in include/sonar.h with the same number of arguments, and every C type
it names must be a type of the header."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    text = open(os.path.join(ROOT, "include", "sonar.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"\b(sonar_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_args(args))
    types = set(re.findall(r"\}\s*(sonar_[a-z0-9_]+)\s*;", text)) | set(re.findall(r"typedef\s+struct\s+\w+\s+(sonar_[a-z0-9_]+)\s*;", text))
    return protos, types


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def _go_calls():
    calls = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "go")):
        for f in files:
            if not f.endswith(".go"):
                continue
            src = open(os.path.join(dirpath, f)).read()
            src = re.sub(r"//[^\n]*", "", src)
            for m in re.finditer(r"C\.(sonar_[a-z0-9_]+)\s*(\()?", src):
                name = m.group(1)
                if not m.group(2):
                    calls.append((f, name, None))
                    continue
                i, depth = m.end(), 1
                while depth and i < len(src):
                    depth += src[i] in "([{"
                    depth -= src[i] in ")]}"
                    i += 1
                inner = src[m.end():i - 1]
                calls.append((f, name, 0 if not inner.strip() else len(_split_args(inner))))
    return calls


def test_go_shim_calls_match_the_header():
    protos, types = _header()
    assert len(protos) >= 50, "header parse lost the prototypes"
    calls = _go_calls()
    assert len(calls) > 30
    seen = 0
    for f, name, nargs in calls:
        if name in protos and nargs is not None:
            # a struct type and a function never share a name; C.type(x) conversions have one argument and are types
            assert nargs == protos[name], f"{f}: C.{name} called with {nargs} arguments, header declares {protos[name]}"
            seen += 1
        elif name in protos:
            pytest.fail(f"{f}: C.{name} is a function but is used without a call")
        else:
            assert name in types, f"{f}: C.{name} is neither a function nor a type of include/sonar.h"
    assert seen >= 20


def test_go_shim_binds_the_three_public_entry_points():
    """GenerateFingerprint, ExtractAlignmentFeatures and Compare (BASELINE north star) all reach the C ABI."""
    names = {n for _, n, _ in _go_calls()}
    for need in ("sonar_fingerprint_f64", "sonar_align_xcorr_f64", "sonar_dtw_f64", "sonar_compare_f64", "sonar_compare_batch_f64"):
        assert need in names, need
