"""CPU: the C-ABI boundary.  The library must load without a GPU, export every symbol include/uzliti_edge.h
declares, agree with the header on struct layout, and FAIL LOUDLY (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "uzliti_edge.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from uzliti_slam_b200 import binding
    lib = C.CDLL(binding.lib_path())
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in uzliti_edge.h but not exported"
    assert sorted(binding.EXPORTED_SYMBOLS) == declared       # the Python stub binds exactly the header


def test_header_is_plain_c_and_struct_layout_matches_binding(built, tmp_path):
    from uzliti_slam_b200 import binding
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "uzliti_edge.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(uz_edge_result), offsetof(uz_edge_result, mse),'
                    ' offsetof(uz_edge_result, T), sizeof(uz_params), sizeof(uz_features), offsetof(uz_features, n),'
                    ' offsetof(uz_features, desc_bytes), offsetof(uz_features, sensor_frame));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(binding.EdgeResult), binding.EdgeResult.mse.offset, binding.EdgeResult.T.offset,
                     C.sizeof(binding.Params), C.sizeof(binding.Features), binding.Features.n.offset,
                     binding.Features.desc_bytes.offset, binding.Features.sensor_frame.offset]
    assert sizes[0] == 176 == binding.RESULT_DTYPE.itemsize


def test_default_params_are_the_reference_production_values(built):
    from uzliti_slam_b200 import binding
    lib = binding.load_library()
    p = binding.Params()
    lib.uz_default_params(C.byref(p))
    # iti_slam_launch/yaml/slam.yaml:35-36, cfg/FeatureLinkEstimation.cfg:12, feature_transformation_estimator.cpp:47,67
    assert (p.ransac_threshold, p.ransac_iterations, p.break_percentage) == (0.1, 100, 0.6)
    assert (p.ratio_num, p.ratio_den, p.min_keypoints, p.cross_check, p.do_prosac) == (99, 100, 7, 0, 1)
    assert lib.uz_version().startswith(b"uzliti_edge_b200")


def test_no_cpu_fallback_without_a_gpu(built):
    """On a box without a GPU the product must refuse to run rather than fall back to CPU code."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on the CPU box")
    from uzliti_slam_b200 import EdgeEstimator, UzError
    with pytest.raises(UzError) as e:
        EdgeEstimator(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_references_the_oracle():
    """oracle/ is test infrastructure: nothing under uzliti_slam_b200/ or adapter/ may include, import or load it."""
    offenders = []
    for base in ("uzliti_slam_b200", "adapter", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".c", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(import\s+oracle|from\s+oracle|oracle/|libuz_oracle|uzo_)", txt):
                        offenders.append(os.path.join(dp, f))
    assert offenders == []
