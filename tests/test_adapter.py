"""The C++ host-side mirror of the reference interface (adapter/): builds on CPU, runs on the GPU box."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "adapter")


def _build():
    subprocess.check_call(["make", "-C", ADAPTER, "CXX=g++"], stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ADAPTER, "test_adapter")) or not os.path.exists(os.path.join(ADAPTER, "libuz_adapter.so")):
        subprocess.check_call(["make", "-C", ADAPTER, "CXX=g++"])


def test_adapter_builds_and_exposes_the_reference_interface(built):
    _build()
    syms = subprocess.check_output(["nm", "-DC", os.path.join(ADAPTER, "libuz_adapter.so")], text=True).replace("[abi:cxx11]", "")
    for name in ("TransformationEstimator::estimateEdge(SlamNode&, SlamNode&)",
                 "GpuFeatureTransformationEstimator::estimateEdgeImpl(SlamNode&, SlamNode&, SlamEdge&)",
                 "GpuFeatureTransformationEstimator::estimateEdgeDirect(",
                 "GpuFeatureTransformationEstimator::estimateSVD(",
                 "GpuFeatureTransformationEstimator::consensus3D(",
                 "GpuFeatureTransformationEstimator::setConfig(",
                 "GpuFeatureTransformationEstimator::estimateEdgeBatch(",
                 "GpuFeatureTransformationEstimator::acceptEdges(",
                 "GpuFeatureTransformationEstimator::estimateSVDBatch(",
                 "GpuLshSetRecognizer::addNode(SlamNode const&)",
                 "GpuLshSetRecognizer::searchAndAddPlace(SlamNode const&)",
                 "GpuLshSetRecognizer::searchPlace(SlamNode const&)",
                 "GpuLshSetRecognizer::addPlace(SlamNode const&)",
                 "GpuLshSetRecognizer::removePlace(",
                 "GpuLshSetRecognizer::recognizedPlaces()",
                 "GpuLshSetRecognizer::hasRecognizedPlaces()",
                 "GpuLshSetRecognizer::setConfig(",
                 "GpuLshSetRecognizer::clear()"):
        assert name in syms, name
    # the adapter reaches the GPU only through the C-ABI
    undefined = subprocess.check_output(["nm", "-Du", os.path.join(ADAPTER, "libuz_adapter.so")], text=True)
    assert "uz_estimate_edges" in undefined and "cuda" not in undefined.lower()


@pytest.mark.gpu
def test_adapter_end_to_end_on_gpu(built):
    _build()
    out = subprocess.run([os.path.join(ADAPTER, "test_adapter")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ADAPTER OK" in out.stdout


@pytest.mark.gpu
def test_adapter_on_two_gpus(built):
    """the same C++ class over a device list (uz_group behind the adapter): same checks, edges identical to one device"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _build()
    env = dict(os.environ, UZ_TEST_DEVICES="0,1")
    out = subprocess.run([os.path.join(ADAPTER, "test_adapter")], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ADAPTER OK" in out.stdout and "2 device(s)" in out.stdout


@pytest.mark.gpu
def test_adapter_bench_on_the_benchmark_map(built):
    """bench_adapter generates the benchmark's map in C++ (include/uz_synth.h) and drives it through the adapter's queue:
    the map must be the one the Python generator builds, every edge must come back, and resident == cold answers"""
    import json
    from uzliti_slam_b200 import synth_splitmix as SM
    _build()
    out = subprocess.run([os.path.join(ADAPTER, "bench_adapter"), "150", "2000", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    kfs, pairs, _ = SM.make_map(150, native=False)
    assert d["map_checksum_desc"] == SM.checksum(np.stack([k["desc"] for k in kfs]))
    assert d["map_checksum_pos"] == SM.checksum(np.stack([k["pos"] for k in kfs]))
    assert d["pairs"] == 2000 and d["same_edges"] is True and d["edges_score_ge_15"] > 1000
    # the same pairs through the Python binding: the sum of matching scores agrees (scores are integers)
    from uzliti_slam_b200 import EdgeEstimator
    est = EdgeEstimator(0)
    h = est.add_keyframes(kfs)
    res = est.estimateEdges(h[pairs[:2000, 0]], h[pairs[:2000, 1]])
    est.close()
    assert int(np.where(res["ok"] != 0, res["consensus"], 0).sum()) == d["score_sum_resident"]
