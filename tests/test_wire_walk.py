"""SURVEY 8f-4, second half: resume.  RosbagStorage keeps one serialised graph_slam_msgs/Node per node
(rosbag_storage.cpp:62-105), loadGraph reads them back (:135-211) and GraphSlamNode::load re-adds every node
(graph_slam_node.cpp:875-888).  The envelope walk finds the Feature[] fields inside such a record on the host; the bulk ingest
decodes all of them in one launch."""
import struct

import numpy as np
import pytest

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ros_wire as W  # noqa: E402
from uzliti_slam_b200 import binding, synthetic as S


def _blob(oracle, cam):
    return oracle.wire_encode(cam)


def test_walk_finds_every_sensor_of_a_node_record(built, oracle):
    fa, fb, _ = S.make_pair(37, seed=1)
    fw, _, _ = S.make_pair(21, seed=2, desc_bytes=64)
    ba, bb, bw = _blob(oracle, fa), _blob(oracle, fb), _blob(oracle, fw)
    sensors = [W.sensor_data(1, "/cam0", ba, 2, gist=(1.0, 2.0, 3.0)),
               W.sensor_data(4, "/laser", struct.pack("<I", 0), 0, ranges=tuple(range(17))),          # a LASERSCAN sensor in between
               W.sensor_data(1, "/cam1_long_frame_name", bb, 2, depth=b"\x01" * 16),
               W.sensor_data(1, "/cam2", bw, 3)]
    msg = W.node("node_0042", sensors)
    node_id, found = binding.wire_walk_node(msg)
    assert node_id == "node_0042" and len(found) == 4
    assert [s.sensor_type for s in found] == [1, 4, 1, 1]
    assert [s.descriptor_type for s in found] == [2, 0, 2, 3]
    assert [s.n_features for s in found] == [37, 0, 37, 21]
    for s, blob, frame in zip(found, (ba, struct.pack("<I", 0), bb, bw), ("/cam0", "/laser", "/cam1_long_frame_name", "/cam2")):
        assert msg[s.features_offset:s.features_offset + s.features_bytes] == bytes(blob)
        assert msg[s.sensor_frame_offset:s.sensor_frame_offset + s.sensor_frame_len].decode() == frame
        assert struct.unpack("<3d", msg[s.displacement_offset:s.displacement_offset + 24]) == (0.1, 0.2, 0.3)
    # one SensorData on its own (what /sensor_data carries), with bytes behind it
    s, used = binding.wire_walk_sensor_data(sensors[2] + b"tail")
    assert used == len(sensors[2]) and s.n_features == 37
    # truncated or corrupt records are refused, never read past the end
    for cut in (3, 40, len(msg) // 2, len(msg) - 1):
        with pytest.raises(binding.UzError):
            binding.wire_walk_node(msg[:cut])
    bad = bytearray(msg)
    bad[found[0].features_offset:found[0].features_offset + 4] = struct.pack("<I", 0x7FFFFFFF)
    with pytest.raises(binding.UzError):
        binding.wire_walk_node(bytes(bad))


@pytest.mark.gpu
def test_bulk_reload_of_a_serialised_map_equals_the_original_store(est, oracle):
    est.clear()
    kn, pairs, _ = S.make_map(90, n_features=300, cluster=9, pool=300, n_shared=180, k_candidates=4, cross_cluster=1, seed=41)
    kw, _, _ = S.make_map(10, n_features=200, cluster=5, pool=200, n_shared=120, k_candidates=2, seed=42, desc_bytes=64)
    kfs = [[k] for k in kn] + [[a, b] for a, b in zip(kw[0::2], kw[1::2])]            # single cameras and two-camera rigs
    h0 = est.add_keyframes(kfs)
    want = est.estimateEdges(h0[pairs[:, 0]], h0[pairs[:, 1]])
    # "shutdown": every node becomes one serialised Node record, written from the store (FeatureData::toMsg on the device)
    records = []
    for i, kf in enumerate(kfs):
        sensors = [W.sensor_data(1, "/cam%d" % c, est.wire_encode(int(h0[i]), c, capacity=400), 3 if cam["desc"].shape[1] == 64 else 2)
                   for c, cam in enumerate(kf)]
        records.append(W.node("n%04d" % i, sensors))
    est.clear()
    # "restart": walk every record on the host, ONE bulk ingest
    frames = {}
    keyframes, ids = [], []
    for rec in records:
        nid, sens = binding.wire_walk_node(rec)
        ids.append(nid)
        arr = np.frombuffer(rec, np.uint8)
        cams = []
        for s in sens:
            if s.sensor_type != 1:
                continue
            frame = rec[s.sensor_frame_offset:s.sensor_frame_offset + s.sensor_frame_len]
            cams.append((arr[s.features_offset:s.features_offset + s.features_bytes], s.descriptor_type, frames.setdefault(frame, len(frames))))
        keyframes.append(cams)
    assert ids[7] == "n0007"
    n0 = est.launch_count()
    h1 = est.add_keyframes_wire(keyframes)
    assert est.launch_count() - n0 <= 4                       # gather + decode + layouts, whatever the number of keyframes
    assert est.store_size() == len(kfs)
    for i in (0, 13, 89, 90, 94):
        for c, cam in enumerate(kfs[i]):
            back = est.read_keyframe(int(h1[i]), c)
            assert np.array_equal(back["desc"], cam["desc"]) and np.array_equal(back["pos"], cam["pos"])
            assert np.array_equal(back["valid"], cam["valid"])
    got = est.estimateEdges(h1[pairs[:, 0]], h1[pairs[:, 1]])
    assert got.tobytes() == want.tobytes()
    # one by one through uz_store_add_wire: the same store
    h2 = est.add_keyframe_wire(bytes(keyframes[3][0][0]), 2, 0)
    a, b = est.read_keyframe(h2), est.read_keyframe(int(h1[3]))
    assert np.array_equal(a["desc"], b["desc"]) and np.array_equal(a["pos"], b["pos"])
    est.clear()
