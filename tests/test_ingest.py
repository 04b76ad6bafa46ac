"""8f-3 / 8f-4, the data formats in front of the store: extract3dFeatures' back-projection
(feature_extraction_core.cpp:254-295) and FeatureData::fromMsg's decode of the serialised graph_slam_msgs/Feature[]
(sensor_data.cpp:124-171), both on the device, against the oracle's literal restatements."""
import struct

import numpy as np
import pytest

from uzliti_slam_b200 import synthetic as S


def encode_features(desc_f32, u, v, is3d, strength, pos):
    """ROS1 serialisation of graph_slam_msgs/Feature[] (Feature.msg field order; little endian; no padding)."""
    out = [struct.pack("<I", len(u))]
    for i in range(len(u)):
        out.append(struct.pack("<iiBfI", int(u[i]), int(v[i]), int(bool(is3d[i])), float(strength[i]), desc_f32.shape[1]))
        out.append(np.asarray(desc_f32[i], "<f4").tobytes())
        out.append(struct.pack("<ddd", *[float(x) for x in pos[i]]))
    return b"".join(out)


def _depth_image(rng, h=480, w=640):
    d = rng.uniform(0.4, 9.0, (h, w)).astype(np.float32)
    d[rng.random((h, w)) < 0.1] = 0.0
    d[rng.random((h, w)) < 0.05] = np.nan
    return d


def test_oracle_backproject_semantics(oracle):
    rng = np.random.default_rng(0)
    d = _depth_image(rng)
    u = np.array([0, 639, 700, -5, 320, 100], np.int32)
    v = np.array([0, 479, 10, 500, 240, 100], np.int32)
    d[0, 0] = 2.0; d[479, 639] = 0.0; d[10, 639] = 7.0; d[479, 0] = 7.5; d[240, 320] = np.nan; d[100, 100] = 3.0
    pos, valid = oracle.backproject(u, v, d)
    assert valid.tolist() == [1, 0, 1, 0, 0, 1]
    assert np.array_equal(pos[1], [0, 0, -1]) and np.array_equal(pos[3], [0, 0, -1])       # zero depth / beyond max_depth
    assert pos[2, 2] == 7.0 and pos[2, 0] == (639 - 319.5) * 7.0 / 525.0                   # clamped into the image
    assert pos[5, 1] == (100 - 239.5) * 3.0 / 525.0
    rpos, rvalid = oracle.backproject(u, v, d, reverse=True)                                # the reference's own order
    assert np.array_equal(rpos, pos[::-1]) and np.array_equal(rvalid, valid[::-1])
    assert oracle.backproject(u, v, d, max_depth=0.0)[1].tolist() == [1, 0, 1, 1, 0, 1]     # 0 = no limit


def test_oracle_wire_decode(oracle):
    f, _, _ = S.make_pair(50, seed=1)
    rng = np.random.default_rng(2)
    u = rng.integers(0, 640, 50); v = rng.integers(0, 480, 50)
    blob = encode_features(f["desc"].astype(np.float32), u, v, f["valid"], np.full(50, -1.0), f["pos"])
    assert len(blob) == 4 + 50 * 169
    desc, pos, valid, uv = oracle.wire_decode(blob)
    assert np.array_equal(desc, f["desc"]) and np.array_equal(pos, f["pos"]) and np.array_equal(valid, f["valid"])
    assert np.array_equal(uv[:, 0], u) and np.array_equal(uv[:, 1], v)
    assert oracle.wire_decode(blob[:-1]) is None                                            # truncated
    weird = f["desc"].astype(np.float32)
    weird[0, :6] = [255.9, 256.0, -1.0, 1e10, np.nan, 3.7]                                  # the (unsigned char) cast of gcc/x86
    desc, _, _, _ = oracle.wire_decode(encode_features(weird, u, v, f["valid"], np.zeros(50), f["pos"]))
    assert desc[0, :6].tolist() == [255, 0, 255, 0, 0, 3]


@pytest.mark.gpu
@pytest.mark.parametrize("reverse", [False, True])
def test_gpu_backproject(est, oracle, reverse):
    rng = np.random.default_rng(3)
    d = _depth_image(rng)
    n = 3000
    u = rng.integers(-20, 680, n).astype(np.int32)
    v = rng.integers(-20, 520, n).astype(np.int32)
    for md in (7.0, 0.0):
        gp, gv = est.backproject(u, v, d, max_depth=md, reverse=reverse)
        op, ov = oracle.backproject(u, v, d, max_depth=md, reverse=reverse)
        assert np.array_equal(gv, ov) and gp.tobytes() == op.tobytes()
    padded = np.zeros((480, 704), np.float32)                                               # row-padded image
    padded[:, :640] = d
    gp, gv = est.backproject(u, v, padded[:, :640], reverse=reverse)
    op, ov = oracle.backproject(u, v, d, reverse=reverse)
    assert np.array_equal(gv, ov) and gp.tobytes() == op.tobytes()


@pytest.mark.gpu
def test_gpu_rgbd_ingest_feeds_the_path(est, oracle):
    """pixels + depth image -> store -> edge: same result as the host-side back-projection fed through uz_store_add"""
    rng = np.random.default_rng(4)
    est.clear()
    kf = []
    for s in range(2):
        d = _depth_image(rng)
        n = 700
        u = rng.integers(0, 640, n).astype(np.int32); v = rng.integers(0, 480, n).astype(np.int32)
        desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        kf.append((desc, u, v, d))
    kf[1] = (kf[0][0].copy(), kf[0][1], kf[0][2], kf[0][3])          # same scene twice: identity edge
    for reverse in (False, True):
        hs = [est.add_keyframe_rgbd(desc, u, v, d, reverse=reverse) for desc, u, v, d in kf]
        back = est.read_keyframe(hs[0])
        op, ov = oracle.backproject(kf[0][1], kf[0][2], kf[0][3], reverse=reverse)
        want_desc = kf[0][0][::-1] if reverse else kf[0][0]
        assert np.array_equal(back["desc"], want_desc) and back["pos"].tobytes() == op.tobytes() and np.array_equal(back["valid"], ov)
        r = est.estimateEdges([hs[0]], [hs[1]])[0]
        cams = [dict(desc=np.ascontiguousarray(want_desc), pos=op, valid=ov, feature_type=2, sensor_frame=0)] * 2
        o = oracle.estimate_edge([cams[0]], [cams[1]])
        assert bool(r["ok"]) == o["ok"] and r["consensus"] == o["consensus"] and np.array_equal(r["T"].reshape(4, 4), o["T"])
        assert r["ok"] and np.allclose(r["T"].reshape(4, 4), np.eye(4), atol=1e-6)
    est.clear()


@pytest.mark.gpu
def test_gpu_wire_decode_and_ingest(est, oracle):
    f, t, _ = S.make_pair(900, seed=6)
    rng = np.random.default_rng(7)
    blobs = []
    for kfr in (f, t):
        n = len(kfr["desc"])
        fl = kfr["desc"].astype(np.float32)
        blobs.append(encode_features(fl, rng.integers(0, 640, n), rng.integers(0, 480, n), kfr["valid"], np.full(n, -1.0), kfr["pos"]))
    gd, gp, gv, guv = est.wire_decode(blobs[0])
    od, op, ov, ouv = oracle.wire_decode(blobs[0])
    assert np.array_equal(gd, od) and gp.tobytes() == op.tobytes() and np.array_equal(gv, ov) and np.array_equal(guv, ouv)
    weird = f["desc"].astype(np.float32)
    weird[:, 0] = rng.choice([255.9, 256.0, -1.0, 1e10, -1e10, np.nan, np.inf, 3.7, 511.0, -300.5], len(weird))
    wb = encode_features(weird, np.zeros(len(weird)), np.zeros(len(weird)), f["valid"], np.zeros(len(weird)), f["pos"])
    assert np.array_equal(est.wire_decode(wb)[0], oracle.wire_decode(wb)[0])
    # straight into the store, then through the path
    est.clear()
    hf, ht = est.add_keyframe_wire(blobs[0]), est.add_keyframe_wire(blobs[1])
    r = est.estimateEdges([hf], [ht])[0]
    o = oracle.estimate_edge([f], [t])
    assert r["consensus"] == o["consensus"] and np.array_equal(r["T"].reshape(4, 4), o["T"])
    # malformed input is refused, not decoded
    with pytest.raises(Exception):
        est.add_keyframe_wire(blobs[0][:-10])
    b48 = encode_features(np.zeros((3, 48), np.float32), [0] * 3, [0] * 3, [1] * 3, [0] * 3, np.zeros((3, 3)))
    with pytest.raises(Exception):                      # neither 32 nor 64 columns
        est.add_keyframe_wire(b48)
    mixed = encode_features(np.zeros((2, 64), np.float32), [0] * 2, [0] * 2, [1] * 2, [0] * 2, np.zeros((2, 3)))
    mixed = struct.pack("<I", 3) + mixed[4:] + b48[4:4 + 17 + 48 * 4 + 24] + b"\0" * 64      # third element has another length
    with pytest.raises(Exception):
        est.add_keyframe_wire(mixed)
    assert est.read_keyframe(est.add_keyframe_wire(struct.pack("<I", 0)))["desc"].shape == (0, 32)
    est.clear()


@pytest.mark.parametrize("nb", [32, 64])
def test_oracle_wire_encode_is_the_inverse_of_decode(oracle, nb):
    """FeatureData::toMsg (sensor_data.cpp:77-122): u/v from feature_positions_2d_, keypoint_strength -1, one float32 per
    descriptor byte - against the field-by-field struct packing above, and decode(encode(x)) == x"""
    f, _, _ = S.make_pair(70, seed=3, desc_bytes=nb)
    rng = np.random.default_rng(4)
    uv = np.stack([rng.integers(0, 640, 70), rng.integers(0, 480, 70)], 1).astype(np.int32)
    blob = oracle.wire_encode(f, uv)
    assert blob == encode_features(f["desc"].astype(np.float32), uv[:, 0], uv[:, 1], f["valid"], np.full(70, -1.0), f["pos"])
    desc, pos, valid, uv2 = oracle.wire_decode(blob, cols=nb)
    assert np.array_equal(desc, f["desc"]) and np.array_equal(pos, f["pos"]) and np.array_equal(valid, f["valid"])
    assert np.array_equal(uv2, uv)
    assert oracle.wire_encode(f)[:4 + 8] == struct.pack("<Iii", 70, 0, 0)                   # no 2-D positions: zeros


@pytest.mark.gpu
@pytest.mark.parametrize("nb", [32, 64])
def test_gpu_wire_encode(est, oracle, nb):
    f, t, _ = S.make_pair(500, seed=8, desc_bytes=nb)
    rng = np.random.default_rng(9)
    uv = np.stack([rng.integers(0, 640, 500), rng.integers(0, 480, 500)], 1).astype(np.int32)
    est.clear()
    h = est.add_keyframe([f])
    assert est.wire_encode(h, uv=uv) == oracle.wire_encode(f, uv)
    assert est.wire_encode(h) == oracle.wire_encode(f)
    # store -> message -> store: the reloaded keyframe is the same keyframe (what a RosbagStorage resume does)
    h2 = est.add_keyframe_wire(est.wire_encode(h, uv=uv), feature_type=f["feature_type"])
    back = est.read_keyframe(h2)
    assert np.array_equal(back["desc"], f["desc"]) and np.array_equal(back["pos"], f["pos"]) and np.array_equal(back["valid"], f["valid"])
    ht = est.add_keyframe([t])
    r = est.estimateEdges([h, h2], [ht, ht])
    assert r[0].tobytes() == r[1].tobytes() and r[0]["consensus"] == oracle.estimate_edge([f], [t])["consensus"]
    empty = est.add_keyframe([dict(desc=np.zeros((0, nb), np.uint8), pos=np.zeros((0, 3)), valid=np.zeros(0, np.uint8))])
    assert est.wire_encode(empty) == struct.pack("<I", 0)
    est.clear()
