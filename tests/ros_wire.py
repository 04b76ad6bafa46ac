"""ROS1 serialisation (little endian, unpadded) of the reference's messages, written from the .msg definitions for the tests:
graph_slam_msgs/msg/{Node,SensorDataArray,SensorData,Features,Feature,DepthImage}.msg plus std_msgs/Header,
geometry_msgs/{Pose,Point}, sensor_msgs/{CameraInfo,Image,LaserScan,RegionOfInterest}.  Independent of the C walk."""
import struct

import numpy as np


def _str(s):
    b = s.encode() if isinstance(s, str) else bytes(s)
    return struct.pack("<I", len(b)) + b


def header(seq=0, secs=0, nsecs=0, frame_id=""):
    return struct.pack("<III", seq, secs, nsecs) + _str(frame_id)


def pose(p=(0, 0, 0), q=(0, 0, 0, 1)):
    return struct.pack("<7d", *p, *q)


def image(h=0, w=0, encoding="", data=b""):
    return header() + struct.pack("<II", h, w) + _str(encoding) + struct.pack("<BI", 0, w) + struct.pack("<I", len(data)) + bytes(data)


def camera_info(d=(0.1, 0.2, 0.0, 0.0, 0.0)):
    out = header(frame_id="/camera_rgb_optical_frame") + struct.pack("<II", 480, 640) + _str("plumb_bob")
    out += struct.pack("<I", len(d)) + struct.pack("<%dd" % len(d), *d)
    out += struct.pack("<9d", *([525, 0, 319.5, 0, 525, 239.5, 0, 0, 1])) + struct.pack("<9d", *np.eye(3).ravel())
    out += struct.pack("<12d", *np.eye(3, 4).ravel()) + struct.pack("<II", 1, 1) + struct.pack("<IIIIB", 0, 0, 0, 0, 0)
    return out


def laserscan(ranges=(), intensities=()):
    out = header(frame_id="/laser") + struct.pack("<7f", -1.5, 1.5, 0.01, 0.0, 0.1, 0.05, 30.0)
    out += struct.pack("<I", len(ranges)) + struct.pack("<%df" % len(ranges), *ranges)
    out += struct.pack("<I", len(intensities)) + struct.pack("<%df" % len(intensities), *intensities)
    return out


def sensor_data(sensor_type=1, sensor_frame="/camera_rgb_optical_frame", features_blob=struct.pack("<I", 0), descriptor_type=2,
                displacement=((0.1, 0.2, 0.3), (0, 0, 0, 1)), gist=(), ranges=(), depth=b"", stamp=(12, 34)):
    """one graph_slam_msgs/SensorData; features_blob is the serialised Feature[] field (uint32 count + elements)"""
    out = header(seq=7, secs=stamp[0], nsecs=stamp[1], frame_id=sensor_frame) + struct.pack("<i", sensor_type)
    out += pose(*displacement) + _str(sensor_frame)
    out += header(secs=stamp[0], nsecs=stamp[1], frame_id=sensor_frame) + struct.pack("<i", descriptor_type) + bytes(features_blob) + camera_info()
    out += image(2, 2, "32FC1", depth) + image()
    out += struct.pack("<I", len(gist)) + struct.pack("<%df" % len(gist), *gist)
    out += laserscan(ranges, ranges) + struct.pack("<3d", 1.0, 2.0, 3.0)
    return out


def node(node_id, sensors, stamps=((1, 2), (3, 4)), edge_ids=("e1", "edge-22")):
    out = struct.pack("<I", len(stamps)) + b"".join(struct.pack("<II", *s) for s in stamps)
    out += _str(node_id) + pose((1, 2, 3)) + pose((4, 5, 6))
    out += header(frame_id="/map") + struct.pack("<I", len(sensors)) + b"".join(sensors)
    out += struct.pack("<I", len(edge_ids)) + b"".join(_str(e) for e in edge_ids)
    out += struct.pack("<Bd", 1, 0.25)
    return out
