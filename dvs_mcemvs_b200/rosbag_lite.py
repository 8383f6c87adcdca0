"""Minimal ROS bag (format 2.0, uncompressed chunks) reader / writer for the two message types the
mapping path consumes — no ROS installation needed.

  geometry_msgs/PoseStamped, geometry_msgs/PoseWithCovarianceStamped, nav_msgs/Odometry, vicon/Subject
                             -> STAMPED_POSE_DTYPE control poses (what parse_rosbag_gt collects,
                                mapper_emvs_stereo/src/data_loading.cpp:305-465; the bundled
                                data/DSEC/*/pose.bag files are PoseStamped)
  dvs_msgs/EventArray        -> EVENT_DTYPE events (data_loading.cpp:33-219); a serialised
                                dvs_msgs/Event is 13 bytes {uint16 x, uint16 y, time ts, bool polarity},
                                the in-memory struct the engine takes is the 16-byte padded one
  sensor_msgs/CameraInfo     -> dict(width, height, distortion_model, D, K, R, P)

parse_rosbag() is data_loading::parse_rosbag itself: the time window relative to the process-wide first stamp, the
re-timing of events and poses, the stop-after-the-message rule and the final sort.

The writer exists so that tests can round-trip synthetic streams; it emits one uncompressed chunk per
call plus the index records rosbag tools expect.  bz2 / lz4 chunk compression is not supported.
"""
import struct

import numpy as np

from ._capi import EVENT_DTYPE, STAMPED_POSE_DTYPE

MAGIC = b"#ROSBAG V2.0\n"
OP_MSG, OP_BAG_HEADER, OP_INDEX, OP_CHUNK, OP_CHUNK_INFO, OP_CONNECTION = 2, 3, 4, 5, 6, 7


# ---- record level ----------------------------------------------------------------------------------
def _parse_header(buf):
    fields, o = {}, 0
    while o < len(buf):
        (n,) = struct.unpack_from("<I", buf, o)
        o += 4
        k, v = bytes(buf[o:o + n]).split(b"=", 1)
        fields[k.decode()] = v
        o += n
    return fields


def _records(buf, start, end):
    o = start
    while o < end:
        (hl,) = struct.unpack_from("<I", buf, o)
        hdr = _parse_header(buf[o + 4:o + 4 + hl])
        o += 4 + hl
        (dl,) = struct.unpack_from("<I", buf, o)
        yield hdr, o + 4, dl
        o += 4 + dl


def read_messages(path, topics=None, types=None):
    """Yields (topic, msg_type, receive_time(sec, nsec), payload memoryview) in file order."""
    buf = memoryview(open(path, "rb").read())
    if bytes(buf[:len(MAGIC)]) != MAGIC:
        raise ValueError(f"{path}: not a ROS bag v2.0")
    conns = {}

    def handle(hdr, off, n, base):
        op = hdr["op"][0]
        if op == OP_CONNECTION:
            info = _parse_header(base[off:off + n])
            conns[struct.unpack("<I", hdr["conn"])[0]] = (hdr["topic"].decode(), info["type"].decode())
        elif op == OP_MSG:
            topic, mtype = conns[struct.unpack("<I", hdr["conn"])[0]]
            if (topics is None or topic in topics) and (types is None or mtype in types):
                return topic, mtype, struct.unpack("<II", hdr["time"]), base[off:off + n]
        return None

    for hdr, off, n in _records(buf, len(MAGIC), len(buf)):
        op = hdr["op"][0]
        if op == OP_CHUNK:
            if hdr["compression"] != b"none":
                raise NotImplementedError(f"{path}: {hdr['compression'].decode()} chunks are not supported")
            for h2, o2, n2 in _records(buf, off, off + n):
                m = handle(h2, o2, n2, buf)
                if m:
                    yield m
        else:
            m = handle(hdr, off, n, buf)
            if m:
                yield m


# ---- message level ---------------------------------------------------------------------------------
def _skip_std_header(p, o=0):
    """std_msgs/Header {uint32 seq; time stamp; string frame_id} -> (stamp(sec, nsec), offset after it)."""
    _, sec, nsec, flen = struct.unpack_from("<IIII", p, o)
    return (sec, nsec), o + 16 + flen


def read_poses(path, topic=None):
    """All geometry_msgs/PoseStamped messages [of `topic`] as STAMPED_POSE_DTYPE, sorted by header stamp
    (the std::map<ros::Time, Transformation> ordering of the reference; later duplicates of a stamp win)."""
    out = {}
    for _, _, _, p in read_messages(path, topics=None if topic is None else {topic}, types={"geometry_msgs/PoseStamped"}):
        stamp, o = _skip_std_header(p)
        px, py, pz, qx, qy, qz, qw = struct.unpack_from("<7d", p, o)
        out[stamp] = ((qw, qx, qy, qz), (px, py, pz))
    arr = np.zeros(len(out), STAMPED_POSE_DTYPE)
    for i, stamp in enumerate(sorted(out)):
        arr["sec"][i], arr["nsec"][i] = stamp
        arr["T"]["q"][i], arr["T"]["t"][i] = out[stamp]
    return arr


_EV_WIRE = np.dtype([("x", "<u2"), ("y", "<u2"), ("sec", "<u4"), ("nsec", "<u4"), ("polarity", "u1")])   # 13 bytes, packed
assert _EV_WIRE.itemsize == 13


def read_events(path, topic=None, t_min=None, t_max=None, sort=True):
    """All dvs_msgs/EventArray messages [of `topic`] concatenated as EVENT_DTYPE, optionally restricted to
    t_min <= t < t_max (seconds) and sorted by timestamp like parse_rosbag does (data_loading.cpp:196-214)."""
    parts = []
    for _, _, _, p in read_messages(path, topics=None if topic is None else {topic}, types={"dvs_msgs/EventArray"}):
        _, o = _skip_std_header(p)
        _, _, n = struct.unpack_from("<III", p, o)   # height, width, events.size()
        parts.append(np.frombuffer(p, _EV_WIRE, count=n, offset=o + 12))
    wire = np.concatenate(parts) if parts else np.zeros(0, _EV_WIRE)
    ev = np.zeros(wire.shape[0], EVENT_DTYPE)
    for f in ("x", "y", "sec", "nsec", "polarity"):
        ev[f] = wire[f]
    t = ev["sec"].astype(np.float64) + 1e-9 * ev["nsec"]
    keep = np.ones(ev.shape[0], bool)
    if t_min is not None:
        keep &= t >= t_min
    if t_max is not None:
        keep &= t < t_max
    ev, t = ev[keep], t[keep]
    if sort:
        key = ev["sec"].astype(np.uint64) * np.uint64(1_000_000_000) + ev["nsec"].astype(np.uint64)
        ev = ev[np.argsort(key, kind="stable")]
    return ev


# ---- data_loading::parse_rosbag ------------------------------------------------------------------------
class TimeOrigin:
    """The `static ros::Time initial_timestamp` of data_loading.cpp:30-31: the first stamp seen by the first
    parse_rosbag call of a process becomes the origin of every later call (left and right bags share it)."""

    def __init__(self):
        self.stamp = None     # (sec, nsec)

    def to_sec(self):
        return self.stamp[0] + 1e-9 * self.stamp[1]


DEFAULT_ORIGIN = TimeOrigin()


def _time_from_sec(t):
    """ros::Time(double): sec = floor(t), nsec = round((t - sec) * 1e9), normalised; negative times throw."""
    if t < 0 or t >= 4294967296.0:
        raise ValueError(f"ros::Time out of range: {t}")
    sec = int(np.floor(t))
    nsec = int(np.floor((t - sec) * 1e9 + 0.5))
    if nsec >= 1_000_000_000:
        sec, nsec = sec + 1, nsec - 1_000_000_000
    return sec, nsec


def parse_camera_info(p):
    """sensor_msgs/CameraInfo payload -> dict(width, height, distortion_model, D, K, R, P)."""
    _, o = _skip_std_header(p)
    height, width, n = struct.unpack_from("<III", p, o)
    o += 12
    model = bytes(p[o:o + n]).decode()
    o += n
    (nd,) = struct.unpack_from("<I", p, o)
    o += 4
    D = np.frombuffer(p, "<f8", count=nd, offset=o).copy()
    o += 8 * nd
    K = np.frombuffer(p, "<f8", count=9, offset=o).reshape(3, 3).copy()
    R = np.frombuffer(p, "<f8", count=9, offset=o + 72).reshape(3, 3).copy()
    P = np.frombuffer(p, "<f8", count=12, offset=o + 144).reshape(3, 4).copy()
    return dict(width=int(width), height=int(height), distortion_model=model, D=D, K=K, R=R, P=P)


# Pose-carrying message types of data_loading.cpp.  parse_rosbag (events + poses, :33-219) accepts the first three,
# parse_rosbag_gt (:305-465) all four.  Wire layouts (ROS1 serialisation, little endian):
#   geometry_msgs/PoseStamped                 Header, Point position (3 f64), Quaternion orientation (x y z w f64)
#   geometry_msgs/PoseWithCovarianceStamped   the same followed by float64[36] covariance
#   nav_msgs/Odometry                         Header, string child_frame_id, PoseWithCovariance pose, TwistWithCovariance twist
#   vicon/Subject (EVIMO2)                    Header, string name, position (3 f64), orientation (x y z w f64), then
#                                             occlusion flag / markers.  The vicon package is not vendored in the reference
#                                             tree; this is the field order of the EVIMO recording tools' Subject.msg, and only
#                                             the fields data_loading.cpp:115-141 reads (header.stamp, position, orientation)
#                                             are decoded.
_POSE_TYPES = {"geometry_msgs/PoseStamped", "geometry_msgs/PoseWithCovarianceStamped", "vicon/Subject"}
_POSE_TYPES_GT = _POSE_TYPES | {"nav_msgs/Odometry"}


def _skip_string(p, o):
    (n,) = struct.unpack_from("<I", p, o)
    return o + 4 + n


def parse_pose_message(mtype, p):
    """-> (stamp (sec, nsec), quaternion (w, x, y, z), position (x, y, z)) of one pose-carrying message."""
    stamp, o = _skip_std_header(p)
    if mtype in ("nav_msgs/Odometry", "vicon/Subject"):
        o = _skip_string(p, o)            # child_frame_id / subject name
    elif mtype not in _POSE_TYPES_GT:
        raise ValueError(f"not a pose message type: {mtype}")
    px, py, pz, qx, qy, qz, qw = struct.unpack_from("<7d", p, o)
    return stamp, (qw, qx, qy, qz), (px, py, pz)


def parse_rosbag_gt(path, pose_topic, tmin=0.0, tmax=float("inf"), origin=None):
    """data_loading::parse_rosbag_gt (data_loading.cpp:305-465): the control poses of `pose_topic` — vicon/Subject,
    geometry_msgs/PoseStamped, geometry_msgs/PoseWithCovarianceStamped or nav_msgs/Odometry messages — restricted to
    [tmin, tmax] seconds after the process-wide first stamp and re-timed relative to it.  Same literal rules as
    parse_rosbag: a pose past tmax is still inserted and stops the loop afterwards; an equal re-timed stamp does not
    replace an earlier pose (std::map::insert).  Returns STAMPED_POSE_DTYPE sorted by stamp."""
    origin = DEFAULT_ORIGIN if origin is None else origin
    msgs = sorted(read_messages(path, topics={pose_topic}), key=lambda m: m[2])
    poses, go_on = {}, True
    for topic, mtype, _, p in msgs:
        if not go_on:
            break
        if mtype not in _POSE_TYPES_GT:
            continue
        stamp, q, t = parse_pose_message(mtype, p)
        if origin.stamp is None:
            origin.stamp = stamp
        rel = (stamp[0] - origin.stamp[0]) + 1e-9 * (stamp[1] - origin.stamp[1])
        if rel < tmin:
            continue
        if rel > tmax:
            go_on = False
        poses.setdefault(_time_from_sec((stamp[0] + 1e-9 * stamp[1]) - origin.to_sec()), (q, t))
    arr = np.zeros(len(poses), STAMPED_POSE_DTYPE)
    for i, k in enumerate(sorted(poses)):
        arr["sec"][i], arr["nsec"][i] = k
        arr["T"]["q"][i], arr["T"]["t"][i] = poses[k]
    return arr


def parse_rosbag(path, event_topic, camera_info_topic=None, pose_topic=None, tmin=0.0, tmax=float("inf"),
                 events_offset=0.0, origin=None):
    """data_loading::parse_rosbag (data_loading.cpp:33-219 with a pose topic, :221-303 without): events, control
    poses and the last CameraInfo of one bag, restricted to [tmin, tmax] seconds after the process-wide first stamp
    and re-timed relative to it (events additionally shifted by -events_offset).

    Kept literally: messages are visited in bag-time order; an event / pose past tmax stops the loop only AFTER its
    message has been consumed (so the last message may contribute stamps > tmax); empty EventArrays are skipped; the
    events are sorted by timestamp at the end; poses come back sorted by their re-timed stamp (std::map), a later
    pose with an equal stamp does not replace an earlier one (std::map::insert).
    Returns (events EVENT_DTYPE, poses STAMPED_POSE_DTYPE, camera_info dict or None)."""
    origin = DEFAULT_ORIGIN if origin is None else origin
    topics = {t for t in (event_topic, camera_info_topic, pose_topic) if t}
    msgs = sorted(read_messages(path, topics=topics), key=lambda m: m[2])      # rosbag::View: by receive time (stable)
    ev_parts, poses, cam_info, go_on = [], {}, None, True
    for topic, mtype, _, p in msgs:
        if not go_on:
            break
        if topic == event_topic and mtype == "dvs_msgs/EventArray":
            _, o = _skip_std_header(p)
            _, _, n = struct.unpack_from("<III", p, o)
            if n == 0:
                continue
            wire = np.frombuffer(p, _EV_WIRE, count=n, offset=o + 12)
            if origin.stamp is None:
                origin.stamp = (int(wire["sec"][0]), int(wire["nsec"][0]))
            # (a - b).toSec() of two ros::Time values: the exact sec / nsec difference as a double
            rel = (wire["sec"].astype(np.int64) - origin.stamp[0]).astype(np.float64) \
                + 1e-9 * (wire["nsec"].astype(np.int64) - origin.stamp[1]).astype(np.float64)
            if (rel > tmax).any():
                go_on = False
            keep = ~(rel < tmin)
            if keep.any():
                w = wire[keep]
                t_new = (w["sec"].astype(np.float64) + 1e-9 * w["nsec"].astype(np.float64)) - origin.to_sec() - events_offset
                if (t_new < 0).any():
                    raise ValueError("parse_rosbag: a re-timed event stamp is negative (ros::Time would throw)")
                sec = np.floor(t_new)
                nsec = np.floor((t_new - sec) * 1e9 + 0.5)
                carry = nsec >= 1e9
                sec, nsec = sec + carry, nsec - carry * 1e9
                e = np.zeros(w.shape[0], EVENT_DTYPE)
                e["x"], e["y"], e["polarity"] = w["x"], w["y"], w["polarity"]
                e["sec"], e["nsec"] = sec.astype(np.uint32), nsec.astype(np.uint32)
                ev_parts.append(e)
        elif topic == camera_info_topic and mtype == "sensor_msgs/CameraInfo":
            cam_info = parse_camera_info(p)
        elif topic == pose_topic and mtype in _POSE_TYPES:     # vicon/Subject, PoseStamped, PoseWithCovarianceStamped (:111-219)
            stamp, q, t = parse_pose_message(mtype, p)
            if origin.stamp is None:
                origin.stamp = stamp
            rel = (stamp[0] - origin.stamp[0]) + 1e-9 * (stamp[1] - origin.stamp[1])
            if rel < tmin:
                continue
            if rel > tmax:
                go_on = False
            key = _time_from_sec((stamp[0] + 1e-9 * stamp[1]) - origin.to_sec())
            poses.setdefault(key, (q, t))
    ev = np.concatenate(ev_parts) if ev_parts else np.zeros(0, EVENT_DTYPE)
    key = ev["sec"].astype(np.uint64) * np.uint64(1_000_000_000) + ev["nsec"].astype(np.uint64)
    ev = ev[np.argsort(key, kind="stable")]
    arr = np.zeros(len(poses), STAMPED_POSE_DTYPE)
    for i, k in enumerate(sorted(poses)):
        arr["sec"][i], arr["nsec"][i] = k
        arr["T"]["q"][i], arr["T"]["t"][i] = poses[k]
    return ev, arr, cam_info


# ---- writer (tests / synthetic data) ------------------------------------------------------------------
def _hdr(**fields):
    b = b"".join(struct.pack("<I", len(k) + 1 + len(v)) + k.encode() + b"=" + v for k, v in fields.items())
    return struct.pack("<I", len(b)) + b


def _record(header, data):
    return header + struct.pack("<I", len(data)) + data


def _std_header(seq, sec, nsec, frame_id=b""):
    return struct.pack("<IIII", seq, sec, nsec, len(frame_id)) + frame_id


def serialize_camera_info(info, sec=0, nsec=0):
    model = info["distortion_model"].encode()
    D = np.asarray(info["D"], "<f8").reshape(-1)
    return (_std_header(0, sec, nsec) + struct.pack("<III", info["height"], info["width"], len(model)) + model +
            struct.pack("<I", len(D)) + D.tobytes() + np.asarray(info["K"], "<f8").tobytes() +
            np.asarray(info["R"], "<f8").tobytes() + np.asarray(info["P"], "<f8").tobytes() +
            struct.pack("<II", 0, 0) + struct.pack("<IIIIB", 0, 0, 0, 0, 0))


_POSE_MD5 = {"geometry_msgs/PoseStamped": "d3812c3cbc69362b77dc0b19b345f8f5",
             "geometry_msgs/PoseWithCovarianceStamped": "953b798c0f514ff060a53a3498ce6246",
             "nav_msgs/Odometry": "cd5e73d190d741a2f92e81eda573aca7",
             "vicon/Subject": "*"}


def write_bag(path, poses=None, pose_topic="/pose", events=None, event_topic="/dvs/events", sensor=(640, 480),
              events_per_message=5000, camera_info=None, camera_info_topic="/dvs/camera_info", pose_with_covariance=False,
              pose_type=None):
    """Writes one bag holding the given control poses (STAMPED_POSE_DTYPE), events (EVENT_DTYPE) and / or one
    sensor_msgs/CameraInfo (a dict as parse_camera_info returns).  pose_type: one of the four pose-carrying message
    types (default geometry_msgs/PoseStamped; pose_with_covariance=True selects PoseWithCovarianceStamped)."""
    conns, msgs = [], []   # msgs: (conn id, (sec, nsec), payload)
    if poses is not None:
        cid = len(conns)
        if pose_type is None:
            pose_type = "geometry_msgs/PoseWithCovarianceStamped" if pose_with_covariance else "geometry_msgs/PoseStamped"
        conns.append((pose_topic, pose_type, _POSE_MD5[pose_type]))
        for i, p in enumerate(np.asarray(poses, STAMPED_POSE_DTYPE)):
            q, t = p["T"]["q"], p["T"]["t"]
            pose = struct.pack("<7d", t[0], t[1], t[2], q[1], q[2], q[3], q[0])
            hdr = _std_header(i, int(p["sec"]), int(p["nsec"]))
            if pose_type == "geometry_msgs/PoseStamped":
                payload = hdr + pose
            elif pose_type == "geometry_msgs/PoseWithCovarianceStamped":
                payload = hdr + pose + struct.pack("<36d", *([0.0] * 36))
            elif pose_type == "nav_msgs/Odometry":      # child_frame_id, PoseWithCovariance, TwistWithCovariance
                child = b"base_link"
                payload = (hdr + struct.pack("<I", len(child)) + child + pose + struct.pack("<36d", *([0.0] * 36)) +
                           struct.pack("<6d", *([0.0] * 6)) + struct.pack("<36d", *([0.0] * 36)))
            else:                                       # vicon/Subject: name, position, orientation, occluded, no markers
                name = b"subject"
                payload = hdr + struct.pack("<I", len(name)) + name + pose + struct.pack("<BI", 0, 0)
            msgs.append((cid, (int(p["sec"]), int(p["nsec"])), payload))
    if camera_info is not None:
        cid = len(conns)
        conns.append((camera_info_topic, "sensor_msgs/CameraInfo", "c9a58c1b0b154e0e6da7578cb991d214"))
        first = min([m[1] for m in msgs], default=(0, 0))
        msgs.append((cid, first, serialize_camera_info(camera_info, *first)))
    if events is not None:
        cid = len(conns)
        conns.append((event_topic, "dvs_msgs/EventArray", "5e8beee5a6c107e504c2e78903c224b8"))
        ev = np.asarray(events, EVENT_DTYPE)
        for i, s in enumerate(range(0, len(ev), events_per_message)):
            e = ev[s:s + events_per_message]
            wire = np.zeros(len(e), _EV_WIRE)
            for f in ("x", "y", "sec", "nsec", "polarity"):
                wire[f] = e[f]
            stamp = (int(e["sec"][-1]), int(e["nsec"][-1]))
            msgs.append((cid, stamp, _std_header(i, *stamp) + struct.pack("<III", sensor[1], sensor[0], len(e)) + wire.tobytes()))
    msgs.sort(key=lambda m: m[1])
    conn_recs = [_record(_hdr(op=bytes([OP_CONNECTION]), conn=struct.pack("<I", c), topic=t.encode()),
                         _hdr(topic=t.encode(), type=ty.encode(), md5sum=md5.encode(), message_definition=b"")[4:])
                 for c, (t, ty, md5) in enumerate(conns)]
    chunk, index = b"".join(conn_recs), {c: [] for c in range(len(conns))}
    for c, stamp, payload in msgs:
        index[c].append((stamp, len(chunk)))
        chunk += _record(_hdr(op=bytes([OP_MSG]), conn=struct.pack("<I", c), time=struct.pack("<II", *stamp)), payload)
    chunk_pos = len(MAGIC) + 4096
    body = _record(_hdr(op=bytes([OP_CHUNK]), compression=b"none", size=struct.pack("<I", len(chunk))), chunk)
    for c, entries in index.items():
        body += _record(_hdr(op=bytes([OP_INDEX]), ver=struct.pack("<I", 1), conn=struct.pack("<I", c),
                             count=struct.pack("<I", len(entries))),
                        b"".join(struct.pack("<III", s[0], s[1], off) for s, off in entries))
    index_pos = chunk_pos + len(body)
    t0, t1 = (msgs[0][1], msgs[-1][1]) if msgs else ((0, 0), (0, 0))
    tail = b"".join(conn_recs) + _record(
        _hdr(op=bytes([OP_CHUNK_INFO]), ver=struct.pack("<I", 1), chunk_pos=struct.pack("<Q", chunk_pos),
             start_time=struct.pack("<II", *t0), end_time=struct.pack("<II", *t1), count=struct.pack("<I", len(conns))),
        b"".join(struct.pack("<II", c, len(e)) for c, e in index.items()))
    bag_hdr = _hdr(op=bytes([OP_BAG_HEADER]), index_pos=struct.pack("<Q", index_pos), conn_count=struct.pack("<I", len(conns)),
                   chunk_count=struct.pack("<I", 1))
    pad = 4096 - len(bag_hdr) - 4
    with open(path, "wb") as f:
        f.write(MAGIC + bag_hdr + struct.pack("<I", pad) + b" " * pad + body + tail)
