"""Host-side sharding plan for the multi-GPU DSI build.

Voting is a sum over events, so DSIs of disjoint event sets add — the reference relies on the
same additivity for its temporal arithmetic-mean fusion (process2.cpp:229-236).  The unit of
sharding is the PACKET (1024 events sharing one pose, mapper_emvs_stereo.cpp:88-126): packet
boundaries are decided once on the unsharded stream, so sharded == unsharded up to the order of
the float additions, and the integer vote counts add up exactly.

Plan: rank r of `world` builds, for EVERY camera, the contiguous packet range r of that
camera's packet list (balanced to within one packet).  One sum-exchange per camera DSI follows
(ncclAllReduce, or the peer-memory reduce fused into the fuse+argmax sweep), then fusion.
"""


def split_range(n_items, n_parts):
    """Contiguous, balanced partition of range(n_items) into n_parts (lo, hi) pairs; the first
    n_items % n_parts parts get one extra item.  Parts may be empty when n_items < n_parts."""
    if n_parts < 1:
        raise ValueError("n_parts must be >= 1")
    base, extra = divmod(int(n_items), int(n_parts))
    out, lo = [], 0
    for p in range(n_parts):
        hi = lo + base + (1 if p < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def plan(n_packets_per_camera, world, rank):
    """-> [(camera, packet_lo, packet_hi)] for `rank`: its sub-interval of every camera."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return [(cam, *split_range(n, world)[rank]) for cam, n in enumerate(n_packets_per_camera)]


def camera_groups(n_cams, world):
    """Camera x sub-interval sharding: when the GPUs divide evenly over the cameras, camera c is
    built by the group of g = world / n_cams consecutive ranks [c*g, (c+1)*g).  -> g, or 0 when that is not possible
    (fewer GPUs than cameras, or not a multiple): every rank then builds a sub-interval of every camera (plan())."""
    if n_cams >= 1 and world >= n_cams and world % n_cams == 0:
        return world // n_cams
    return 0


def plan2d(n_packets_per_camera, world, rank):
    """-> [(camera, packet_lo, packet_hi)] for `rank` under camera x sub-interval sharding: ONE camera per rank, its
    packet list split over the g ranks of the camera's group.  Compared with plan() a rank votes twice as many events
    of half as many cameras: half the per-slab merge / re-zero / exchange work for the same votes, and nothing to sum
    across groups.  Falls back to plan() when the GPUs do not divide evenly over the cameras."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    g = camera_groups(len(n_packets_per_camera), world)
    if g == 0:
        return plan(n_packets_per_camera, world, rank)
    cam, part = divmod(rank, g)
    return [(cam, *split_range(n_packets_per_camera[cam], g)[part])]


def participants(n_cams, world, two_d=True):
    """builds[c][r] = 1 iff rank r builds camera c (the table emvs_exchange_set_participants takes)."""
    g = camera_groups(n_cams, world) if two_d else 0
    return [[1 if (g == 0 or r // g == c) else 0 for r in range(world)] for c in range(n_cams)]


def row_bands(dimY, world):
    """Row band of the depth/confidence maps owned by each rank in the fused reduce+fuse+argmax
    sweep: [(row_lo, row_hi)] * world."""
    return split_range(dimY, world)
