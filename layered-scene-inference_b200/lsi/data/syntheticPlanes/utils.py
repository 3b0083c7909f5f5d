"""Mirror of lsi/data/syntheticPlanes/utils.py (reference tree): world-layout helpers of the synthetic planar-room generator.

Same function names and arguments.  The texture loader of the reference (QueuedRandomTextureLoader: PASCAL object crops and
SUN backgrounds through TF file queues, utils.py:206-295) has no counterpart: those datasets are not available, textures are
procedural and generated on the GPU (lsi.data.syntheticPlanes.data.WorldGenerator).
"""
import math

import numpy as np


def resize_instrinsic(intrinsic, scale_x, scale_y):
    """utils.py:29-33 (name as in the reference): diag(sx, sy, 1) @ K."""
    return np.diag([float(scale_x), float(scale_y), 1.0]) @ np.asarray(intrinsic, dtype=np.float64)


def dims2kmat(w_plane, h_plane, w_tex, h_tex):
    """utils.py:36-51 -- intrinsics that map a w_plane x h_plane fronto-parallel plane at z = 1 onto a w_tex x h_tex texture."""
    return np.array([[w_tex / float(w_plane), 0.0, 0.5 * w_tex], [0.0, h_tex / float(h_plane), 0.5 * h_tex], [0.0, 0.0, 1.0]])


def _axis(v):
    v = np.asarray(v, dtype=np.float64).reshape(3, 1)
    return v / np.sqrt(float((v * v).sum()))


def get_centre(pt, x_dir, y_dir, w, h, off_x=0.5, off_y=0.5):
    """utils.py:54-75 -- centre of a w x h plane given a point that sits (off_x w, off_y h) from its top-left corner."""
    shift = (0.5 - off_x) * w * _axis(x_dir) + (0.5 - off_y) * h * _axis(y_dir)
    return np.asarray(pt, dtype=np.float64).reshape(3, 1) + shift


def canonical_transform(centre_s, x_dir, y_dir, trans_init=None):
    """utils.py:78-106 -- (rot, trans) taking the canonical plane (axes e_x, e_y, centre trans_init = (0,0,1)) to the plane with
    axes x_dir, y_dir and centre centre_s."""
    ex, ey = _axis(x_dir), _axis(y_dir)
    rot = np.hstack([ex, ey, np.cross(ex.ravel(), ey.ravel()).reshape(3, 1)])
    origin = np.array([0.0, 0.0, 1.0]) if trans_init is None else np.asarray(trans_init, dtype=np.float64)
    return rot, np.asarray(centre_s, dtype=np.float64).reshape(3, 1) - rot @ origin.reshape(3, 1)


# (anchor corner, x axis, y axis, width extent, height extent) of the five box faces in the reference's order:
# front wall, floor, ceiling, left wall, right wall; corners / extents index into (x0, y0, z0, x1, y1, z1)
_FACES = (((0, 1, 5), (1, 0, 0), (0, 1, 0), (3, 0), (4, 1)),
          ((0, 4, 5), (1, 0, 0), (0, 0, -1), (3, 0), (5, 2)),
          ((0, 1, 5), (1, 0, 0), (0, 0, -1), (3, 0), (5, 2)),
          ((0, 1, 2), (0, 0, 1), (0, 1, 0), (5, 2), (4, 1)),
          ((3, 1, 2), (0, 0, 1), (0, 1, 0), (5, 2), (4, 1)))


def box_planes(extent):
    """utils.py:109-176 -- plane parameters (dicts: pt, x_dir, y_dir, w, h, off_x, off_y) of a box (x0, y0, z0, x1, y1, z1)."""
    e = [float(v) for v in extent]
    return [{'pt': np.array([e[i] for i in corner]), 'x_dir': np.array(xd, dtype=np.float64), 'y_dir': np.array(yd, dtype=np.float64),
             'w': e[wi[0]] - e[wi[1]], 'h': e[hi[0]] - e[hi[1]], 'off_x': 0, 'off_y': 0} for corner, xd, yd, wi, hi in _FACES]


def lookat_rotation(delta):
    """utils.py:189-203 -- rotation R with R delta = (0, 0, |delta|): yaw about y, then pitch about x."""
    dx, dy, dz = (float(v) for v in np.asarray(delta).reshape(3))
    yaw, pitch = -math.atan2(dx, dz), math.asin(dy / math.sqrt(dx * dx + dy * dy + dz * dz))
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    r_yaw = np.array([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]])
    r_pitch = np.array([[1.0, 0.0, 0.0], [0.0, cp, -sp], [0.0, sp, cp]])
    return r_pitch @ r_yaw
