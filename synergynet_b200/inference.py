"""Host-side API semantics around the hot path, batched (reference ``utils/inference.py``).

Only tiny per-face affine / pose algebra and the integer ROI crop live here; vertex
reconstruction itself runs on the GPU (``Engine.reconstruct``).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

STD_SIZE = 120


def parse_param(param: np.ndarray):
    """Slices of one de-whitened 62-vector (utils/inference.py:25-31)."""
    cam = param[:12].reshape(3, 4)
    return cam[:, :3], cam[:, 3:4], param[12:52].reshape(40, 1), param[52:62].reshape(10, 1)


def crop_img(img: np.ndarray, roi_box: Sequence[float]) -> np.ndarray:
    """Integer-rounded ROI crop with zero fill outside the image (utils/inference.py:95-125).
    Index arithmetic is bit-exact with the reference: Python ``round`` then clamping."""
    img_h, img_w = img.shape[:2]
    x0, y0, x1, y1 = (int(round(v)) for v in roi_box[:4])
    out = np.zeros((y1 - y0, x1 - x0) + tuple(img.shape[2:]), dtype=np.uint8)
    src_x0, src_y0 = max(x0, 0), max(y0, 0)
    src_x1, src_y1 = min(x1, img_w), min(y1, img_h)
    dst_x0, dst_y0 = src_x0 - x0, src_y0 - y0
    dst_x1 = (x1 - x0) - (x1 - src_x1)
    dst_y1 = (y1 - y0) - (y1 - src_y1)
    out[dst_y0:dst_y1, dst_x0:dst_x1] = img[src_y0:src_y1, src_x0:src_x1]
    return out


def square_roi(rect: Sequence[float]) -> list:
    """Enlarged square box around a detection (synergy3DMM.py:181-185): side = 1.2 x height,
    ``//`` floor division as in the reference."""
    h_center = (rect[1] + rect[3]) / 2
    w_center = (rect[0] + rect[2]) / 2
    margin = (rect[3] - rect[1]) * 1.2 // 2
    tail = list(rect[4:]) if len(rect) > 4 else [1.0]
    return [w_center - margin, h_center - margin, w_center + margin, h_center + margin] + tail


def rescale_vertices(vertex: np.ndarray, roi_box: Sequence[float]) -> np.ndarray:
    """Crop -> image coordinates for one (3,N) array (utils/inference.py:127-138)."""
    sx, sy, ex, ey = roi_box[:4]
    kx, ky = (ex - sx) / STD_SIZE, (ey - sy) / STD_SIZE
    out = np.array(vertex, copy=True)
    out[0] = out[0] * kx + sx
    out[1] = out[1] * ky + sy
    out[2] *= (kx + ky) / 2
    return out


def roi_affine(roi_boxes: Sequence[Sequence[float]]) -> np.ndarray:
    """(B,5) fp32 rows kx, sx, ky, sy, kz of the crop -> image map of ``_predict_vertices`` / ``predict_pose``
    (utils/inference.py:129-136,150-154).  The reference evaluates the scales as Python floats (double) and numpy
    rounds them to fp32 when they meet the fp32 vertex arrays; the same happens here, once per face, as index-like host
    work -- the per-vertex arithmetic runs on the GPU (``Engine.reconstruct_image``)."""
    out = np.empty((len(roi_boxes), 5), np.float32)
    for i, box in enumerate(roi_boxes):
        sx, sy, ex, ey = box[:4]
        kx, ky = (ex - sx) / STD_SIZE, (ey - sy) / STD_SIZE
        out[i] = (kx, sx, ky, sy, (kx + ky) / 2)
    return out


def decompose_camera(P: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Batched P2sRt (utils/inference.py:33-43): P (B,3,4) -> scale (B,), R (B,3,3), t (B,3)."""
    r1, r2 = P[:, 0, :3], P[:, 1, :3]
    n1 = np.linalg.norm(r1, axis=1, keepdims=True)
    n2 = np.linalg.norm(r2, axis=1, keepdims=True)
    u1, u2 = r1 / n1, r2 / n2
    R = np.stack([u1, u2, np.cross(u1, u2)], axis=1)
    return ((n1 + n2) / 2.0)[:, 0], R, P[:, :, 3]


def rotation_to_euler_deg(R: np.ndarray) -> np.ndarray:
    """Batched matrix2angle_corr (utils/inference.py:45-62) -> (B,3) degrees."""
    r20 = R[:, 2, 0]
    lock = (r20 == 1) | (r20 == -1)
    x = np.arcsin(np.clip(r20, -1, 1))
    cx = np.where(lock, 1.0, np.cos(x))
    y = np.arctan2(R[:, 1, 2] / cx, R[:, 2, 2] / cx)
    z = np.arctan2(R[:, 0, 1] / cx, R[:, 0, 0] / cx)
    neg = r20 == -1
    x = np.where(lock, np.where(neg, np.pi / 2, -np.pi / 2), x)
    y = np.where(lock, np.where(neg, np.arctan2(R[:, 0, 1], R[:, 0, 2]),
                                np.arctan2(-R[:, 0, 1], -R[:, 0, 2])), y)
    z = np.where(lock, 0.0, z)
    return np.stack([x, y, z], 1) * (180.0 / np.pi)


def predict_pose_batch(params: np.ndarray, param_mean: np.ndarray, param_std: np.ndarray,
                       roi_boxes: Sequence[Sequence[float]]):
    """parse_pose + predict_pose (utils/inference.py:86-92,146-157) for B whitened vectors.
    Returns a list of ``[angles(list of 3), t3d(ndarray 3)]`` like the reference."""
    p = params * param_std[:62] + param_mean[:62]
    cam = p[:, :12].reshape(-1, 3, 4)
    _, R, t3d = decompose_camera(cam)
    ang = rotation_to_euler_deg(R)
    out = []
    for i, box in enumerate(roi_boxes):
        sx, sy, ex, ey = box[:4]
        t = t3d[i].copy()
        t[0] = t[0] * ((ex - sx) / STD_SIZE) + sx
        t[1] = t[1] * ((ey - sy) / STD_SIZE) + sy
        out.append([[float(a) for a in ang[i]], t])
    return out


# lighting of the solid-mesh overlay (utils/render.py:18-27), consumed by synergynet_b200.Sim3DR.render
RENDER_CFG = {
    'intensity_ambient': 0.75, 'color_ambient': (1, 1, 1),
    'intensity_directional': 0.7, 'color_directional': (1, 1, 1),
    'intensity_specular': 0.2, 'specular_exp': 5,
    'light_pos': (0, 0, 5), 'view_pos': (0, 0, 5),
}
