"""TEST INFRASTRUCTURE ONLY -- oracle of the stages either side of the 3DMM path.

Nothing under ``synergynet_b200/`` may import this module (``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline legs do).  It holds

* ctypes doors to ``oracle/libsim3dr_port.so`` (this project's C restatement, ``oracle/sim3dr_port.c``) and, when it has
  been built, ``oracle/_ref/libsim3dr_ref.so`` (the reference's own ``Sim3DR/lib/rasterize_kernel.cpp`` compiled where
  it lies, ``oracle/Makefile``);
* a numpy restatement of ``RenderPipeline.__call__`` (``Sim3DR/lighting.py:37-75``) and of ``utils/render.py:31-53``;
* a numpy / torch restatement of the FaceBoxes post-processing (``FaceBoxes/utils/prior_box.py:12-48``,
  ``utils/box_utils.py:177-195``, ``FaceBoxes.py:98-127``, ``utils/nms/py_cpu_nms.py:10-38``).

Pinned by ``tests/test_oracle_render.py`` against ``tests/golden/render_vectors.npz`` (recorded from the unmodified
reference by ``tests/golden/make_golden_render.py``) and against ``_ref`` where present.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(_DIR, 'libsim3dr_port.so')
REF_LIB = os.path.join(_DIR, '_ref', 'libsim3dr_ref.so')


def build(force: bool = False) -> None:
    """Compile the C restatement (and ``_ref`` when /root/reference exists).  Building the checker is not using it."""
    src = os.path.join(_DIR, 'sim3dr_port.c')
    if force or not os.path.exists(PORT_LIB) or os.path.getmtime(PORT_LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-C', _DIR, 'libsim3dr_port.so'], check=True, capture_output=True)
    if os.path.isdir('/root/reference/Sim3DR/lib') and (force or not os.path.exists(REF_LIB)):
        subprocess.run(['make', '-C', _DIR, 'ref'], check=True, capture_output=True)


_libs = {}


def _lib(kind: str):
    if kind not in _libs:
        path = PORT_LIB if kind == 'port' else REF_LIB
        if kind == 'port':
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _libs[kind] = C.CDLL(path)
    return _libs[kind]


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def get_normal(vertices: np.ndarray, triangles: np.ndarray, kind: str = 'port') -> np.ndarray:
    """``Sim3DR.get_normal`` (Sim3DR/Sim3DR.py:8-11): vertices (nver,3) f32, triangles (ntri,3) int32 -> (nver,3)."""
    v = np.ascontiguousarray(vertices, np.float32)
    t = np.ascontiguousarray(triangles, np.int32)
    out = np.zeros_like(v)
    fn = getattr(_lib(kind), 'port_get_normal' if kind == 'port' else 'ref_get_normal')
    fn(_p(out), _p(v), _p(t), C.c_int(v.shape[0]), C.c_int(t.shape[0]))
    return out


def rasterize(vertices, triangles, colors, bg, reverse: bool = False, alpha: float = 1.0, kind: str = 'port', return_depth=False):
    """``Sim3DR.rasterize`` (Sim3DR/Sim3DR.py:14-29) on a uint8 background; draws into and returns ``bg``."""
    v = np.ascontiguousarray(vertices, np.float32)
    t = np.ascontiguousarray(triangles, np.int32)
    col = np.ascontiguousarray(colors, np.float32)
    assert bg.dtype == np.uint8 and bg.flags.c_contiguous
    h, w, c = bg.shape
    depth = np.zeros((h, w), np.float32) - 1e8
    fn = getattr(_lib(kind), 'port_rasterize' if kind == 'port' else 'ref_rasterize')
    fn(_p(bg), _p(v), _p(t), _p(col), _p(depth), C.c_int(t.shape[0]), C.c_int(h), C.c_int(w), C.c_int(c), C.c_float(alpha),
       C.c_int(1 if reverse else 0))
    return (bg, depth) if return_depth else bg


# ---- lighting (Sim3DR/lighting.py) ---------------------------------------------------------------------------------------
RENDER_CFG = dict(intensity_ambient=0.75, color_ambient=(1, 1, 1), intensity_directional=0.7, color_directional=(1, 1, 1),
                  intensity_specular=0.2, specular_exp=5, light_pos=(0, 0, 5), view_pos=(0, 0, 5))   # utils/render.py:18-27


def _row(x):
    return np.array(x, np.float32)[None, :] if isinstance(x, (tuple, list)) else x


def _unit(a):
    return a / np.sqrt(np.sum(a ** 2, axis=1))[:, None]


def lighting(vertices: np.ndarray, normal: np.ndarray, cfg: dict = RENDER_CFG) -> np.ndarray:
    """Per-vertex light of ``RenderPipeline.__call__`` (lighting.py:40-66), float32 numpy like the reference."""
    ia, idr, isp = cfg.get('intensity_ambient', 0.3), cfg.get('intensity_directional', 0.6), cfg.get('intensity_specular', 0.1)
    ca, cd = _row(cfg.get('color_ambient', (1, 1, 1))), _row(cfg.get('color_directional', (1, 1, 1)))
    lp, vp = _row(cfg.get('light_pos', (0, 0, 5))), _row(cfg.get('view_pos', (0, 0, 5)))
    expo = cfg.get('specular_exp', 5)
    light = np.zeros_like(vertices, dtype=np.float32)
    if ia > 0:
        light += ia * ca
    vn = vertices.astype(np.float32).copy()          # norm_vertices, lighting.py:9-14
    vn -= vn.min(0)[None, :]
    vn /= vn.max()
    vn *= 2
    vn -= vn.max(0)[None, :] / 2
    if idr > 0:
        direction = _unit(lp - vn)
        cos = np.sum(normal * direction, axis=1)[:, None]
        light += idr * (cd * np.clip(cos, 0, 1))
        if isp > 0:
            v2v = _unit(vp - vn)
            reflection = 2 * cos * normal - direction
            spe = np.sum((v2v * reflection) ** expo, axis=1)[:, None]
            spe = np.where(cos != 0, np.clip(spe, 0, 1), np.zeros_like(spe))
            light += isp * cd * np.clip(spe, 0, 1)
    return np.clip(light, 0, 1)


def render_faces(img: np.ndarray, ver_lst, tri: np.ndarray, cfg: dict = RENDER_CFG, kind: str = 'port') -> np.ndarray:
    """The loop of ``utils/render.py:40-45``: every (3,N) vertex array lit and drawn, in order, onto a copy of ``img``."""
    overlap = img.copy()
    for ver_ in ver_lst:
        ver = np.ascontiguousarray(ver_.astype(np.float32).T)
        normal = get_normal(ver, tri, kind)
        overlap = rasterize(ver, tri, lighting(ver, normal, cfg), overlap, kind=kind)
    return overlap


# ---- FaceBoxes post-processing -------------------------------------------------------------------------------------------
FB_MIN_SIZES, FB_STEPS, FB_VARIANCE = [[32, 64, 128], [256], [512]], [32, 64, 128], [0.1, 0.2]   # utils/config.py


def prior_boxes(im_h: int, im_w: int) -> np.ndarray:
    """``PriorBox(image_size).forward()`` (prior_box.py:22-48): Python-double arithmetic, one rounding to float32."""
    out = []
    for k, step in enumerate(FB_STEPS):
        rows, cols = math.ceil(im_h / step), math.ceil(im_w / step)
        for i in range(rows):
            for j in range(cols):
                for ms in FB_MIN_SIZES[k]:
                    sub = {32: (0, 0.25, 0.5, 0.75), 64: (0, 0.5)}.get(ms, (0.5,))
                    for dy in sub:
                        for dx in sub:
                            out += [(j + dx) * step / im_w, (i + dy) * step / im_h, ms / im_w, ms / im_h]
    return np.array(out, np.float64).astype(np.float32).reshape(-1, 4)


def decode_boxes(loc, priors):
    """``decode`` (box_utils.py:177-195) with torch float32 CPU arithmetic, variances 0.1 / 0.2."""
    import torch
    loc, priors = torch.as_tensor(loc), torch.as_tensor(priors)
    boxes = torch.cat((priors[:, :2] + loc[:, :2] * FB_VARIANCE[0] * priors[:, 2:],
                       priors[:, 2:] * torch.exp(loc[:, 2:] * FB_VARIANCE[1])), 1)
    boxes[:, :2] -= boxes[:, 2:] / 2
    boxes[:, 2:] += boxes[:, :2]
    return boxes


def faceboxes_dets(loc, conf, im_h, im_w, scale=1.0, conf_thresh=0.05, top_k=5000):
    """FaceBoxes.py:98-121: decoded, rescaled, thresholded boxes in descending score order as (n,5) float32 rows."""
    import torch
    boxes = decode_boxes(loc, prior_boxes(im_h, im_w))
    boxes = (boxes * torch.Tensor([im_w, im_h, im_w, im_h]) / scale / 1).numpy()
    scores = np.asarray(conf)[:, 1]
    inds = np.where(scores > conf_thresh)[0]
    boxes, scores = boxes[inds], scores[inds]
    order = np.argsort(scores, kind='stable')[::-1][:top_k]     # the reference's unstable argsort leaves ties undefined
    return np.hstack((boxes[order], scores[order][:, None])).astype(np.float32, copy=False)


def py_cpu_nms(dets: np.ndarray, thresh: float):
    """Restatement of utils/nms/py_cpu_nms.py:10-38 (vectorised numpy float32, keeps ``ovr <= thresh``)."""
    x1, y1, x2, y2, s = (dets[:, k] for k in range(5))
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    todo = s.argsort()[::-1]
    kept = []
    while todo.size:
        i, rest = todo[0], todo[1:]
        kept.append(int(i))
        w = np.maximum(0.0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + 1)
        h = np.maximum(0.0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + 1)
        inter = w * h
        todo = rest[inter / (area[i] + area[rest] - inter) <= thresh]
    return kept


def cpu_nms(dets: np.ndarray, thresh: float, ge: bool = True):
    """C restatement of utils/nms/cpu_nms.pyx:17-68 (``ge``) or of py_cpu_nms' comparison (``not ge``)."""
    d = np.ascontiguousarray(dets, np.float32)
    n = d.shape[0]
    order = np.ascontiguousarray(d[:, 4].argsort()[::-1], np.int64)
    keep = np.zeros(max(n, 1), np.int64)
    fn = _lib('port').port_nms
    fn.restype = C.c_int
    cnt = fn(_p(d), _p(order), C.c_int(n), C.c_double(thresh), C.c_int(1 if ge else 0), _p(keep))
    return [int(k) for k in keep[:cnt]]


# ---- the detector network (FaceBoxes/models/faceboxes.py) ----------------------------------------------------------------
def faceboxes_forward(sd, img_u8: np.ndarray):
    """``FaceBoxesNet(phase='test').forward`` (faceboxes.py:112-150) on ``img - (104,117,123)`` (FaceBoxes.py:86-96), restated
    with torch.nn.functional on CPU fp32 -- the ATen kernels the reference's modules dispatch to.  ``sd``: reference-schema
    state dict.  Returns ``(loc (P,4), conf (P,2))`` numpy arrays."""
    import torch
    import torch.nn.functional as F

    def bn_conv(x, name, stride=1, padding=0):
        y = F.conv2d(x, sd[f'{name}.conv.weight'], None, stride, padding)
        return F.batch_norm(y, sd[f'{name}.bn.running_mean'], sd[f'{name}.bn.running_var'], sd[f'{name}.bn.weight'],
                            sd[f'{name}.bn.bias'], False, 0.0, 1e-5)

    def basic(x, name, stride=1, padding=0):                       # BasicConv2d :8-18
        return F.relu(bn_conv(x, name, stride, padding))

    def crelu(x, name, stride, padding):                           # CRelu :50-64
        y = bn_conv(x, name, stride, padding)
        return F.relu(torch.cat([y, -y], 1))

    def inception(x, p):                                           # Inception :21-47
        a = basic(x, f'{p}.branch1x1')
        b = basic(F.avg_pool2d(x, 3, 1, 1), f'{p}.branch1x1_2')
        c = basic(basic(x, f'{p}.branch3x3_reduce'), f'{p}.branch3x3', 1, 1)
        d = basic(basic(basic(x, f'{p}.branch3x3_reduce_2'), f'{p}.branch3x3_2', 1, 1), f'{p}.branch3x3_3', 1, 1)
        return torch.cat([a, b, c, d], 1)

    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray((np.float32(img_u8) - (104, 117, 123)).transpose(2, 0, 1), dtype=np.float32))[None]
        x = F.max_pool2d(crelu(x, 'conv1', 4, 3), 3, 2, 1)
        x = F.max_pool2d(crelu(x, 'conv2', 2, 2), 3, 2, 1)
        for p in ('inception1', 'inception2', 'inception3'):
            x = inception(x, p)
        s0 = x
        s1 = basic(basic(s0, 'conv3_1'), 'conv3_2', 2, 1)
        s2 = basic(basic(s1, 'conv4_1'), 'conv4_2', 2, 1)
        loc, conf = [], []
        for k, s in enumerate((s0, s1, s2)):
            loc.append(F.conv2d(s, sd[f'loc.{k}.weight'], sd[f'loc.{k}.bias'], 1, 1).permute(0, 2, 3, 1).reshape(-1))
            conf.append(F.conv2d(s, sd[f'conf.{k}.weight'], sd[f'conf.{k}.bias'], 1, 1).permute(0, 2, 3, 1).reshape(-1))
        loc = torch.cat(loc).view(-1, 4)
        conf = torch.softmax(torch.cat(conf).view(-1, 2), dim=-1)
    return loc.numpy(), conf.numpy()


def faceboxes_detect(sd, img_u8: np.ndarray, conf_thresh=0.05, top_k=5000, nms_thresh=0.3, keep_top_k=750, vis_thres=0.5):
    """``FaceBoxes.__call__`` (FaceBoxes.py:57-143) for an image that needs no rescaling; NMS = py_cpu_nms' convention with the
    `>=` of the Cython path (identical unless an overlap equals the threshold)."""
    loc, conf = faceboxes_forward(sd, img_u8)
    h, w = img_u8.shape[:2]
    d = faceboxes_dets(loc, conf, h, w, 1.0, conf_thresh, top_k)
    keep = cpu_nms(d, nms_thresh) if d.shape[0] else []
    kept = d[keep][:keep_top_k]
    return [list(b) for b in kept if b[4] > vis_thres]
