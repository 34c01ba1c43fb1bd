"""The FaceBoxes detector on the B200: network, box decode and NMS all on the device.

Reference-shaped surface: :class:`FaceBoxes` has the constructor and call signature of ``FaceBoxes/FaceBoxes.py:46-143``
(``FaceBoxes(timer_flag=False)``, ``face_boxes(img_bgr_uint8) -> [[xmin, ymin, xmax, ymax, score], ...]``) and loads the
reference's checkpoint schema (``FaceBoxes/models/faceboxes.py``: ``conv1.conv.weight``, ``inception2.branch3x3.bn.*``,
``loc.0.bias`` ..., optional ``module.`` prefix, ``utils/functions.py:19-43``).  The 33 convolutions, the pools and the
softmax run in ``libsynergy_b200.so`` (``csrc/kernels_detect.cuh``); only ``cv2.resize`` of oversized images stays on
the host, as in the reference.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, detect

# FaceBoxes/FaceBoxes.py:24-25
scale_flag = True
HEIGHT, WIDTH = 720, 1080


def layer_plan() -> List[dict]:
    """The 33 convolutions in execution order as the library reports them (name = state_dict prefix)."""
    lib = _lib.load()
    out = []
    for i in range(lib.syn_fb_num_layers()):
        d = _lib.FbLayerDesc()
        _lib.check(lib.syn_fb_layer_desc(i, C.byref(d)))
        out.append(dict(index=i, name=d.name.decode(), cin=d.cin, cout=d.cout, ksize=d.ksize, stride=d.stride, pad=d.pad,
                        has_bn=bool(d.has_bn), activation=d.activation))
    return out


def state_dict_keys() -> List[str]:
    """Keys of a ``FaceBoxesNet`` state dict (faceboxes.py:8-18, 50-58, 94-106)."""
    keys = []
    for L in layer_plan():
        if L['has_bn']:
            keys.append(f"{L['name']}.conv.weight")
            keys += [f"{L['name']}.bn.{k}" for k in ('weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked')]
        else:
            keys += [f"{L['name']}.weight", f"{L['name']}.bias"]
    return keys


class FaceBoxesNet:
    """Device-side detector network: one ``syn_fb_t`` handle."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device=None):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError('synergynet_b200.faceboxes needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        _lib.check(self._lib.syn_fb_create(self.device.index or 0, C.byref(h)))
        self._h = h
        sd = {(k[7:] if k.startswith('module.') else k): v for k, v in state_dict.items()}       # utils/functions.py:19-24
        f32 = lambda t: torch.as_tensor(t).detach().to(device='cpu', dtype=torch.float32).contiguous()
        for L in layer_plan():
            n = L['name']
            if L['has_bn']:
                w = f32(sd[f'{n}.conv.weight'])
                bn = [f32(sd[f'{n}.bn.{k}']) for k in ('weight', 'bias', 'running_mean', 'running_var')]
                _lib.check(self._lib.syn_fb_set_layer(self._h, L['index'], w.data_ptr(), w.numel(), None,
                                                      *[t.data_ptr() for t in bn], 1e-5))
            else:
                w, b = f32(sd[f'{n}.weight']), f32(sd[f'{n}.bias'])
                _lib.check(self._lib.syn_fb_set_layer(self._h, L['index'], w.data_ptr(), w.numel(), b.data_ptr(), None, None, None, None, 0.0))
        _lib.check(self._lib.syn_fb_commit(self._h))

    def close(self):
        if getattr(self, '_h', None):
            self._lib.syn_fb_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.syn_fb_launch_count(self._h))

    def forward(self, image: torch.Tensor):
        """``image`` (H,W,3) uint8 BGR on the device -> ``(loc (P,4), conf (P,2))`` like ``FaceBoxesNet.forward`` in 'test'
        phase (faceboxes.py:112-150) applied to ``img - (104,117,123)``."""
        if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3 or image.device != self.device or not image.is_contiguous():
            raise ValueError('image must be a contiguous uint8 (H,W,3) tensor on the detector device')
        h, w = int(image.shape[0]), int(image.shape[1])
        p = detect.num_priors(h, w)
        loc = torch.empty((p, 4), dtype=torch.float32, device=self.device)
        conf = torch.empty((p, 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.syn_fb_forward(self._h, image.data_ptr(), h, w, loc.data_ptr(), conf.data_ptr(),
                                                torch.cuda.current_stream(self.device).cuda_stream))
        return loc, conf


class FaceBoxes:
    """``FaceBoxes.FaceBoxes`` (FaceBoxes/FaceBoxes.py:46-143).  ``weights``: a state dict, a checkpoint path, or None for
    ``weights/FaceBoxesProd.pth`` next to this module (where the reference keeps it)."""

    def __init__(self, timer_flag: bool = False, weights=None, device=None):
        if weights is None:
            weights = os.path.join(os.path.dirname(os.path.realpath(__file__)), 'weights', 'FaceBoxesProd.pth')
        if isinstance(weights, (str, os.PathLike)):
            weights = torch.load(weights, map_location='cpu')
            if 'state_dict' in weights:
                weights = weights['state_dict']                                               # utils/functions.py:34-37
        self.net = FaceBoxesNet(weights, device)
        self.timer_flag = timer_flag

    def __call__(self, img_: np.ndarray):
        import cv2
        img_raw = img_
        scale = 1
        if scale_flag:                                                                        # FaceBoxes.py:62-79
            h, w = img_raw.shape[:2]
            if h > HEIGHT:
                scale = HEIGHT / h
            if w * scale > WIDTH:
                scale *= WIDTH / (w * scale)
            if scale != 1:
                img_raw = cv2.resize(img_raw, dsize=(int(scale * w), int(scale * h)))
        im_h, im_w = img_raw.shape[:2]
        image = torch.from_numpy(np.ascontiguousarray(img_raw, dtype=np.uint8)).to(self.net.device)
        loc, conf = self.net.forward(image)
        dets, n = detect.decode_device(loc, conf, im_h, im_w, scale=float(scale))
        n_host = int(n.item())
        if n_host == 0:
            return []
        keep, n_keep = detect.nms_device(dets, detect.nms_threshold, _lib.NMS_CPU_NMS, n=n_host)
        kept = dets[keep[:int(n_keep.item())].long()][:detect.keep_top_k].cpu().numpy()
        return [[b[0], b[1], b[2], b[3], b[4]] for b in kept if b[4] > detect.vis_thres]
