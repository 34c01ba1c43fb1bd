"""Sim3DR on the B200: vertex normals, lighting and z-buffer rasterisation.

Reference-shaped surface (``Sim3DR/Sim3DR.py:8-29``, ``Sim3DR/lighting.py:23-79``): ``get_normal``, ``rasterize``,
``RenderPipeline`` take and return numpy arrays like the reference's Cython module does.  Underneath sits
:class:`MeshRenderer`, which works on device tensors and on a whole BATCH of meshes per launch -- it reads the dense
vertices ``reconstruct_vertex_62(dense=True)`` / ``syn_reconstruct_image`` leave on the GPU in place (any strides), so
the 638 KB per face never visit the host.  All arithmetic runs in ``libsynergy_b200.so`` (``csrc/kernels_render.cuh``);
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _light_cfg(**kw) -> _lib.LightCfg:
    """Defaults of ``RenderPipeline.__init__`` (Sim3DR/lighting.py:24-32)."""
    v3 = lambda x: (C.c_float * 3)(*[float(t) for t in x])
    return _lib.LightCfg(float(kw.get('intensity_ambient', 0.3)), float(kw.get('intensity_directional', 0.6)),
                         float(kw.get('intensity_specular', 0.1)), v3(kw.get('color_ambient', (1, 1, 1))),
                         v3(kw.get('color_directional', (1, 1, 1))), v3(kw.get('light_pos', (0, 0, 5))),
                         v3(kw.get('view_pos', (0, 0, 5))), int(kw.get('specular_exp', 5)))


class MeshRenderer:
    """Batched renderer over one triangle list.

    ``triangles``: (ntri,3) integer array, 0-based (``utils/render.py:32-33``).  Vertex arguments are float32 CUDA
    tensors indexed ``v[b, i, k]`` = coordinate k of vertex i of mesh b with ANY strides: pass
    ``dense.transpose(1, 2)`` for the (B,3,N) output of the 3DMM stage (a view, nothing is copied).
    """

    def __init__(self, triangles, nver: int, device=None):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError('synergynet_b200.Sim3DR needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        tri = np.ascontiguousarray(np.asarray(triangles), dtype=np.int32)
        if tri.ndim != 2 or tri.shape[1] != 3:
            raise ValueError('triangles must be (ntri, 3)')
        self.nver, self.ntri = int(nver), int(tri.shape[0])
        start = np.zeros(self.nver + 1, np.int32)
        lst = np.zeros(max(3 * self.ntri, 1), np.int32)
        _lib.check(self._lib.syn_mesh_incidence_host(tri.ctypes.data, self.ntri, self.nver, start.ctypes.data, lst.ctypes.data))
        self.tri = torch.from_numpy(tri).to(self.device)
        self._inc_start = torch.from_numpy(start).to(self.device)
        self._inc_tri = torch.from_numpy(lst).to(self.device)
        self.launches = 0

    # -- helpers ----------------------------------------------------------------------------------------------------------
    def _view(self, vertices: torch.Tensor):
        if vertices.dim() == 2:
            vertices = vertices.unsqueeze(0)
        if vertices.dtype != torch.float32 or vertices.device != self.device or vertices.dim() != 3 \
                or vertices.shape[1] != self.nver or vertices.shape[2] != 3:
            raise ValueError(f'vertices must be float32 (B,{self.nver},3) on {self.device}; got {tuple(vertices.shape)} '
                             f'{vertices.dtype} on {vertices.device}')
        sb, sv, sc = vertices.stride()
        if vertices.shape[0] == 1:
            sb = max(sb, 1)
        if min(sb, sv, sc) <= 0:
            raise ValueError('vertices must have positive strides (no expanded / flipped views)')
        return vertices, (vertices.data_ptr(), int(sb), int(sv), int(sc), int(vertices.shape[0]), self.nver)

    def normals(self, vertices: torch.Tensor) -> torch.Tensor:
        """``get_normal`` for every mesh: (B,nver,3) unit normals (NaN for a vertex no triangle touches, like the reference)."""
        v, view = self._view(vertices)
        b = view[4]
        ws = torch.empty((b, self.ntri, 3), dtype=torch.float32, device=self.device)
        out = torch.empty((b, self.nver, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.syn_mesh_normals(*view, self.tri.data_ptr(), self.ntri, self._inc_start.data_ptr(),
                                                  self._inc_tri.data_ptr(), ws.data_ptr(), out.data_ptr(), _stream_ptr(self.device)))
        self.launches += 2
        return out

    def colors(self, vertices: torch.Tensor, normals: torch.Tensor, cfg: Optional[_lib.LightCfg] = None,
               texture: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Per-vertex light of ``RenderPipeline.__call__`` (times ``texture`` (nver,3) if given): (B,nver,3) in [0,1]."""
        v, view = self._view(vertices)
        b = view[4]
        cfg = cfg or _light_cfg()
        normals = normals.contiguous()
        if tuple(normals.shape) != (b, self.nver, 3) or normals.dtype != torch.float32 or normals.device != self.device:
            raise ValueError('normals must be float32 (B,nver,3) on the renderer device')
        tex_ptr = None
        if texture is not None:
            texture = texture.to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(texture.shape) != (self.nver, 3):
                raise ValueError('texture must be (nver, 3)')
            tex_ptr = texture.data_ptr()
        stats = torch.empty((b, 6), dtype=torch.int32, device=self.device)
        out = torch.empty((b, self.nver, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.syn_mesh_lighting(*view, normals.data_ptr(), C.byref(cfg), tex_ptr, stats.data_ptr(),
                                                   out.data_ptr(), _stream_ptr(self.device)))
        self.launches += 2
        return out

    def rasterize(self, image: torch.Tensor, vertices: torch.Tensor, colors: torch.Tensor, reverse: bool = False,
                  return_depth: bool = False):
        """Draw the B meshes, in order, onto ``image`` (H,W,C) uint8 on the device, IN PLACE (alpha = 1)."""
        v, view = self._view(vertices)
        b = view[4]
        if image.dtype != torch.uint8 or image.dim() != 3 or image.device != self.device or not image.is_contiguous():
            raise ValueError('image must be a contiguous uint8 (H,W,C) tensor on the renderer device')
        h, w, c = (int(s) for s in image.shape)
        colors = colors.contiguous()
        if tuple(colors.shape) != (b, self.nver, c) or colors.dtype != torch.float32 or colors.device != self.device:
            raise ValueError(f'colors must be float32 (B,nver,{c}) on the renderer device')
        keys = torch.empty((b, h, w), dtype=torch.int64, device=self.device)
        depth = torch.empty((b, h, w), dtype=torch.float32, device=self.device) if return_depth else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.syn_rasterize(image.data_ptr(), h, w, c, *view, self.tri.data_ptr(), self.ntri, colors.data_ptr(),
                                               1.0, 1 if reverse else 0, keys.data_ptr(),
                                               depth.data_ptr() if depth is not None else None, _stream_ptr(self.device)))
        self.launches += 2
        return (image, depth) if return_depth else image

    def render(self, image: torch.Tensor, vertices: torch.Tensor, cfg: Optional[_lib.LightCfg] = None,
               texture: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``RenderPipeline.__call__`` for a batch of meshes drawn one after the other onto ``image`` (in place)."""
        nrm = self.normals(vertices)
        return self.rasterize(image, vertices, self.colors(vertices, nrm, cfg, texture))


_renderers = {}


def _renderer_for(triangles: np.ndarray, nver: int) -> MeshRenderer:
    if not torch.cuda.is_available():
        raise RuntimeError('synergynet_b200.Sim3DR needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    tri = np.ascontiguousarray(triangles, dtype=np.int32)
    key = (hash(tri.tobytes()), tri.shape, int(nver), torch.cuda.current_device())
    r = _renderers.get(key)
    if r is None:
        if len(_renderers) > 8:
            _renderers.clear()
        r = _renderers[key] = MeshRenderer(tri, nver)
    return r


def _verts_dev(vertices: np.ndarray, r: MeshRenderer) -> torch.Tensor:
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    if v.ndim != 2 or v.shape[1] != 3:
        raise ValueError('vertices must be (nver, 3)')
    return torch.from_numpy(v).to(r.device).unsqueeze(0)


def get_normal(vertices: np.ndarray, triangles: np.ndarray) -> np.ndarray:
    """``Sim3DR.get_normal`` (Sim3DR/Sim3DR.py:8-11): (nver,3) float32 vertices, (ntri,3) int32 triangles -> (nver,3)."""
    r = _renderer_for(triangles, vertices.shape[0])
    return r.normals(_verts_dev(vertices, r))[0].cpu().numpy()


def rasterize(vertices, triangles, colors, bg=None, height=None, width=None, channel=None, reverse=False):
    """``Sim3DR.rasterize`` (Sim3DR/Sim3DR.py:14-29).  ``bg`` (H,W,C) uint8 is drawn into and returned, as the reference's
    C routine writes through the array it is given.  Without ``bg`` the reference builds a float32 canvas that its own
    ``unsigned char`` Cython signature then rejects; here that case starts from a black uint8 canvas."""
    if bg is not None:
        height, width, channel = bg.shape
    else:
        assert height is not None and width is not None and channel is not None
        bg = np.zeros((height, width, channel), dtype=np.uint8)
    if bg.dtype != np.uint8:
        raise ValueError("Buffer dtype mismatch, expected 'unsigned char'")       # what the reference's Cython layer raises
    colors = np.ascontiguousarray(colors, dtype=np.float32)
    r = _renderer_for(triangles, vertices.shape[0])
    img = torch.from_numpy(np.ascontiguousarray(bg)).to(r.device)
    r.rasterize(img, _verts_dev(vertices, r), torch.from_numpy(colors).to(r.device).unsqueeze(0), reverse=reverse)
    out = img.cpu().numpy()
    if bg.flags.writeable and bg.flags.c_contiguous:
        bg[...] = out
        return bg
    return out


def convert_type(obj):
    if isinstance(obj, (tuple, list)):
        return np.array(obj, dtype=np.float32)[None, :]
    return obj


class RenderPipeline(object):
    """``Sim3DR.RenderPipeline`` (Sim3DR/lighting.py:23-79): same constructor keywords and call signature."""

    def __init__(self, **kwargs):
        self._kw = dict(kwargs)
        self.light_pos = convert_type(kwargs.get('light_pos', (0, 0, 5)))

    def update_light_pos(self, light_pos):
        self.light_pos = convert_type(light_pos)

    def _cfg(self) -> _lib.LightCfg:
        kw = dict(self._kw)
        kw['light_pos'] = tuple(np.asarray(self.light_pos, np.float32).reshape(-1))
        return _light_cfg(**kw)

    def __call__(self, vertices, triangles, bg, texture=None):
        r = _renderer_for(triangles, vertices.shape[0])
        v = _verts_dev(vertices, r)
        tex = None if texture is None else torch.from_numpy(np.ascontiguousarray(texture, dtype=np.float32))
        col = r.colors(v, r.normals(v), self._cfg(), tex)
        if texture is not None and isinstance(texture, np.ndarray) and texture.flags.writeable:
            texture[...] = col[0].cpu().numpy()                                    # `texture *= light` is in place (:73)
        if bg.dtype != np.uint8:
            raise ValueError("Buffer dtype mismatch, expected 'unsigned char'")
        img = torch.from_numpy(np.ascontiguousarray(bg)).to(r.device)
        r.rasterize(img, v, col)
        out = img.cpu().numpy()
        if bg.flags.writeable and bg.flags.c_contiguous:
            bg[...] = out
            return bg
        return out


def render(img: np.ndarray, ver_lst, tri, alpha: float = 0.6, wfp=None, tex=None, cfg: Optional[dict] = None):
    """``utils/render.py:31-53`` as one batched call: every (3,N) vertex array of ``ver_lst`` is lit and drawn, in order,
    onto a copy of ``img`` (the loop :41-45 becomes one launch sequence over the batch), then blended with
    ``cv2.addWeighted(img, 1 - alpha, overlap, alpha, 0)``.  ``tri``: (ntri,3) 0-based triangles (the reference re-reads
    ``3dmm_data/tri.mat`` on every call).  Returns ``(blended, overlap)``."""
    import cv2
    from .inference import RENDER_CFG
    ver = np.stack([np.asarray(v, dtype=np.float32) for v in ver_lst])             # (B,3,N)
    r = _renderer_for(tri, ver.shape[2])
    v = torch.from_numpy(ver).to(r.device).transpose(1, 2)                          # strided view, no transpose copy
    canvas = torch.from_numpy(np.ascontiguousarray(img)).to(r.device)
    texture = None if tex is None else torch.from_numpy(np.ascontiguousarray(tex, dtype=np.float32))
    r.render(canvas, v, _light_cfg(**(cfg or RENDER_CFG)), texture)
    overlap = canvas.cpu().numpy()
    res = cv2.addWeighted(img, 1 - alpha, overlap, alpha, 0)
    if wfp is not None:
        cv2.imwrite(wfp[:-4] + '_solid' + '.png', overlap)
        cv2.imwrite(wfp, res)
    return res, overlap
