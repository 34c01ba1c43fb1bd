"""Per-device handle of the sm_100a library: weight hand-over and the compute entry points.

PyTorch is used for device memory, streams and (in bench.py) ``torch.distributed`` only; all
arithmetic of the path runs in ``libsynergy_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .backbone import HEAD_DIMS, conv_plan

N_PARAMS = 62


def _host_f32(t) -> torch.Tensor:
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    return t.detach().to(device='cpu', dtype=torch.float32).contiguous()


class Engine:
    """Owns one ``syn_handle_t`` bound to ``cuda:<device>``."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError('synergynet_b200 needs a CUDA device (B200, sm_100a); there is no '
                               'CPU fallback for the inference hot path')
        self.device = torch.device('cuda', int(device))
        h = C.c_void_p()
        _lib.check(self._lib.syn_create(int(device), C.byref(h)))
        self._h = h
        self.n_pts = 0
        self.n_vert = 0
        self._keep = []

    def close(self):
        if getattr(self, '_h', None):
            self._lib.syn_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- weights --------------------------------------------------------------------------------
    def load_backbone(self, sd: Dict[str, torch.Tensor], prefix: str = 'I2P.backbone.') -> None:
        """Hand the 52 conv+BN pairs and the three heads of a reference-schema state dict
         to the library."""
        for spec in conv_plan():
            w = _host_f32(sd[f'{prefix}{spec.conv_key}.weight'])
            bn = [_host_f32(sd[f'{prefix}{spec.bn_key}.{k}'])
                  for k in ('weight', 'bias', 'running_mean', 'running_var')]
            _lib.check(self._lib.syn_set_conv_bn(self._h, spec.index, w.data_ptr(), w.numel(),
                                                 *[t.data_ptr() for t in bn], 1e-5))
        heads = []
        for name, _ in HEAD_DIMS:
            heads += [_host_f32(sd[f'{prefix}{name}.1.weight']), _host_f32(sd[f'{prefix}{name}.1.bias'])]
        _lib.check(self._lib.syn_set_heads(self._h, *[t.data_ptr() for t in heads]))

    def load_3dmm(self, param_mean, param_std, u_base, w_shp_base, w_exp_base, u=None, w_shp=None,
                  w_exp=None) -> None:
        mean, std = _host_f32(param_mean).reshape(-1)[:62].contiguous(), _host_f32(param_std).reshape(-1)[:62].contiguous()
        _lib.check(self._lib.syn_set_whitening(self._h, mean.data_ptr(), std.data_ptr()))
        ub, wsb, web = _host_f32(u_base), _host_f32(w_shp_base), _host_f32(w_exp_base)
        self.n_pts = ub.numel() // 3
        _lib.check(self._lib.syn_set_basis_sparse(self._h, ub.data_ptr(), wsb.data_ptr(), web.data_ptr(), self.n_pts))
        if u is not None:
            ud, wsd, wed = _host_f32(u), _host_f32(w_shp), _host_f32(w_exp)
            self.n_vert = ud.numel() // 3
            _lib.check(self._lib.syn_set_basis_dense(self._h, ud.data_ptr(), wsd.data_ptr(), wed.data_ptr(), self.n_vert))

    def commit(self) -> None:
        _lib.check(self._lib.syn_commit(self._h))

    def set_engine(self, engine: int) -> None:
        _lib.check(self._lib.syn_set_engine(self._h, int(engine)))

    @property
    def engine(self) -> int:
        return self._lib.syn_get_engine(self._h)

    @property
    def launch_count(self) -> int:
        return int(self._lib.syn_launch_count(self._h))

    def set_timing(self, on: bool) -> None:
        _lib.check(self._lib.syn_set_timing(self._h, int(on)))

    def timings(self):
        """[(kernel label, ms)] of the last device-buffer call (needs set_timing(True) before it)."""
        ms = (C.c_float * 64)()
        names = (C.c_char_p * 64)()
        n = C.c_int(0)
        _lib.check(self._lib.syn_get_timings(self._h, ms, names, 64, C.byref(n)))
        return [(names[i].decode(), float(ms[i])) for i in range(n.value)]

    def poll_error(self) -> int:
        """Device sync + sticky in-kernel timeout flag (0 = clean)."""
        flag = C.c_int(0)
        _lib.check(self._lib.syn_poll_error(self._h, C.byref(flag)))
        return int(flag.value)

    # ---- compute --------------------------------------------------------------------------------
    def _check_x(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 120, 120):
            raise RuntimeError(f'expected (B,3,120,120) input, got {tuple(x.shape)}')
        if x.device != self.device:
            raise RuntimeError(f'input on {x.device}, engine on {self.device}')
        return x.to(torch.float32).contiguous()

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward(self, x: torch.Tensor, want_pool: bool = False):
        x = self._check_x(x)
        b = x.shape[0]
        params = torch.empty((b, N_PARAMS), device=self.device, dtype=torch.float32)
        pool = torch.empty((b, 1280), device=self.device, dtype=torch.float32) if want_pool else None
        _lib.check(self._lib.syn_forward(self._h, x.data_ptr(), b, params.data_ptr(),
                                         pool.data_ptr() if want_pool else None, self._stream()))
        return (params, pool) if want_pool else params

    def reconstruct(self, params: torch.Tensor, dense: bool = False, whitening: bool = True,
                    transform: bool = True) -> torch.Tensor:
        if params.dim() != 2 or params.shape[1] != N_PARAMS:
            raise RuntimeError('length of params mismatch')          # model_building.py:116-119
        params = params.to(device=self.device, dtype=torch.float32).contiguous()
        b = params.shape[0]
        n = self.n_vert if dense else self.n_pts
        if n == 0:
            raise RuntimeError('dense basis not loaded' if dense else 'sparse basis not loaded')
        out = torch.empty((b, 3, n), device=self.device, dtype=torch.float32)
        _lib.check(self._lib.syn_reconstruct(self._h, params.data_ptr(), b, int(dense), int(whitening),
                                             int(transform), out.data_ptr(), self._stream()))
        return out

    def forward_landmarks(self, x: torch.Tensor, want_params: bool = False):
        """x: fp32 normalised crops, or raw uint8 crops (normalised on the device)."""
        if x.dtype == torch.uint8:
            return self._forward_landmarks_u8(x, want_params)
        x = self._check_x(x)
        b = x.shape[0]
        lmk = torch.empty((b, 3, self.n_pts), device=self.device, dtype=torch.float32)
        params = torch.empty((b, N_PARAMS), device=self.device, dtype=torch.float32) if want_params else None
        _lib.check(self._lib.syn_forward_landmarks(self._h, x.data_ptr(), b,
                                                   params.data_ptr() if want_params else None,
                                                   lmk.data_ptr(), self._stream()))
        return (lmk, params) if want_params else lmk

    def _forward_landmarks_u8(self, x: torch.Tensor, want_params: bool):
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 120, 120) or x.device != self.device:
            raise RuntimeError(f'expected uint8 (B,3,120,120) on {self.device}, got {tuple(x.shape)} on {x.device}')
        x = x.contiguous()
        b = x.shape[0]
        lmk = torch.empty((b, 3, self.n_pts), device=self.device, dtype=torch.float32)
        params = torch.empty((b, N_PARAMS), device=self.device, dtype=torch.float32) if want_params else None
        _lib.check(self._lib.syn_forward_landmarks_u8(self._h, x.data_ptr(), b,
                                                      params.data_ptr() if want_params else None,
                                                      lmk.data_ptr(), self._stream()))
        return (lmk, params) if want_params else lmk

    def forward_landmarks_host(self, x_host: torch.Tensor, lmk_host: Optional[torch.Tensor] = None,
                               params_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """End-to-end call on HOST tensors (pinned recommended): H2D, forward, landmarks, D2H."""
        if x_host.is_cuda or x_host.dtype not in (torch.float32, torch.uint8) or not x_host.is_contiguous():
            raise RuntimeError('x_host must be a contiguous fp32 (normalised) or uint8 (raw) CPU tensor')
        b = x_host.shape[0]
        if lmk_host is None:
            lmk_host = torch.empty((b, 3, self.n_pts), dtype=torch.float32)
        fn = self._lib.syn_forward_landmarks_host_u8 if x_host.dtype == torch.uint8 else self._lib.syn_forward_landmarks_host
        _lib.check(fn(
            self._h, x_host.data_ptr(), b, params_host.data_ptr() if params_host is not None else None,
            lmk_host.data_ptr()))
        return lmk_host

    def debug_forward_until(self, x: torch.Tensor, layer: int) -> torch.Tensor:
        x = self._check_x(x)
        spec = conv_plan()[layer]
        out = torch.empty((x.shape[0], spec.h_out, spec.h_out, spec.cout), device=self.device, dtype=torch.float32)
        _lib.check(self._lib.syn_debug_forward_until(self._h, x.data_ptr(), x.shape[0], layer,
                                                     out.data_ptr(), self._stream()))
        return out
