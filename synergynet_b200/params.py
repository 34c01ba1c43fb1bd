"""3DMM basis loader: same files, attribute names and gathers as the reference ``ParamsPack``
(utils/params.py:8-37), plus an in-memory constructor for the seeded synthetic model."""
from __future__ import annotations

import os
import pickle
from typing import Dict, Optional

import numpy as np

_ENV = 'SYNERGY_3DMM_DIR'


def default_data_dir() -> str:
    return os.environ.get(_ENV) or os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '3dmm_data')


class ParamsPack:
    """keypoints, w_shp, w_exp, param_mean/std, u, and the 68-landmark gathers (index work is
    bit-exact with the reference: plain fancy indexing with the stored ``keypoints``)."""

    std_size = 120

    def __init__(self, data_dir: Optional[str] = None, arrays: Optional[Dict[str, np.ndarray]] = None):
        try:
            a = arrays if arrays is not None else self._read(data_dir or default_data_dir())
            self.keypoints = np.asarray(a['keypoints'])
            self.w_shp = np.asarray(a['w_shp'], np.float32)
            self.w_exp = np.asarray(a['w_exp'], np.float32)
            self.param_mean = np.asarray(a['param_mean'], np.float32)
            self.param_std = np.asarray(a['param_std'], np.float32)
            self.u_shp = np.asarray(a['u_shp'], np.float32)
            self.u_exp = np.asarray(a['u_exp'], np.float32)
            self.tri = a.get('tri')
            self.u = self.u_shp + self.u_exp
            kp = self.keypoints
            self.u_base = self.u[kp].reshape(-1, 1)
            self.w_shp_base = self.w_shp[kp]
            self.w_exp_base = self.w_exp[kp]
            self.dim = self.w_shp.shape[0] // 3
        except Exception as exc:
            raise RuntimeError('Missing data') from exc           # utils/params.py:36-37

    @staticmethod
    def _read(d: str) -> Dict[str, np.ndarray]:
        out = {
            'keypoints': np.load(os.path.join(d, 'keypoints_sim.npy')),
            'w_shp': np.load(os.path.join(d, 'w_shp_sim.npy')),
            'w_exp': np.load(os.path.join(d, 'w_exp_sim.npy')),
            'u_shp': np.load(os.path.join(d, 'u_shp.npy')),
            'u_exp': np.load(os.path.join(d, 'u_exp.npy')),
        }
        with open(os.path.join(d, 'param_whitening.pkl'), 'rb') as f:
            meta = pickle.load(f)
        out['param_mean'], out['param_std'] = meta.get('param_mean'), meta.get('param_std')
        tri_fp = os.path.join(d, 'tri.mat')
        if os.path.exists(tri_fp):
            import scipy.io as sio
            out['tri'] = sio.loadmat(tri_fp)['tri']
        return out

    # derived quantities the reference also exposes (utils/params.py:26-29)
    @property
    def w(self):
        return np.concatenate((self.w_shp, self.w_exp), axis=1)

    @property
    def w_base(self):
        return self.w[self.keypoints]

    @property
    def w_norm(self):
        return np.linalg.norm(self.w, axis=0)

    @property
    def w_base_norm(self):
        return np.linalg.norm(self.w_base, axis=0)


_pack: Optional[ParamsPack] = None


def get_param_pack() -> ParamsPack:
    """Process-wide pack (the reference builds one at import, model_building.py:8-9)."""
    global _pack
    if _pack is None:
        _pack = ParamsPack()
    return _pack


def set_param_pack(pack: ParamsPack) -> None:
    global _pack
    _pack = pack
