"""Golden vectors of the stages either side of the 3DMM path, recorded from the
UNMODIFIED reference in the build container (``/root/reference`` does not exist on the GPU box).

    python tests/golden/make_golden_render.py        # writes tests/golden/render_vectors.npz

What runs (all reference code, nothing from this repository computes a recorded value):
  * ``Sim3DR``: a scratch copy of ``/root/reference/Sim3DR`` is built with its own ``setup.py build_ext -i`` (Cython +
    ``lib/rasterize_kernel.cpp``); ``Sim3DR.get_normal``, ``Sim3DR.rasterize`` and ``RenderPipeline`` (lighting.py) are
    called through the package exactly as ``utils/render.py:29-45`` does (cfg of ``utils/render.py:18-27``);
  * ``FaceBoxes``: ``PriorBox`` (utils/prior_box.py), ``decode`` (utils/box_utils.py), ``py_cpu_nms``
    (utils/nms/py_cpu_nms.py -- the Cython ``cpu_nms`` does not build with Cython 3 / numpy 2),
    and the post-processing body of ``FaceBoxes.__call__`` (FaceBoxes.py:98-143) executed on the reference's own
    network (``FaceBoxesNet`` + the shipped ``weights/FaceBoxesProd.pth``) for a synthetic image.
Inputs come from ``synergynet_b200.synthetic`` (seeded) and are stored next to the outputs.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference'

from synergynet_b200 import synthetic  # noqa: E402


def scratch_reference() -> str:
    tmp = tempfile.mkdtemp(prefix='synergy_ref_render_')
    dst = os.path.join(tmp, 'ref')
    shutil.copytree(REF, dst, ignore=shutil.ignore_patterns('*.ipynb', 'img', 'demo', '.git'))
    os.system(f'chmod -R u+w {dst}')
    subprocess.run([sys.executable, 'setup.py', 'build_ext', '-i'], cwd=os.path.join(dst, 'Sim3DR'), check=True, capture_output=True)
    with open(os.path.join(dst, 'FaceBoxes/utils/nms/cpu_nms.py'), 'w') as f:
        f.write('from .py_cpu_nms import py_cpu_nms as cpu_nms\n'
                'def cpu_soft_nms(*a, **k):\n    raise NotImplementedError\n')
    sys.path.insert(0, dst)
    os.chdir(dst)
    return dst


RENDER_CFG = {   # utils/render.py:18-27 (importing utils/render.py itself needs scipy + a 3dmm_data directory)
    'intensity_ambient': 0.75, 'color_ambient': (1, 1, 1), 'intensity_directional': 0.7, 'color_directional': (1, 1, 1),
    'intensity_specular': 0.2, 'specular_exp': 5, 'light_pos': (0, 0, 5), 'view_pos': (0, 0, 5)}


def main():
    scratch_reference()
    import Sim3DR as ref_s3d                              # reference package: get_normal, rasterize, RenderPipeline
    from FaceBoxes.utils.prior_box import PriorBox
    from FaceBoxes.utils.box_utils import decode
    from FaceBoxes.utils.nms.py_cpu_nms import py_cpu_nms
    from FaceBoxes.utils.config import cfg as fb_cfg
    import importlib
    fb_mod = importlib.import_module('FaceBoxes.FaceBoxes')     # the module (the package re-exports the class under the same name)
    out = {}

    # ---- f2: a small batch of meshes on a 96 x 128 canvas ---------------------------------------------------------------
    rows, cols, h, w = 40, 50, 96, 128
    tri = synthetic.make_render_topology(rows, cols)
    verts = synthetic.make_render_meshes(3, h, w, seed=2, rows=rows, cols=cols, size=60)        # (B,3,N)
    bg = (np.arange(h * w * 3, dtype=np.int64).reshape(h, w, 3) * 7 % 251).astype(np.uint8)
    app = ref_s3d.RenderPipeline(**RENDER_CFG)
    overlap = bg.copy()
    normals, lights, steps = [], [], []
    for b in range(verts.shape[0]):
        ver = np.ascontiguousarray(verts[b].astype(np.float32).T)                             # utils/render.py:42-44
        normals.append(ref_s3d.get_normal(ver, tri))
        # the light the pipeline rasterises with: recovered by rendering a texture of ones (`texture *= light`, lighting.py:73)
        tex = np.ones_like(ver)
        app(ver, tri, bg.copy(), texture=tex)
        lights.append(tex.copy())
        overlap = app(ver, tri, overlap)
        steps.append(overlap.copy())
    out.update(render_tri=tri, render_verts=verts, render_bg=bg, render_normals=np.stack(normals), render_light=np.stack(lights),
               render_steps=np.stack(steps))
    # plain rasterize with given colours, both orientations
    rng = np.random.default_rng(11)
    colors = rng.uniform(0, 1, (rows * cols, 3)).astype(np.float32)
    ver0 = np.ascontiguousarray(verts[0].T)
    out['raster_colors'] = colors
    out['raster_plain'] = ref_s3d.rasterize(ver0, tri, colors, bg=bg.copy())
    out['raster_reverse'] = ref_s3d.rasterize(ver0, tri, colors, bg=bg.copy(), reverse=True)

    # ---- f3: prior boxes, decode, NMS ------------------------------------------------------------------------------------------
    for (ih, iw) in ((96, 160), (250, 333)):
        out[f'priors_{ih}x{iw}'] = PriorBox(image_size=(ih, iw)).forward().numpy()
    g = torch.Generator().manual_seed(5)
    pri = torch.from_numpy(out['priors_250x333'])
    loc = torch.randn((pri.shape[0], 4), generator=g) * 1.5
    out['decode_loc'] = loc.numpy()
    out['decode_boxes'] = decode(loc, pri, fb_cfg['variance']).numpy()
    # NMS on clustered boxes with distinct scores
    centers = rng.uniform(40, 400, (60, 2))
    cidx = rng.integers(0, 60, 1500)
    wh = rng.uniform(20, 90, (1500, 2))
    c = centers[cidx] + rng.normal(0, 6, (1500, 2))
    scores = rng.permutation(1500).astype(np.float32) / 1500 * 0.95 + 0.05
    dets = np.hstack([c - wh / 2, c + wh / 2, scores[:, None]]).astype(np.float32)
    out['nms_dets'] = dets
    for thr in (0.3, 0.5):
        out[f'nms_keep_{int(thr * 10)}'] = np.array(py_cpu_nms(dets, thr), np.int64)

    # the detector end to end on a synthetic image: reference network + weights, then FaceBoxes.py:98-143 verbatim
    torch.manual_seed(0)
    net = fb_mod.FaceBoxes()
    ih, iw = 240, 320
    yy, xx = np.mgrid[0:ih, 0:iw]
    img = np.zeros((ih, iw, 3), np.float32)
    for (cy, cx, r) in ((80, 90, 38), (150, 230, 46)):                                       # two face-ish blobs
        d2 = ((yy - cy) / (1.25 * r)) ** 2 + ((xx - cx) / r) ** 2
        img += np.where(d2 < 1, 1.0, 0.0)[..., None] * np.array([120, 150, 200], np.float32)
        for (ey, ex) in ((-0.3, -0.35), (-0.3, 0.35)):
            img -= np.where(((yy - cy - ey * r) ** 2 + (xx - cx - ex * r) ** 2) < (0.12 * r) ** 2, 1.0, 0.0)[..., None] * 110
        img -= np.where((((yy - cy - 0.45 * r) / 0.1) ** 2 + ((xx - cx) / 0.4) ** 2) < r * r, 1.0, 0.0)[..., None] * 70
    img += rng.normal(0, 6, img.shape)
    img_u8 = np.clip(img + 40, 0, 255).astype(np.uint8)
    x = np.float32(img_u8) - (104, 117, 123)
    x = torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1), dtype=np.float32)).unsqueeze(0)
    with torch.no_grad():
        loc_n, conf_n = net.net(x)
    out.update(fb_image=img_u8, fb_loc=loc_n.squeeze(0).numpy(), fb_conf=conf_n.squeeze(0).numpy())
    # random-score variant (the synthetic blobs rarely cross vis_thres): FaceBoxes.py:98-143 on recorded network outputs
    conf_r = torch.softmax(torch.randn((loc_n.shape[1], 2), generator=g) * 2.0, dim=-1)
    for tag, conf_t in (('net', conf_n.squeeze(0)), ('rnd', conf_r)):
        priors = PriorBox(image_size=(ih, iw)).forward()
        boxes = decode(loc_n.squeeze(0), priors, fb_cfg['variance'])
        boxes = boxes * torch.Tensor([iw, ih, iw, ih]) / 1 / fb_mod.resize
        boxes = boxes.cpu().numpy()
        sc = conf_t.numpy()[:, 1]
        inds = np.where(sc > fb_mod.confidence_threshold)[0]
        boxes, sc = boxes[inds], sc[inds]
        order = sc.argsort()[::-1][:fb_mod.top_k]
        boxes, sc = boxes[order], sc[order]
        d = np.hstack((boxes, sc[:, np.newaxis])).astype(np.float32, copy=False)
        keep = py_cpu_nms(d, fb_mod.nms_threshold) if d.shape[0] else []
        kept = d[keep, :][:fb_mod.keep_top_k, :]
        out[f'fb_{tag}_conf'] = conf_t.numpy()
        out[f'fb_{tag}_dets_sorted'] = d
        out[f'fb_{tag}_keep'] = np.array(keep, np.int64)
        out[f'fb_{tag}_final'] = np.array([b for b in kept if b[4] > fb_mod.vis_thres], np.float32).reshape(-1, 5)
    # the detector with the seeded synthetic checkpoint the GPU tests can rebuild (the shipped weights cannot travel):
    # reference network class, strict load, its forward; then the reference's own FaceBoxes.__call__ with that network
    sd = synthetic.make_faceboxes_state_dict(0)
    ref_keys = list(net.net.state_dict().keys())
    assert ref_keys == __import__('synergynet_b200.faceboxes', fromlist=['x']).state_dict_keys(), 'key schema differs from the reference'
    net.net.load_state_dict(sd, strict=True)
    net.net.eval()
    for (sh, sw, seed) in ((250, 333, 0), (120, 96, 1)):
        scene = synthetic.make_scene_u8(sh, sw, seed)
        xs = torch.from_numpy(np.ascontiguousarray((np.float32(scene) - (104, 117, 123)).transpose(2, 0, 1), dtype=np.float32)).unsqueeze(0)
        with torch.no_grad():
            l_s, c_s = net.net(xs)
        out[f'fbs_loc_{sh}x{sw}'] = l_s.squeeze(0).numpy()
        out[f'fbs_conf_{sh}x{sw}'] = c_s.squeeze(0).numpy()
        out[f'fbs_final_{sh}x{sw}'] = np.array(net(scene), np.float32).reshape(-1, 5)            # FaceBoxes.__call__, unmodified
    path = os.path.join(ROOT, 'tests', 'golden', 'render_vectors.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
