"""One invocation of every kernel of the Sim3DR / FaceBoxes stages for an ncu capture (bench workload sizes): 8 meshes on a
720 x 1080 canvas, the detector network + decode + NMS on one 720 x 1080 image.  No warm-up loop: ncu replays each launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synergynet_b200 import Sim3DR, detect, faceboxes, synthetic  # noqa: E402
from synergynet_b200.inference import RENDER_CFG  # noqa: E402

dev = torch.device('cuda', 0)
tri = synthetic.make_render_topology()
verts = torch.from_numpy(synthetic.make_render_meshes(8, 720, 1080, seed=0)).to(dev)
r = Sim3DR.MeshRenderer(tri, verts.shape[2], dev)
canvas = torch.zeros((720, 1080, 3), dtype=torch.uint8, device=dev)
r.render(canvas, verts.transpose(1, 2), Sim3DR._light_cfg(**RENDER_CFG))
net = faceboxes.FaceBoxesNet(synthetic.make_faceboxes_state_dict(0), dev)
loc, conf = net.forward(torch.from_numpy(synthetic.make_scene_u8(720, 1080, 0)).to(dev))
dets, n = detect.decode_device(loc, conf, 720, 1080)
detect.nms_device(dets, 0.3, n=int(n.item()))
torch.cuda.synchronize()
print('done', int(n.item()))
