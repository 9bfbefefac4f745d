import sys, os, numpy as np, torch, tempfile
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import util
from empose_b200 import lib, synthetic
from oracle import ief as oi, sensors, smplh_lbs
npz = synthetic.write_synthetic_smplh(os.path.join(tempfile.gettempdir(), 'empose_b200_assets'))
osm = smplh_lbs.SmplhModel(npz, dtype=torch.float64); topo = sensors.sensor_topology(osm.faces.numpy())
dev = torch.device('cuda:0')
net = util.build_module(npz, precision=lib.PRECISION_FP32, device=dev)
ctx = net.native_context(dev)
for R in (1, 3, 8, 9, 35):
    p = synthetic.synth_window_params(R, 1, seed=4, offsets=True)
    t = lambda a: torch.from_numpy(np.asarray(a))
    poses = t(p['poses']).reshape(R, 66); shapes = t(p['shapes']); off_r = t(p['offset_r']); off_t = t(p['offset_t'])
    pos, ori, joints = ctx.sensor_project(poses.to(dev), shapes.to(dev), off_r.to(dev), off_t.to(dev))
    with torch.no_grad():
        o_pos, o_ori, o_j = oi.project_sensors(osm, topo, poses.double(), shapes.double(), off_r.double(), off_t.double())
    e = (pos.cpu().double() - o_pos).norm(dim=-1)
    print('R=%d joints err %.2e; pos err per frame (max over sensors):' % (R, (joints.cpu().double()-o_j).abs().max()), np.round(e.max(dim=1).values.numpy(), 4))
    if R == 3: print(np.round(e.numpy(), 4))
