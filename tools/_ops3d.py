import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from bio_image_unet_b200.unet3d import Session, UNet3D
dev = torch.device('cuda', 0)
sd = bench.kaiming_state_dict(UNet3D(n_filter=16), 0)
params = {'state_dict': sd, 'n_filter': 16, 'in_channels': 1, 'out_channels': 1}
vol = np.random.default_rng(0).integers(0, 4000, (64, 1024, 1024)).astype('uint16')
vol_dev = torch.from_numpy(vol).to(dev)
ses = Session(params, (64, 128, 128), device=dev, precision='bf16', workspace_gb=16.0)
for _ in range(3): ses.predict(vol_dev, to_host=False)
torch.cuda.synchronize()
ses.engine.set_profile(True)
acc = None
for _ in range(5):
    ses.predict(vol_dev, to_host=False); torch.cuda.synchronize()
    kinds, ms = ses.engine.read_profile()
    ms = np.array(ms); acc = ms if acc is None else np.minimum(acc, ms)
ses.engine.set_profile(False)
print('batch', ses._planner.tile_batch, 'ops', len(kinds), 'sum %.3f ms' % acc.sum())
print(' '.join(f'{k}:{m:.3f}' for k, m in zip(kinds, acc)))
