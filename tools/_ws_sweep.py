import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from bio_image_unet_b200.unet3d import Session, UNet3D
dev = torch.device('cuda', 0)
sd = bench.kaiming_state_dict(UNet3D(n_filter=16), 0)
params = {'state_dict': sd, 'n_filter': 16, 'in_channels': 1, 'out_channels': 1}
vol = np.random.default_rng(0).integers(0, 4000, (256, 1024, 1024)).astype('uint16')
vol_dev = torch.from_numpy(vol).to(dev)
for ws in [2, 4, 8, 16, 30, 60]:
    ses = Session(params, (64, 128, 128), device=dev, precision='bf16', workspace_gb=float(ws))
    for _ in range(2): ses.predict(vol_dev, to_host=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ses.predict(vol_dev, to_host=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f'ws {ws} GB: batch {ses._planner.tile_batch}  {ms:.2f} ms  {vol.size / ms / 1e6:.2f} Gvox/s', flush=True)
    ses.close(); del ses; torch.cuda.empty_cache()
