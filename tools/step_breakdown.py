"""Where does a cfg-2 step go outside the network forward? CUDA-event timing of the pipeline stages."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from bio_image_unet_b200 import pipeline2d as P
from bio_image_unet_b200.unet import Session
ses = Session(bench.random_checkpoint(), resize_dim=bench.TILE, add_tile=1, normalization_mode='single', clip_threshold=(0., 99.8),
              device='cuda:0', precision='bf16', workspace_gb=40.0)
frames = torch.from_numpy(bench.synth_frames(8)).cuda()
for _ in range(3):
    ses.predict_device(frames)
torch.cuda.synchronize()
def timed(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out
t_all, _ = timed(lambda: ses.predict_device(frames))
t_norm, norm = timed(lambda: ses.normalise_device(frames))
n_x, n_y, xs, ys = P.tiling.grid_2d(2048, 2048, bench.TILE, 1)
t_gather, tiles = timed(lambda: P.E.gather_tiles(norm.view(8, 1, 2048, 2048), [0], xs, ys, (1, 512, 512), 0))
t_fwd, res = timed(lambda: P.run_tiles(ses.engine, tiles, ses.tile_batch))
t_stitch, _ = timed(lambda: P.E.stitch_mean_u8(res[0], 8, 1, (2048, 2048), xs, ys, (512, 512)))
print(f'step {t_all:.3f} ms = normalise {t_norm:.3f} + gather {t_gather:.3f} + forward {t_fwd:.3f} + stitch {t_stitch:.3f} '
      f'(sum {t_norm + t_gather + t_fwd + t_stitch:.3f})')
