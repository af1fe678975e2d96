import sys, os, numpy as np, torch, argparse
sys.path.insert(0, '.')
import bench
from bio_image_unet_b200.dist import DistContext
dev = torch.device('cuda', 0)
c = int(sys.argv[1])
args = argparse.Namespace(config=c, frames=None, chunk_frames=8, add_patch=0, precision=None, vol='64,256,256' if c == 4 else None)
wl = bench.WORKLOADS[c](args, dev, DistContext(False)).setup()
wl.step_device(); torch.cuda.synchronize()
