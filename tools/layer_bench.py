"""Time one conv layer through the C-ABI (debug / profiling aid). Usage: layer_bench.py cin cout H W B [reps] [esz]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bio_image_unet_b200 import _lib

cin, cout, H, W, B = (int(v) for v in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
esz = int(sys.argv[7]) if len(sys.argv) > 7 else 2
lib = _lib.load()
dt = torch.bfloat16 if esz == 2 else torch.float32
x = torch.randn(B, H, W, cin, device='cuda').to(dt)
w = (torch.randn(9, cout, cin, device='cuda') / (9 * cin) ** 0.5).to(dt)
scale = torch.ones(cout, device='cuda'); shift = torch.zeros(cout, device='cuda')
out = torch.empty(B, H, W, cout, device='cuda', dtype=dt)
def run():
    _lib.check(lib.biu_conv_tc(esz, _lib.ptr(x), cin, 0, cin, B, 1, H, W, 1, 3, 3, _lib.ptr(w), cout, _lib.ptr(scale),
                               _lib.ptr(shift), 0.1, _lib.ptr(out), cout, 0, _lib.stream_ptr()))
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2.0 * B * H * W * cout * 9 * cin
print(f'conv {cin}->{cout} @{H}x{W} B={B} esz={esz} v1={os.environ.get("BIU_CONV_V1","0")}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  '
      f'in+out {(x.numel() + out.numel()) * esz / ms / 1e6:.0f} GB/s')
