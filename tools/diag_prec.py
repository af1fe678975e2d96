"""Sigmoid / logit error of each precision mode against the CPU oracle for several weight regimes (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import models as om
from tests import _golden
from tests.test_gpu_unet import stress_state_dict
from bio_image_unet_b200.engine import Engine
from bio_image_unet_b200.unet import Unet

def report(tag, sd, nf, tiles):
    with torch.no_grad():
        ref, logits = om.unet_forward(sd, tiles.float() / 255)
    out = [f'{tag:34s} logit std {logits.std().item():6.3f}']
    for prec in ('fp32', 'tf32', 'bf16'):
        eng = Engine('unet2d', sd, nf, 1, [('', 1, None)], precision=prec, device='cuda:0')
        eng.plan(tiles.shape[0], tuple(tiles.shape[2:]))
        val, _ = eng.forward(tiles.cuda(), want_val=True)
        lg = val.cpu()
        sig_err = (torch.sigmoid(lg) - ref).abs().max().item()
        out.append(f'{prec}: logit err {(lg - logits).abs().max().item():.2e} sig err {sig_err:.2e}')
        eng.close()
    print(' | '.join(out))

g = torch.Generator().manual_seed(3)
tiles64 = torch.randint(0, 256, (2, 1, 64, 64), dtype=torch.uint8, generator=g)
tiles256 = torch.randint(0, 256, (1, 1, 256, 256), dtype=torch.uint8, generator=g)
for name in ('unet_single', 'unet_all_invert'):
    gd = _golden.load(name)
    report('golden ' + name, _golden.state_dict(gd), 4, torch.from_numpy(gd['patches'][:4].copy()))
torch.manual_seed(0)
report('default init nf=32 64x64', Unet(n_filter=32).state_dict(), 32, tiles64)
report('default init nf=32 256x256', Unet(n_filter=32).state_dict(), 32, tiles256)
for hg in (4.0, 1.0):
    report(f'stress nf=32 head_gain={hg}', stress_state_dict(32, 132, hg), 32, tiles64)
# stress with logits normalised to std 1
sd = stress_state_dict(32, 132, 1.0)
with torch.no_grad():
    _, lg = om.unet_forward(sd, tiles64.float() / 255)
sd['final.0.weight'] = sd['final.0.weight'] / lg.std()
sd['final.0.bias'] = sd['final.0.bias'] - (lg.mean() / lg.std())
report('stress nf=32 logits ~ N(0,1)', sd, 32, tiles64)
report('stress nf=32 logits ~ N(0,1) 256', sd, 32, tiles256)
