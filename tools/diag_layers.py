"""Per-layer error of the engine against the CPU oracle (debug aid, run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import models as om
from tests.test_gpu_unet import stress_state_dict
from bio_image_unet_b200.engine import Engine

def pad16(c): return (c + 15) // 16 * 16

def run(nf, tile, batch, precision, head_gain=4.0):
    sd = stress_state_dict(nf, seed=100 + nf, head_gain=head_gain)
    tiles = torch.randint(0, 256, (batch, 1, *tile), dtype=torch.uint8, generator=torch.Generator().manual_seed(7))
    acts = {}
    with torch.no_grad():
        ref, logits = om.unet_forward(sd, tiles.float() / 255, collect=acts)
    eng = Engine('unet2d', sd, nf, 1, [('', 1, 'sigmoid')], precision=precision, device='cuda:0')
    eng.plan(batch, tile)
    val, _ = eng.forward(tiles.cuda(), want_val=True)
    torch.cuda.synchronize()
    print(f'--- nf={nf} tile={tile} B={batch} {precision}: sigmoid max-abs err {(val.cpu()-ref).abs().max().item():.3e}  logits std {logits.std().item():.2f}')
    ch = [nf * 2 ** i for i in range(5)]
    # (buffer, level, phys channels, [(oracle name, phys offset, count)])
    spec = [('e1', 0, pad16(ch[0]), [('e1', 0, ch[0])]),
            ('cat4', 0, 2 * pad16(ch[0]), [('u4', 0, ch[0]), ('e2', pad16(ch[0]), ch[0])]),
            ('m1', 1, pad16(ch[0]), [('m1', 0, ch[0])]), ('e3', 1, pad16(ch[1]), [('e3', 0, ch[1])]),
            ('cat3', 1, 2 * pad16(ch[1]), [('u3', 0, ch[1]), ('e4', pad16(ch[1]), ch[1])]),
            ('e5', 2, pad16(ch[2]), [('e5', 0, ch[2])]),
            ('cat2', 2, 2 * pad16(ch[2]), [('u2', 0, ch[2]), ('e6', pad16(ch[2]), ch[2])]),
            ('e7', 3, pad16(ch[3]), [('e7', 0, ch[3])]),
            ('cat1', 3, 2 * pad16(ch[3]), [('u1', 0, ch[3]), ('e8', pad16(ch[3]), ch[3])]),
            ('m4', 4, pad16(ch[3]), [('m4', 0, ch[3])]), ('mid1', 4, pad16(ch[4]), [('mid1', 0, ch[4])]),
            ('mid2', 4, pad16(ch[4]), [('mid2', 0, ch[4])]), ('d1', 3, pad16(ch[3]), [('d1', 0, ch[3])]),
            ('d2', 3, pad16(ch[3]), [('d2', 0, ch[3])]), ('d3', 2, pad16(ch[2]), [('d3', 0, ch[2])]),
            ('d4', 2, pad16(ch[2]), [('d4', 0, ch[2])]), ('d5', 1, pad16(ch[1]), [('d5', 0, ch[1])]),
            ('d6', 1, pad16(ch[1]), [('d6', 0, ch[1])]), ('d7', 0, pad16(ch[0]), [('d7', 0, ch[0])])]
    for buf, lvl, cphys, parts in spec:
        a = eng.debug_activation(buf, cphys, lvl)[:, 0]           # (B, h, w, C)
        for oname, off, cnt in parts:
            r = acts[oname].permute(0, 2, 3, 1).numpy()
            g = a[..., off:off + cnt]
            err = np.abs(g - r).max(); scale = np.abs(r).max()
            print(f'  {oname:5s} max|ref|={scale:8.3f} max err={err:9.3e} rel={err / max(scale, 1e-9):9.3e}')
    eng.close()

if __name__ == '__main__':
    for prec in ('tf32', 'bf16'):
        run(32, (64, 64), 2, prec)
    run(4, (32, 32), 1, 'tf32')
    run(32, (64, 64), 2, 'bf16', head_gain=1.0)
