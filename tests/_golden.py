"""Helpers to read the golden fixtures (tests/golden/*.npz, produced by the unmodified reference)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    return {k: z[k] for k in z.files}


def state_dict(g):
    return {k[3:]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith('sd/')}


def sub(g, prefix):
    return {k[len(prefix) + 1:]: v for k, v in g.items() if k.startswith(prefix + '/')}


MO3D_HEADS = {'seg': {'channels': 1, 'activation': 'sigmoid'}, 'flow': {'channels': 2, 'activation': None},
              'dist': {'channels': 1, 'activation': 'relu'}}
