"""CPU restatement (torch fp32) of the Silero-class VAD described in whisperjav_b200/vad.py -- TEST
INFRASTRUCTURE.  Parity unpinned: the real Silero / TEN weights and code (silero-vad 6.2.1, ten-vad 1.0.6.8;
reference call sites speech_segmentation/backends/silero.py:269, silero_v6.py:205, ten.py:237) are not
installable offline, so this oracle and the CUDA kernel are checked against each other on seeded weights."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


def vad_probs(sd: Dict[str, torch.Tensor], audio: np.ndarray) -> np.ndarray:
    x = torch.from_numpy(np.ascontiguousarray(audio)).float()
    n = x.numel()
    nw = (n + 511) // 512
    xp = F.pad(x, (64, nw * 512 - n))  # left context of the first window is silence
    win = torch.hann_window(256, periodic=True, dtype=torch.float64)
    nn_ = torch.arange(256, dtype=torch.float64)
    kk = torch.arange(129, dtype=torch.float64)
    ang = 2 * math.pi * nn_[:, None] * kk[None, :] / 256
    cos_b = (win[:, None] * torch.cos(ang)).float()
    sin_b = (-win[:, None] * torch.sin(ang)).float()
    h = torch.zeros(128)
    c = torch.zeros(128)
    out = np.zeros(nw, dtype=np.float32)
    for w in range(nw):
        u = xp[w * 512: w * 512 + 576]                       # 64 ctx + 512
        # samples past the clip end are zero (already padded); reflect the last 64 of the window
        u = torch.cat([u, torch.flip(u[-65:-1], dims=[0])])   # u[576 + r] = window[510 - r]
        frames = torch.stack([u[128 * s: 128 * s + 256] for s in range(4)], 1)  # [256, 4]
        re, im = cos_b.t() @ frames, sin_b.t() @ frames
        z = torch.sqrt(re * re + im * im)[None]             # [1, 129, 4]
        z = F.relu(F.conv1d(z, sd["conv1.weight"], sd["conv1.bias"], padding=1))
        z = F.relu(F.conv1d(z, sd["conv2.weight"], sd["conv2.bias"], stride=2, padding=1))
        z = F.relu(F.conv1d(z, sd["conv3.weight"], sd["conv3.bias"], stride=2, padding=1))
        z = F.relu(F.conv1d(z, sd["conv4.weight"], sd["conv4.bias"], padding=1))
        xt = z[0, :, 0]
        gates = sd["lstm.weight_ih"] @ xt + sd["lstm.bias_ih"] + sd["lstm.weight_hh"] @ h + sd["lstm.bias_hh"]
        i, f, g, o = gates[:128], gates[128:256], gates[256:384], gates[384:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[w] = torch.sigmoid(sd["out.weight"].reshape(128) @ F.relu(h) + sd["out.bias"][0]).item()
    return out
