"""Silero-class voice-activity model on the GPU (csrc/vad.cu) and its weight layout.

Architecture (restated from the published Silero VAD v5 description; the JIT/ONNX weights the reference
downloads at whisperjav/modules/speech_segmentation/backends/silero.py:199-206 / silero_v6.py:143 cannot be
fetched offline, so weights are seeded synthetic and parity is against oracle/vad_oracle.py -- unpinned):

    window = 512 samples @16 kHz with 64 samples of left context and 64 reflected samples on the right
    STFT: 256-tap Hann DFT, hop 128 -> 4 steps x 129 magnitudes
    Conv1d(129,128,3,p1)+ReLU -> Conv1d(128,64,3,s2,p1)+ReLU -> Conv1d(64,64,3,s2,p1)+ReLU -> Conv1d(64,128,3,p1)+ReLU
    LSTMCell(128,128) carried across windows -> ReLU -> Conv1d(128,1,1) -> sigmoid

``load_state_dict`` accepts torch-shaped tensors (conv ``[out, in, 3]``, LSTM ``weight_ih [512,128]`` ...) so a
real checkpoint of this shape can be dropped in.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from . import _lib

WINDOW = 512


def synth_vad_weights(seed: int = 5) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std):
        return torch.randn(*shape, generator=g) * std

    w = {}
    for name, (co, ci) in {"conv1": (128, 129), "conv2": (64, 128), "conv3": (64, 64), "conv4": (128, 64)}.items():
        w[name + ".weight"] = rn(co, ci, 3, std=1.6 / math.sqrt(3 * ci))
        w[name + ".bias"] = rn(co, std=0.1)
    w["lstm.weight_ih"] = rn(512, 128, std=1.0 / math.sqrt(128))
    w["lstm.weight_hh"] = rn(512, 128, std=1.0 / math.sqrt(128))
    w["lstm.bias_ih"] = rn(512, std=0.1)
    w["lstm.bias_hh"] = rn(512, std=0.1)
    w["out.weight"] = rn(1, 128, 1, std=4.0 / math.sqrt(128))
    w["out.bias"] = torch.tensor([-0.5])
    return w


def energy_vad_weights(seed: int = 5, noise: float = 1e-3, band=(3, 64), level: float = 0.13, gain: float = 20.0) -> Dict[str, torch.Tensor]:
    """Hand-built weights that make the stack a smoothed band-energy detector (mean 94 Hz-2 kHz STFT
    magnitude vs ``level``), perturbed by seeded noise so no term of the network is degenerate.  Gives
    speech-like segments on the synthetic audio, which random weights do not."""
    g = torch.Generator().manual_seed(seed)

    def nz(*shape):
        return torch.randn(*shape, generator=g) * noise

    w = {}
    c1 = nz(128, 129, 3)
    c1[:, band[0]:band[1], 1] += 1.0 / (band[1] - band[0])
    w["conv1.weight"], w["conv1.bias"] = c1, nz(128).abs()
    for name, (co, ci) in {"conv2": (64, 128), "conv3": (64, 64)}.items():
        c = nz(co, ci, 3)
        c[:, :, 1] += 0.5 / ci
        c[:, :, 2] += 0.5 / ci
        w[name + ".weight"], w[name + ".bias"] = c, nz(co).abs()
    c4 = nz(128, 64, 3)
    c4[:, :, 1] += 1.0 / 64
    w["conv4.weight"], w["conv4.bias"] = c4, nz(128).abs()
    wih = nz(512, 128)
    wih[256:384] += gain / 128
    bih = nz(512)
    bih[0:128] += 4.0
    bih[128:256] += 1.5
    bih[256:384] += -gain * level
    bih[384:512] += 4.0
    w["lstm.weight_ih"], w["lstm.bias_ih"] = wih, bih
    w["lstm.weight_hh"], w["lstm.bias_hh"] = nz(512, 128) * 10, nz(512)
    w["out.weight"] = (nz(1, 128, 1) + 8.0 / 128)
    w["out.bias"] = torch.tensor([-2.0])
    return w


def dft_basis() -> torch.Tensor:
    """[256 taps][258] = Hann(256, periodic) * (cos | -sin)(2 pi k n / 256), k = 0..128."""
    n = torch.arange(256, dtype=torch.float64)
    win = 0.5 - 0.5 * torch.cos(2 * math.pi * n / 256)
    k = torch.arange(129, dtype=torch.float64)
    ang = 2 * math.pi * n[:, None] * k[None, :] / 256
    return torch.cat([win[:, None] * torch.cos(ang), -win[:, None] * torch.sin(ang)], dim=1).float()


def pack_vad_weights(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """fp32 blob in the layout of csrc/vad.cu::VadOff (every matrix K-major-transposed: [K][N])."""
    def conv_t(w):  # [co, ci, 3] -> [tap*ci + c][co]
        co, ci, _ = w.shape
        return w.permute(2, 1, 0).reshape(3 * ci, co)

    parts = [dft_basis(),
             conv_t(sd["conv1.weight"]), sd["conv1.bias"], conv_t(sd["conv2.weight"]), sd["conv2.bias"],
             conv_t(sd["conv3.weight"]), sd["conv3.bias"], conv_t(sd["conv4.weight"]), sd["conv4.bias"],
             sd["lstm.weight_ih"].t(), sd["lstm.bias_ih"] + sd["lstm.bias_hh"], sd["lstm.weight_hh"].t(),
             sd["out.weight"].reshape(128), sd["out.bias"].reshape(1), torch.zeros(3)]
    blob = torch.cat([p.contiguous().reshape(-1).float() for p in parts])
    assert blob.numel() * 4 == _lib.load().wjb_vad_weights_bytes(), (blob.numel() * 4, _lib.load().wjb_vad_weights_bytes())
    return blob


class VadB200:
    """Per-window speech probabilities for a batch of clips in one device pass."""

    def __init__(self, state_dict: Dict[str, torch.Tensor] = None, device="cuda", seed: int = 5):
        if not torch.cuda.is_available():
            raise _lib.WjbError("VadB200 needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.state_dict = state_dict or energy_vad_weights(seed)
        self._blob = pack_vad_weights(self.state_dict).to(self.device)

    def probs(self, audio: torch.Tensor, n_samples: torch.Tensor) -> torch.Tensor:
        """audio fp32 [B, S] (device), n_samples int32 [B] (device) -> fp32 [B, ceil(S/512)] probabilities
        (windows past a clip's end are 0)."""
        B, S = audio.shape
        nw = (S + WINDOW - 1) // WINDOW
        out = torch.empty(B, nw, dtype=torch.float32, device=self.device)
        ws = torch.empty(self.lib.wjb_vad_workspace_bytes(B, nw), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.wjb_vad_forward(_lib.ptr(audio), audio.stride(0), _lib.ptr(n_samples), B, _lib.ptr(self._blob),
                                               _lib.ptr(out), nw, _lib.ptr(ws), _lib.stream_ptr()), "wjb_vad_forward")
        return out
