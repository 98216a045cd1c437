"""Deterministic synthetic weights and audio (there are no checkpoints and no network on
either box).

``synth_weights`` fills the exact Whisper architecture (names follow the openai-whisper
``state_dict``) with seeded values chosen so that greedy decoding is *non-degenerate*:
attention is peaky enough that logits depend on the audio window, the tied embedding is
spread so the argmax changes step to step, and the EOT row is boosted so sequences end
at varied lengths.  ``speech_shaped_audio`` is the
"Japanese-speech-shaped" generator of SURVEY.md section 8d.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch


@dataclass(frozen=True)
class Dims:
    """Whisper ``ModelDimensions`` (upstream whisper/model.py)."""
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


DIMS = {
    "tiny": Dims(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4),
    "base": Dims(80, 1500, 512, 8, 6, 51865, 448, 512, 8, 6),
    "small": Dims(80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12),
    "medium": Dims(80, 1500, 1024, 16, 24, 51865, 448, 1024, 16, 24),
    "large-v1": Dims(80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
    "large-v2": Dims(80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
    "large-v3": Dims(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
    "large": Dims(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),  # upstream alias of large-v3
    "large-v3-turbo": Dims(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
    "turbo": Dims(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
}

EOT = 50257


def _sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1)


# Per-architecture generator settings used by load_model(), bench.py and the tests: seeds / EOT boosts picked (with
# scripts/synth_stats.py, CPU oracle) so that greedy sequences end at varied lengths in both timestamp modes.
SYNTH_PRESETS = {
    "tiny": {"seed": 5},
    "large-v3": {"seed": 11, "eot_boost": 4.0, "attn_logit_std": 6.0, "cross_attn_logit_std": 14.0, "cross_gain": 0.3},
}


def synth_preset(name: str) -> dict:
    return dict(SYNTH_PRESETS.get(name, {"seed": 11}))


def synth_weights(dims: Dims, seed: int = 11, dtype=torch.float16, eot_boost: float = 1.8,
                  logit_std: float = 7.0, attn_logit_std: float = 8.0, res_gain: float = None,
                  cross_gain: float = 0.4, bias_std: float = 0.02, emb_norm_sigma: float = 0.0,
                  eot_drift: float = 0.0, eot_t0: float = 100.0,
                  enc_attn_logit_std: float = 2.0, cross_attn_logit_std: float = 8.0,
                  emb_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded weights in openai-whisper ``state_dict`` naming.  Stored in ``dtype`` (fp16:
    what the reference's ``fp16=True`` run holds after ``model.half()``).

    The defaults are tuned on the CPU oracle (scripts/synth_stats.py; asserted in tests/test_oracle_synth.py) for what a
    parity test needs from a random-init model:
    * decoder self- and cross-attention are sharp (score std 8) so the hidden state, and with it the arg-max, moves from step
      to step and depends on which audio frames the step attends to (a soft cross-attention adds the same vector every step:
      repeats, no EOT); encoder self-attention is soft (2) -- sharp softmaxes amplify fp16 rounding differences (encoder
      fp16-vs-fp32 relative difference 7.6e-3 at std 10, 7e-4 at std 2) without adding anything the test needs;
    * embedding rows are small against the block outputs (a tied embedding's own logit grows with |row|^2 / rms(residual):
      large rows make the model repeat its input token) and the final LayerNorm gain sets the logit spread to
      ``logit_std`` = 7: top logits around 30, avg_logprob around -0.6 (above the reference's logprob_threshold of -1.0, so the
      temperature ladder is not permanently triggered), oracle top-2 margins with median about 1.2;
    * 50 k exchangeable Gaussian logits have a top-2 gap that is exponentially distributed with mean sigma / sqrt(2 ln V):
      about 6 % of steps have a margin below 0.1 whatever the seed.  A trained model is more decisive than that; a random
      one cannot be, so the parity tests judge near-ties explicitly (oracle/parity.py) instead of assuming there are none."""
    g = torch.Generator().manual_seed(seed)
    if res_gain is None:
        res_gain = math.sqrt(32.0 / dims.n_text_layer)

    def rn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    w: Dict[str, torch.Tensor] = {}

    def ln(prefix, n):
        w[prefix + ".weight"] = (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype)
        w[prefix + ".bias"] = rn(n, std=bias_std)

    def attn(prefix, n, n_head, gain=1.0, logit_std_=None):
        d = n // n_head
        # q.k/sqrt(d) with unit-variance inputs has std  sq*sk*n*sqrt(d)/sqrt(d) = sq*sk*n
        s_qk = math.sqrt((logit_std_ or attn_logit_std) / n)
        w[prefix + ".query.weight"] = rn(n, n, std=s_qk)
        w[prefix + ".query.bias"] = rn(n, std=bias_std)
        w[prefix + ".key.weight"] = rn(n, n, std=s_qk)
        w[prefix + ".value.weight"] = rn(n, n, std=1.0 / math.sqrt(n))
        w[prefix + ".value.bias"] = rn(n, std=bias_std)
        w[prefix + ".out.weight"] = rn(n, n, std=gain * res_gain / math.sqrt(n))
        w[prefix + ".out.bias"] = rn(n, std=0.5 * bias_std)

    def mlp(prefix, n):
        w[prefix + ".0.weight"] = rn(4 * n, n, std=1.0 / math.sqrt(n))
        w[prefix + ".0.bias"] = rn(4 * n, std=bias_std)
        w[prefix + ".2.weight"] = rn(n, 4 * n, std=0.8 * res_gain / math.sqrt(4 * n))
        w[prefix + ".2.bias"] = rn(n, std=0.5 * bias_std)

    n = dims.n_audio_state
    w["encoder.conv1.weight"] = rn(n, dims.n_mels, 3, std=1.0 / math.sqrt(dims.n_mels * 3))
    w["encoder.conv1.bias"] = rn(n, std=0.1)
    w["encoder.conv2.weight"] = rn(n, n, 3, std=1.5 / math.sqrt(n * 3))
    w["encoder.conv2.bias"] = rn(n, std=0.1)
    w["encoder.positional_embedding"] = _sinusoids(dims.n_audio_ctx, n)  # fp32 buffer upstream
    for i in range(dims.n_audio_layer):
        p = f"encoder.blocks.{i}"
        ln(p + ".attn_ln", n)
        attn(p + ".attn", n, dims.n_audio_head, logit_std_=enc_attn_logit_std)
        ln(p + ".mlp_ln", n)
        mlp(p + ".mlp", n)
    ln("encoder.ln_post", n)

    n = dims.n_text_state
    # typical |embedding row| = emb_scale; the final LayerNorm gain brings the logit spread to logit_std (tied embeddings: a
    # row's own logit grows with |row|^2 / rms(residual), so rows are kept small against the block outputs and the scale of
    # the logits is set on the way out instead)
    e_scale = emb_scale if emb_scale else logit_std
    emb = torch.randn(dims.n_vocab, n, generator=g) * (e_scale / math.sqrt(n))
    if emb_norm_sigma > 0:
        # Zipf-like vocabulary: row norms are log-normal, so a few hundred "frequent" ids carry most of the arg-max mass and the
        # top-2 margin of a step is heavier-tailed than that of 50 k exchangeable Gaussians (scripts/synth_stats.py)
        norms = torch.exp(emb_norm_sigma * torch.randn(dims.n_vocab, generator=g))
        norms = norms / norms.pow(2).mean().sqrt()
        norms[EOT] = 1.0
        emb = emb * norms[:, None]
    emb[EOT] *= eot_boost
    w["decoder.token_embedding.weight"] = emb.to(dtype)
    pos = torch.randn(dims.n_text_ctx, n, generator=g) * (0.5 * e_scale / math.sqrt(n))
    if eot_drift != 0:
        # the EOT logit drifts upwards with the position, so that sequences end at audio-dependent but bounded lengths
        e_hat = emb[EOT] / emb[EOT].norm()
        t = torch.arange(dims.n_text_ctx, dtype=torch.float32)
        pos = pos + (eot_drift * e_scale * (t - eot_t0) / (dims.n_text_ctx / 2))[:, None] * e_hat[None, :]
    w["decoder.positional_embedding"] = pos.to(dtype)
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}"
        ln(p + ".attn_ln", n)
        attn(p + ".attn", n, dims.n_text_head)
        ln(p + ".cross_attn_ln", n)
        attn(p + ".cross_attn", n, dims.n_text_head, gain=cross_gain, logit_std_=cross_attn_logit_std)
        ln(p + ".mlp_ln", n)
        mlp(p + ".mlp", n)
    ln("decoder.ln", n)
    if emb_scale:
        w["decoder.ln.weight"] = (w["decoder.ln.weight"].float() * (logit_std / e_scale)).to(dtype)
        w["decoder.ln.bias"] = (w["decoder.ln.bias"].float() * (logit_std / e_scale)).to(dtype)
    return w


# ------------------------------------------------------------------------------- audio
def _resonator(x: np.ndarray, f: float, bw: float, sr: int) -> np.ndarray:
    """Two-pole resonator (formant) applied with scipy.signal.lfilter."""
    from scipy.signal import lfilter
    r = math.exp(-math.pi * bw / sr)
    theta = 2 * math.pi * f / sr
    a = [1.0, -2 * r * math.cos(theta), r * r]
    b = [1.0 - r]
    return lfilter(b, a, x)


def speech_shaped_audio(seconds: float, seed: int, sr: int = 16000, duty: float = 0.55) -> np.ndarray:
    """"Japanese-speech-shaped" mono fp32 audio in [-1, 1]: harmonic
    source on f0 in U(110, 320) Hz with slow contour, three formant resonators, mora-rate AM
    6.5-8.5 Hz, 20 % unvoiced fricative morae, utterances 0.4-4 s separated by 0.15-2.5 s
    pauses, pink-ish background noise at -35..-20 dBFS, peak -3 dBFS, int16-quantised."""
    from scipy.signal import lfilter
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    out = np.zeros(n, dtype=np.float64)
    t = 0
    # background: white -> one-pole low-pass ~ pink-ish
    bg = lfilter([0.05], [1.0, -0.95], rng.standard_normal(n))
    bg *= 10 ** (rng.uniform(-35, -20) / 20) / (np.abs(bg).max() + 1e-9)
    pause_lo, pause_hi = 0.15, 2.5
    # scale pause range so that the speech duty is roughly ``duty``
    while t < n:
        pause = math.exp(rng.uniform(math.log(pause_lo), math.log(pause_hi))) * (1 - duty) / 0.45
        t += int(pause * sr)
        if t >= n:
            break
        dur = rng.uniform(0.4, 4.0)
        m = min(int(dur * sr), n - t)
        if m < 160:
            break
        tt = np.arange(m) / sr
        f0 = rng.uniform(110, 320)
        contour = f0 * (1 + 0.15 * np.sin(2 * math.pi * rng.uniform(0.3, 1.2) * tt + rng.uniform(0, 6.28)))
        phase = 2 * math.pi * np.cumsum(contour) / sr
        src = np.zeros(m)
        for h in range(1, 25):
            if h * f0 * 1.15 > sr / 2:
                break
            src += np.sin(h * phase) / h
        mora = rng.uniform(6.5, 8.5)
        am = 0.55 + 0.45 * np.sin(2 * math.pi * mora * tt + rng.uniform(0, 6.28))
        # unvoiced morae: replace 20 % of mora periods with high-passed noise
        mora_idx = (tt * mora).astype(int)
        unv = rng.random(mora_idx.max() + 1) < 0.2
        noise = rng.standard_normal(m)
        noise = noise - lfilter([0.3], [1.0, -0.7], noise)  # crude high-pass
        sig = np.where(unv[mora_idx], 0.5 * noise, src)
        v = np.zeros(m)
        for (lo, hi), bw in zip(((300, 900), (900, 2500), (2500, 3500)), (90, 120, 160)):
            v += _resonator(sig, rng.uniform(lo, hi), bw, sr)
        env = np.minimum(1.0, np.minimum(tt, tt[::-1]) / 0.03)
        out[t:t + m] += v * am * env
        t += m
    out /= (np.abs(out).max() + 1e-9)
    out = out * 10 ** (-3 / 20) * 0.9 + bg
    out = np.clip(out, -1.0, 1.0)
    q = np.round(out * 32767.0).astype(np.int16)  # the WAV hand-off is PCM16
    return (q.astype(np.float32) / 32768.0).astype(np.float32)


def speech_shaped_stream(seconds: float, seed: int, n_base: int = 40, clip_s: float = 30.0, sr: int = 16000) -> np.ndarray:
    """A long "stream" (BASELINE configs 3 / 4: 2 h) assembled from ``n_base`` distinct speech-shaped clips drawn in a seeded
    order (generating two hours sample by sample would take minutes of host time per stream and add nothing)."""
    rng = np.random.default_rng(seed)
    base = [speech_shaped_audio(clip_s, 100000 + 131 * seed + i, sr) for i in range(n_base)]
    n = int(round(seconds * sr))
    out = np.empty(n, dtype=np.float32)
    t = 0
    while t < n:
        c = base[int(rng.integers(n_base))]
        m = min(len(c), n - t)
        out[t: t + m] = c[:m]
        t += m
    return out


def film_audio(seconds: float, seed: int, n_base: int = 40, clip_s: float = 30.0, sr: int = 16000) -> np.ndarray:
    """A film-shaped stream for the scene detector (scene_detection_backends/auditok_backend.py: chapters separated by long
    silences, oversized chapters holding shorter pauses): chapters of 2-80 s cut from seeded speech-shaped clips, separated by
    0.4-4 s of room tone at about -75 dBFS (below the pass-1 gate of 32 dB re 1 LSB), pauses of 1.0-1.6 s at -68 dBFS inside the
    long chapters (what pass 2 splits on), and now and then a 35-50 s "murmur" chapter at about 35 dB that passes the pass-1 gate
    but not the pass-2 gate (the brute-force branch).  int16-quantised like every WAV hand-off of the reference."""
    rng = np.random.default_rng(seed)
    base = [speech_shaped_audio(clip_s, 200000 + 131 * seed + i, sr) for i in range(n_base)]
    n = int(round(seconds * sr))
    out = np.empty(n, dtype=np.float32)

    def tone(m, dbfs):
        return (rng.standard_normal(m) * 10 ** (dbfs / 20)).astype(np.float32)

    t = 0
    while t < n:
        kind = rng.random()
        if kind < 0.06:
            m = int(rng.uniform(35.0, 50.0) * sr)
            chunk = tone(m, -55.5)
        else:
            m = int(math.exp(rng.uniform(math.log(2.0), math.log(80.0))) * sr)
            chunk = np.empty(m, dtype=np.float32)
            u = 0
            while u < m:
                c = base[int(rng.integers(n_base))]
                o = int(rng.integers(0, len(c) // 2))
                k = min(len(c) - o, m - u)
                chunk[u: u + k] = c[o: o + k]
                u += k
            u = int(rng.uniform(4.0, 18.0) * sr)
            while u < m - sr:
                k = min(int(rng.uniform(1.0, 1.6) * sr), m - sr - u)
                chunk[u: u + k] = tone(k, -68.0)
                u += k + int(rng.uniform(4.0, 24.0) * sr)
        k = min(m, n - t)
        out[t: t + k] = chunk[:k]
        t += k
        if t >= n:
            break
        g = min(int(math.exp(rng.uniform(math.log(0.4), math.log(4.0))) * sr), n - t)
        out[t: t + g] = tone(g, -75.0)
        t += g
    q = np.round(np.clip(out, -1.0, 1.0) * 32767.0).astype(np.int16)
    return (q.astype(np.float32) / 32768.0).astype(np.float32)
