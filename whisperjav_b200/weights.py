"""Pack a Whisper ``state_dict`` into the device blob libwjb200.so consumes.

Stands where ``whisper.load_model(name, device)`` stands in the reference
(whisperjav/modules/whisper_pro_asr.py:182): it owns weight placement.  Accepts openai-whisper
naming (``encoder.blocks.0.attn.query.weight``) or HF ``transformers`` naming
(``model.encoder.layers.0.self_attn.q_proj.weight``); fused / re-ordered tensors
(q|k|v, cross k|v, k-major conv taps) are assembled here.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib


def hf_to_openai(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Rename HF WhisperForConditionalGeneration keys to openai-whisper keys."""
    out = {}
    rep = [
        ("model.encoder.layers.", "encoder.blocks."), ("model.decoder.layers.", "decoder.blocks."),
        ("model.encoder.", "encoder."), ("model.decoder.", "decoder."),
        (".self_attn_layer_norm.", ".attn_ln."), (".encoder_attn_layer_norm.", ".cross_attn_ln."),
        (".final_layer_norm.", ".mlp_ln."), (".self_attn.", ".attn."), (".encoder_attn.", ".cross_attn."),
        (".q_proj.", ".query."), (".k_proj.", ".key."), (".v_proj.", ".value."), (".out_proj.", ".out."),
        (".fc1.", ".mlp.0."), (".fc2.", ".mlp.2."),
        ("encoder.embed_positions.weight", "encoder.positional_embedding"),
        ("decoder.embed_positions.weight", "decoder.positional_embedding"),
        ("decoder.embed_tokens.weight", "decoder.token_embedding.weight"),
        ("encoder.layer_norm.", "encoder.ln_post."), ("decoder.layer_norm.", "decoder.ln."),
    ]
    for k, v in sd.items():
        if k.startswith("proj_out."):
            continue
        for a, b in rep:
            k = k.replace(a, b)
        out[k] = v
    return out


def _source(name: str, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Blob tensor ``name`` (see csrc/api.cu::build_layout) from an openai-named state_dict."""
    def g(k):
        return sd[k]

    if name == "enc.conv1.w" or name == "enc.conv2.w":
        w = g("encoder.conv1.weight" if name == "enc.conv1.w" else "encoder.conv2.weight")  # [n, C, 3]
        return w.permute(0, 2, 1).reshape(w.shape[0], -1)  # [n, 3*C], tap-major
    simple = {
        "enc.conv1.b": "encoder.conv1.bias", "enc.conv2.b": "encoder.conv2.bias",
        "enc.pos": "encoder.positional_embedding",
        "enc.ln_post.g": "encoder.ln_post.weight", "enc.ln_post.b": "encoder.ln_post.bias",
        "dec.emb": "decoder.token_embedding.weight", "dec.pos": "decoder.positional_embedding",
        "dec.ln.g": "decoder.ln.weight", "dec.ln.b": "decoder.ln.bias",
    }
    if name in simple:
        return g(simple[name])
    side, idx, *rest = name.split(".")
    leaf = ".".join(rest)
    p = f"{'encoder' if side == 'enc' else 'decoder'}.blocks.{idx}."
    ln_names = {"ln1": "attn_ln", "ln2": "mlp_ln"} if side == "enc" else {"ln1": "attn_ln", "ln2": "cross_attn_ln", "ln3": "mlp_ln"}
    part, kind = leaf.split(".")
    if part in ln_names:
        return g(p + ln_names[part] + (".weight" if kind == "g" else ".bias"))
    wb = "weight" if kind == "w" else "bias"
    if part == "qkv":
        q, k, v = g(p + f"attn.query.{wb}"), g(p + "attn.key.weight"), g(p + f"attn.value.{wb}")
        if kind == "w":
            return torch.cat([q, k, v], 0)
        return torch.cat([q, torch.zeros_like(q), v], 0)  # key has no bias
    if part == "out":
        return g(p + f"attn.out.{wb}")
    if part == "cq":
        return g(p + f"cross_attn.query.{wb}")
    if part == "ckv":
        if kind == "w":
            return torch.cat([g(p + "cross_attn.key.weight"), g(p + "cross_attn.value.weight")], 0)
        vb = g(p + "cross_attn.value.bias")
        return torch.cat([torch.zeros_like(vb), vb], 0)
    if part == "cout":
        return g(p + f"cross_attn.out.{wb}")
    if part == "fc1":
        return g(p + f"mlp.0.{wb}")
    if part == "fc2":
        return g(p + f"mlp.2.{wb}")
    raise KeyError(name)


def layout(dims) -> Dict[str, tuple]:
    """name -> (offset, nbytes, dtype) straight from the C side."""
    lib = _lib.load()
    d = _lib.make_dims(dims)
    out = {}
    buf = C.create_string_buffer(64)
    off, nb, dt = C.c_size_t(), C.c_size_t(), C.c_int()
    for i in range(lib.wjb_weight_count(C.byref(d))):
        _lib.check(lib.wjb_weight_info(C.byref(d), i, buf, 64, C.byref(off), C.byref(nb), C.byref(dt)), "wjb_weight_info")
        out[buf.value.decode()] = (off.value, nb.value, dt.value)
    return out


def pack_weights(dims, state_dict: Dict[str, torch.Tensor], device="cuda") -> torch.Tensor:
    """Return the uint8 device blob for ``state_dict`` (openai or HF naming)."""
    lib = _lib.load()
    if any(k.startswith("model.") for k in state_dict):
        state_dict = hf_to_openai(state_dict)
    d = _lib.make_dims(dims)
    total = lib.wjb_weights_bytes(C.byref(d))
    if total == 0:
        raise _lib.WjbError("unsupported model dimensions")
    host = torch.zeros(total, dtype=torch.uint8)
    for name, (off, nb, dt) in layout(dims).items():
        src = _source(name, state_dict).detach().to("cpu").contiguous()
        src = src.to(torch.float32 if dt == 1 else torch.float16).contiguous()
        raw = src.view(torch.uint8).reshape(-1)
        if raw.numel() != nb:
            raise _lib.WjbError(f"{name}: expected {nb} bytes, state_dict gives {raw.numel()}")
        host[off:off + nb] = raw
    return host.to(device)
