"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/wjb200.h declares
(no compute calls without a GPU); the product path fails loudly without a device."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from whisperjav_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    src = (ROOT / "include" / "wjb200.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wjb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from whisperjav_b200 import build
    build.build()
    lib = C.CDLL(str(_lib.LIB_PATH))
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in wjb200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_layout_queries_work_without_gpu():
    from whisperjav_b200.synth import DIMS
    from whisperjav_b200.weights import layout
    lib = _lib.load()
    assert lib.wjb_abi_version() == 1
    d = _lib.make_dims(DIMS["large-v3"])
    total = lib.wjb_weights_bytes(C.byref(d))
    # 1.54 B parameters in fp16 (~3.09 GB) plus alignment padding and the fp32 sinusoids
    assert 3.05e9 < total < 3.2e9
    lay = layout(DIMS["tiny"])
    assert "enc.0.qkv.w" in lay and "dec.3.ckv.b" in lay and lay["enc.pos"][2] == 1
    offs = sorted(v[0] for v in lay.values())
    assert all(o % 256 == 0 for o in offs)
    bad = _lib.Dims(80, 1500, 100, 3, 1, 51865, 448, 100, 3, 1)  # head dim != 64
    assert lib.wjb_weights_bytes(C.byref(bad)) == 0


def test_weight_packing_roundtrip_cpu():
    from whisperjav_b200.synth import DIMS, synth_weights
    from whisperjav_b200.weights import hf_to_openai, layout, pack_weights
    d = DIMS["tiny"]
    w = synth_weights(d, seed=3)
    blob = pack_weights(d, w, device="cpu")
    lay = layout(d)
    off, nb, dt = lay["enc.1.qkv.w"]
    got = blob[off: off + nb].view(torch.float16).view(3 * 384, 384)
    assert torch.equal(got[384:768], w["encoder.blocks.1.attn.key.weight"])
    off, nb, dt = lay["enc.conv1.w"]
    got = blob[off: off + nb].view(torch.float16).view(384, 3, 80)
    assert torch.equal(got[:, 2, :], w["encoder.conv1.weight"][:, :, 2])
    off, nb, dt = lay["dec.2.ckv.b"]
    got = blob[off: off + nb].view(torch.float16)
    assert torch.all(got[:384] == 0) and torch.equal(got[384:], w["decoder.blocks.2.cross_attn.value.bias"])
    # HF naming maps onto the same blob
    hf = {k.replace("encoder.blocks.", "model.encoder.layers.").replace(".attn.query.", ".self_attn.q_proj."): v for k, v in w.items()
          if k.startswith("encoder.blocks.0.attn.query")}
    assert "encoder.blocks.0.attn.query.weight" in hf_to_openai(hf)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_product_path_fails_loudly_without_gpu():
    from whisperjav_b200 import model as M
    with pytest.raises(_lib.WjbError):
        M.load_model("tiny")
