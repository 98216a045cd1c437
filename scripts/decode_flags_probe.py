"""Decode-step placement options measured on the box (wjb_debug_set_decode_flags): steady-state ms per decoder step of whisper-large-v3
with every row alive (EOT suppressed), 64 windows greedy and 64 x beam 2, for each combination of
  1 LayerNorms folded into the step GEMMs (row statistics)   2 L2 prefetch of the next Linear's weights
  4 cross-attention K/V primed before the dependency wait      8 evict-first K/V loads
Writes gpurun_out/decode_flags.json (committed as profiles/<tag>_decode_flags.json)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from whisperjav_b200 import _lib, model as WM  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = _lib.load()
m = WM.load_model("large-v3", max_batch=B)
xa = torch.randn(B, 1500, 1280, device="cuda", dtype=torch.float16) * 0.5
tok = WM.Tokens(m.dims.n_vocab, "ja")
sup = f"-1,{tok.eot}"
res = {}
combos = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1, 2, 4, 8, 1 | 2, 4 | 8, 2 | 4 | 8, 1 | 4 | 8, 15]
for mode, kw in (("greedy", {}), ("beam2", {"beam_size": 2, "patience": 1.2})):
    for rep in range(2):   # two rounds: the second is the one reported (clocks settled, every graph captured once)
        for f in combos:
            lib.wjb_debug_set_decode_flags(f)
            m.decode_features(xa, without_timestamps=True, suppress_tokens=sup, sample_len=12, **kw)
            torch.cuda.synchronize()
            s0 = m.stats["decode_steps"]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = m.decode_features(xa, without_timestamps=True, suppress_tokens=sup, sample_len=100, **kw)
            e1.record()
            torch.cuda.synchronize()
            steps = m.stats["decode_steps"] - s0
            res[f"{mode}_flags{f}"] = {"ms_per_step": e0.elapsed_time(e1) / steps, "steps": steps, "tokens_row0": r[0].tokens[:6]}
    print(json.dumps({k: round(v["ms_per_step"], 3) for k, v in res.items() if k.startswith(mode)}), flush=True)
lib.wjb_debug_set_decode_flags(12)
same = len({tuple(v["tokens_row0"]) for k, v in res.items() if k.startswith("greedy")})
res["note"] = "ms per step includes the once-per-run cross-K/V projection (~25 ms / 103 steps); distinct greedy token heads across flags: %d" % same
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/decode_flags.json").write_text(json.dumps(res, indent=1))
