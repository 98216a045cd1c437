"""One shortened hot-path step for ncu (launch list / full captures): large-v3, batch B, decode capped at
``--tokens`` sampled tokens so the launch list stays small.  Usage under gpurun:

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --batch 64 --tokens 6
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from whisperjav_b200 import model as M  # noqa: E402
from whisperjav_b200.synth import DIMS, speech_shaped_audio  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="large-v3")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--tokens", type=int, default=6)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--vad", action="store_true", help="also one VAD pass over the same clips (vad_features_kernel / vad_lstm_kernel)")
ap.add_argument("--align", action="store_true", help="also one word-timestamp alignment pass (capture + align_* kernels)")
ap.add_argument("--beam", type=int, default=0, help="also a beam-search decode of that width (beam_select_kernel, NQ-query cross-attention)")
ap.add_argument("--scenes", action="store_true", help="also the scene-split energy gate over a 10 min film-shaped stream (scene_energy_kernel)")
args = ap.parse_args()
dims = DIMS[args.model]
m = M.load_model(args.model, max_batch=args.batch)
base = [speech_shaped_audio(30.0, 2000 + i) for i in range(4)]
audio = torch.stack([torch.from_numpy(base[i % 4]) for i in range(args.batch)]).cuda()
ns = torch.full((args.batch,), audio.shape[1], dtype=torch.int32, device="cuda")
for _ in range(args.reps):
    mel = m.log_mel(audio, ns, n_frames=3000, layout="time")
    xa = m.encode(mel)
    res = m.decode_features(xa, language="ja", without_timestamps=True, sample_len=args.tokens)
if args.vad:
    from whisperjav_b200.vad import VadB200
    VadB200().probs(audio, ns)
if args.beam:
    m.decode_features(xa, language="ja", without_timestamps=False, max_initial_timestamp=0.0, sample_len=args.tokens, beam_size=args.beam, patience=1.2)
if args.scenes:
    from whisperjav_b200.scenes import B200SceneDetector
    from whisperjav_b200.synth import film_audio
    B200SceneDetector().detect(film_audio(600.0, 77), 16000)
if args.align:
    eot = M.Tokens(dims.n_vocab, "ja").eot
    m.align_windows(xa, [[t for t in r.tokens if t < eot] or [11] for r in res], [3000] * args.batch, language="ja")
torch.cuda.synchronize()
print("ok", [len(r.tokens) for r in res][:4])
