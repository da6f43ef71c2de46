#!/usr/bin/env python3
"""Entry point mirroring python_scripts/inference/inference.py:114-183 of the reference on the B200 path:

    python python_scripts/inference/inference.py --input wav.csv|clip.wav --output pred.json --checkpoint EXP/best.pth
           [--batch_size 32] [--original_sr SR] [--target_sr 32000] [--min_duration 0.32] [--sample_method beam] [--beam_size 3]

`config.yaml` is read from the checkpoint's directory, the model tree is built through the reflection factory, clips
shorter than `min_duration` are dropped (:97-99) and predictions are written as {"predictions": [{"filename",
"tokens"}]}.  The waveforms are resampled to `target_sr` ON THE DEVICE (audiocaption_b200.resample, the polyphase kernel
of csrc/resample.cu) instead of with torchaudio in the data-loader workers."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = str(Path(__file__).resolve().parents[2])
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import audiocaption_b200  # noqa: E402

audiocaption_b200.install_as_captioning()
import captioning.utils.train_util as train_util  # noqa: E402
from audiocaption_b200.resample import resample  # noqa: E402
from captioning.datasets.caption_dataset import InferenceDataset  # noqa: E402


def load_model(cfg, ckpt_path, device):
    """inference.py:18-29"""
    model = train_util.init_model_from_config(cfg["model"], lambda *_: None)
    ckpt = torch.load(ckpt_path, "cpu")
    train_util.load_pretrained_model(model, ckpt, lambda *_: None)
    model = model.eval().to(device)
    tok_cfg = cfg["data"]["train"]["collate_fn"]["tokenizer"]
    tokenizer = train_util.init_obj_from_dict({k: v for k, v in tok_cfg.items() if k in ("type", "args")})
    if not tokenizer.loaded:
        tokenizer.load_state_dict(ckpt["tokenizer"])
    model.set_index(tokenizer.bos, tokenizer.eos, tokenizer.pad)
    return model, tokenizer


def inference(input, output, checkpoint, batch_size=32, original_sr=None, target_sr=32000, min_duration=0.32,
              sample_method="beam", beam_size=3):
    device = torch.device("cuda")
    exp_dir = Path(checkpoint).parent
    cfg = train_util.parse_config_or_kwargs(exp_dir / "config.yaml")
    model, tokenizer = load_model(cfg, checkpoint, device)
    if Path(input).suffix == ".csv":
        import pandas as pd
        df = pd.read_csv(input, sep="\t")
        aid_to_fname = dict(zip(df["audio_id"], df["file_name"]))
    else:
        aid_to_fname = {Path(input).name: input}
    dataset = InferenceDataset(aid_to_fname)
    captions, audio_ids = [], []
    with torch.no_grad():
        for b0 in range(0, len(dataset), batch_size):
            items = [dataset[i] for i in range(b0, min(len(dataset), b0 + batch_size))]
            by_sr = {}
            for it in items:                                    # one resample launch per source sample rate
                sr = it["sample_rate"] or original_sr
                assert sr is not None, "original sample rate must be provided"
                by_sr.setdefault(int(sr), []).append(it)
            wavs, aids = [], []
            for sr, group in by_sr.items():
                n = max(len(it["wav"]) for it in group)
                host = np.zeros((len(group), n), dtype=np.float32)
                for i, it in enumerate(group):
                    host[i, :len(it["wav"])] = it["wav"]
                dev_wav = resample(torch.from_numpy(host).to(device), sr, target_sr)
                for i, it in enumerate(group):
                    n_out = -(-len(it["wav"]) * target_sr // sr)
                    if n_out >= min_duration * target_sr:
                        wavs.append(dev_wav[i, :n_out])
                        aids.append(it["audio_id"])
            if not wavs:
                continue
            lens = torch.tensor([w.shape[0] for w in wavs])
            wav = torch.zeros(len(wavs), int(lens.max()), device=device)
            for i, w in enumerate(wavs):
                wav[i, :w.shape[0]] = w
            input_dict = {"mode": "inference", "wav": wav, "wav_len": lens, "specaug": False, "sample_method": sample_method}
            if sample_method == "beam":
                input_dict["beam_size"] = beam_size
            seq = model(input_dict)["seq"].cpu().numpy()
            captions.extend(tokenizer.decode(seq))
            audio_ids.extend(aids)
    data = {"predictions": [{"filename": aid, "tokens": cap} for aid, cap in zip(audio_ids, captions)]}
    Path(output).parent.mkdir(parents=True, exist_ok=True)
    with open(output, "w") as f:
        json.dump(data, f, indent=4)
    return data


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True)
    ap.add_argument("--output", required=True)
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--batch_size", type=int, default=32)
    ap.add_argument("--original_sr", type=int, default=None)
    ap.add_argument("--target_sr", type=int, default=32000)
    ap.add_argument("--min_duration", type=float, default=0.32)
    ap.add_argument("--sample_method", default="beam")
    ap.add_argument("--beam_size", type=int, default=3)
    a = ap.parse_args()
    inference(a.input, a.output, a.checkpoint, a.batch_size, a.original_sr, a.target_sr, a.min_duration, a.sample_method,
              a.beam_size)
