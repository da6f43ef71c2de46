"""Entry points and input pipeline: tokenizer / collate / dataset host logic (CPU), the on-device resampler against
torchaudio (GPU), and the two scripts python_scripts/train_eval/run.py + python_scripts/inference/inference.py run end to
end as subprocesses on a small synthetic experiment (GPU)."""
import json
import os
import subprocess
import sys
import wave

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


# ------------------------------------------------------------------ host logic (CPU)
def test_dict_tokenizer_roundtrip_and_state_dict():
    from audiocaption_b200.captioning.datasets.text_tokenizer import DictTokenizer
    tok = DictTokenizer(max_length=5)
    for w in "a dog barks while rain falls loudly outside".split():
        tok.add_word(w)
    assert (tok.pad, tok.bos, tok.eos) == (0, 1, 2) and len(tok) == 12
    out = tok(["a dog barks", "rain falls loudly outside while a dog barks", "a zebra"])
    assert out["cap"].shape == (3, 7) and out["cap_len"].tolist() == [5, 7, 4]          # truncated to max_length + 2
    assert out["cap"][2].tolist() == [1, 4, 3, 2, 0, 0, 0]                               # <unk> = 3, <pad> = 0
    assert tok.decode(out["cap"].numpy()) == ["a dog barks", "rain falls loudly outside while", "a <unk>"]
    other = DictTokenizer()
    other.load_state_dict(tok.state_dict())
    assert other.decode(out["cap"].numpy()) == tok.decode(out["cap"].numpy())


def test_text_collate_pads_sorts_and_tokenises():
    from audiocaption_b200.captioning.datasets.caption_dataset import SyntheticCaptionDataset
    from audiocaption_b200.captioning.datasets.collate_func import TextCollate
    from audiocaption_b200.captioning.datasets.text_tokenizer import DictTokenizer
    ds = SyntheticCaptionDataset(size=5, n_samples=4000, vocab_size=50, ragged=True, seed=3)
    tok = DictTokenizer()
    for w in ds.vocabulary():
        tok.add_word(w)
    assert len(tok) == 50
    batch = TextCollate(tok)([ds[i] for i in range(5)])
    assert batch["wav"].shape[0] == 5 and batch["wav"].shape[1] == int(batch["wav_len"].max())
    assert all(2000 <= n <= 4000 for n in batch["wav_len"])
    assert list(batch["cap_len"]) == sorted(batch["cap_len"], reverse=True)              # longest caption first
    assert (batch["cap"][:, 0] == tok.bos).all() and batch["cap"].max() < 50
    for i, n in enumerate(batch["wav_len"]):
        assert (batch["wav"][i, n:] == 0).all()
    assert ds[2]["caption"] == ds[2]["caption"]                                          # seeded


def test_resample_table_matches_torchaudio():
    import math
    from torchaudio.functional.functional import _get_sinc_resample_kernel
    from audiocaption_b200.resample import sinc_resample_kernel
    for o, n in ((32000, 16000), (44100, 32000), (16000, 32000)):
        coef, orig, new, width = sinc_resample_kernel(o, n)
        ref, ref_width = _get_sinc_resample_kernel(o, n, math.gcd(o, n))
        assert width == ref_width and coef.shape == ref[:, 0].shape
        assert (coef - ref[:, 0]).abs().max() < 2e-5
    assert sinc_resample_kernel(32000, 16000)[0].shape == (1, 28)                        # 28 taps, stride 2


# ------------------------------------------------------------------ device resampler vs torchaudio (the reference's call)
@pytest.mark.gpu
@pytest.mark.parametrize("orig,new,n", [(32000, 16000, 320000), (44100, 32000, 44100 * 3 + 17), (16000, 32000, 5001),
                                        (32000, 16000, 27)])
def test_resample_matches_torchaudio(orig, new, n):
    import torchaudio
    from audiocaption_b200.resample import resample
    g = torch.Generator().manual_seed(n)
    wav = 0.1 * torch.randn(3, n, generator=g)
    want = torchaudio.functional.resample(wav, orig, new)
    got = resample(wav.to(DEV), orig, new).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max() < 2e-5          # 459-tap phases (44.1 -> 32 kHz): fp32 summation order


# ------------------------------------------------------------------ run.py train -> inference.py, as subprocesses
def _write_config(path, exp_dir, fused=True):
    vocab = 300
    words = [f"w{i}" for i in range(4, vocab)]
    data = {"dataset": {"type": "captioning.datasets.caption_dataset.SyntheticCaptionDataset",
                        "args": {"size": 8, "n_samples": 48000, "vocab_size": vocab, "min_words": 4, "max_words": 8, "ragged": True}},
            "collate_fn": {"type": "captioning.datasets.collate_func.TextCollate", "args": {},
                           "tokenizer": {"type": "captioning.datasets.text_tokenizer.DictTokenizer", "args": {"max_length": 20},
                                         "vocabulary": words}},
            "dataloader_args": {"batch_size": 4, "shuffle": False, "num_workers": 0}}
    val = json.loads(json.dumps(data))
    val["dataset"]["args"]["seed"] = 9
    with open(os.path.join(ROOT, "tests", "golden", "eg_configs", "clotho_v2_cnn14rnn_trm.yaml")) as f:
        model = yaml.safe_load(f)["model"]                       # the reference's own model section
    model["decoder"]["args"]["vocab_size"] = vocab
    cfg = {"experiment_path": str(exp_dir), "seed": 1, "model": model, "specaug": False, "data": {"train": data, "val": val},
           "optimizer": {"type": "torch.optim.Adam", "args": {"lr": 5e-4, "weight_decay": 1e-6}},
           "lr_scheduler": {"type": "captioning.utils.lr_scheduler.ExponentialDecayScheduler", "args": {"final_lrs": 5e-7}},
           "trainer": {"max_grad_norm": 1.0, "epochs": 3, "save_interval": 1, "fused": fused,
                       "monitor": "cider" if fused else "loss"},      # CIDEr of the beam-search predictions (run.py:150-155) / -val loss
           "inference_args": {"sample_method": "beam", "beam_size": 3},
           "scheduled_sampling": {"use": True, "mode": "linear", "final_ratio": 0.7},
           "loss": {"type": "captioning.losses.loss.LabelSmoothingLoss", "args": {"smoothing": 0.1}},
           "swa": {"use": True, "start": 2}}
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_run_py_train_then_inference_py(tmp_path, fused):
    cfg = tmp_path / "cfg.yaml"
    _write_config(cfg, tmp_path / "exp", fused)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "python_scripts", "train_eval", "run.py"), "train", "--config", str(cfg),
                        "--trainer.epochs=3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    exp = tmp_path / "exp" / "seed_1"
    for name in ("config.yaml", "best.pth", "last.pth", "swa.pth", "train.log"):
        assert (exp / name).exists(), name
    log = (exp / "train.log").read_text()
    losses = [float(line.split("train loss")[1].split()[0]) for line in log.splitlines() if "train loss" in line]
    assert len(losses) == 3 and all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    ckpt = torch.load(exp / "best.pth", map_location="cpu")
    assert set(ckpt) >= {"model", "epoch", "metric_monitor", "not_improve_cnt", "tokenizer"}
    assert not any(k.startswith("encoder.cnn.conv_block") and k.endswith("conv1.weight") for k in ckpt["model"])   # frozen: not saved
    assert any(k.startswith("encoder.cnn.bn0.running_mean") for k in ckpt["model"])                               # buffers are
    assert "decoder.classifier.weight" in ckpt["model"] and "encoder.rnn.network.weight_ih_l0" in ckpt["model"]
    # a 44.1 kHz 16-bit wav file -> inference.py (resampled on the device to 32 kHz)
    wav_path = tmp_path / "clip.wav"
    g = np.random.default_rng(0)
    pcm = (0.1 * g.standard_normal(44100 * 2) * 32767).astype("<i2")
    with wave.open(str(wav_path), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(44100)
        f.writeframes(pcm.tobytes())
    out = tmp_path / "pred.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "python_scripts", "inference", "inference.py"), "--input", str(wav_path),
                        "--output", str(out), "--checkpoint", str(exp / "swa.pth")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    pred = json.loads(out.read_text())["predictions"]
    assert len(pred) == 1 and pred[0]["filename"] == "clip.wav" and isinstance(pred[0]["tokens"], str)
