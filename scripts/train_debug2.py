import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import caption_model as cm, crnn, cnn14 as oc, train_step as ts
import test_train_gpu as T
from audiocaption_b200.train_step import TrainStep
DEV = "cuda:0"
vocab = 4368
for pattern in sys.argv[1:] or ["all_sampled"]:
    m = T._no_dropout(T._train_model(vocab))
    wav, lens = cm.synth_wav(3, 64000, seed=21, ragged=True, varied=True, sample_rate=32000)
    cap, cap_len = ts.synth_captions(3, 7, vocab, seed=4)
    L = cap.size(1) - 1
    coins = {"all_gt": [True] * L, "all_sampled": [False] * L, "last_sampled": [True] * (L - 1) + [False],
             "mixed": [t % 2 == 1 for t in range(L)]}[pattern]
    step = TrainStep(m, total_iters=1000, lr=3e-4 * 5, warmup_iters=10, final_lr=1e-7)
    dec = crnn.build_decoder(6, vocab_size=vocab)
    o = ts.train_step(oc.build_state_dict(3), crnn.build_gru_state_dict(4), dec, wav, lens, cap, cap_len, coins, 3e-4, dtype=torch.float64)
    res = step.step({"wav": wav, "wav_len": lens, "cap": cap, "cap_len": cap_len.numpy()}, coins=coins)
    print(pattern, "cap_len", cap_len.tolist(), "loss", res["loss"].item(), float(o["loss"]))
    print("seq mine\n", step.last_output["seq"].cpu(), "\nseq oracle\n", o["output"]["seq"])
    lg = o["output"]["logit"]; top2 = lg.topk(2, dim=-1).values
    print("min top2 gap per row", (top2[..., 0] - top2[..., 1]).min(1).values)
    print("logit err", (step.last_output["logit"].cpu().double() - lg).abs().max().item())
    for k, p in m.named_parameters():
        if p.requires_grad:
            print(f"{k:55s} {T._relerr(p.grad, o['grads'][k]):.2e}")
