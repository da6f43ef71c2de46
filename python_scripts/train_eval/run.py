#!/usr/bin/env python3
"""Entry point mirroring python_scripts/train_eval/run.py of the reference (`Runner.train / evaluate / debug`) on the B200
path:  python python_scripts/train_eval/run.py train --config CONFIG.yaml [--key.sub=value ...]

Same YAML schema (model / optimizer / lr_scheduler / trainer / scheduled_sampling / loss / swa / inference_args sections,
`type:` + `args:` objects resolved through captioning.utils.train_util, `inherit_from`, dotted CLI overrides) and the
same per-iteration order as run.py:77-148.  Differences, all on the host side:
  * the hot loop body is the fused `audiocaption_b200.train_step.TrainStep` (trainer.fused: True, default) -- or, with
    trainer.fused: False, the reference's literal loop through the mirrors' autograd wrappers (model(input_dict) ->
    loss_fn(output) -> backward -> clip_grad_norm_ -> optimizer.step());
  * `data:` entries are built with the same reflection factory; the datasets shipped here read wav / npy files or
    generate seeded synthetic clips (captioning/datasets), HDF5 needs h5py which this image lacks;
  * validation monitors CIDEr of the beam-search predictions against the validation captions, as run.py:150-155 does,
    with captioning.metrics.cider (a restatement of pycocoevalcap's scorer, which this image lacks) and a plain
    lower-case / punctuation-stripping tokenizer in place of the Java PTB tokenizer; `trainer.monitor: loss` monitors the
    teacher-forced label-smoothing loss of the validation split instead ("score" = -loss).
`fire`, tensorboard and wandb are replaced by argparse and a plain log file."""
import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = str(Path(__file__).resolve().parents[2])
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import audiocaption_b200  # noqa: E402

audiocaption_b200.install_as_captioning()
import captioning.utils.train_util as train_util  # noqa: E402


class Runner:

    def __init__(self, seed=1):
        self.seed = seed
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

    # ------------------------------------------------------------------ data / model (python_scripts/train_eval/base.py:28-75)
    def _get_dataloaders(self):
        loaders = {}
        for split in ("train", "val"):
            cfg = self.config["data"][split]
            dataset = train_util.init_obj_from_dict(cfg["dataset"])
            collate_cfg = dict(cfg["collate_fn"])
            collate_fn = train_util.init_obj_from_dict(collate_cfg, tokenizer=self.tokenizer) \
                if "tokenizer" in collate_cfg else train_util.init_obj_from_dict(collate_cfg)
            loaders[split] = torch.utils.data.DataLoader(dataset, collate_fn=collate_fn, **cfg.get("dataloader_args", {}))
        return loaders["train"], loaders["val"]

    def load_tokenizer(self):
        cfg = self.config["data"]["train"]["collate_fn"]["tokenizer"]
        tokenizer = train_util.init_obj_from_dict(cfg)
        if not tokenizer.loaded and "vocabulary" in cfg:                    # a word list in the config (synthetic runs)
            for word in cfg["vocabulary"]:
                tokenizer.add_word(word)
        return tokenizer

    def _get_model(self, print_fn=sys.stdout.write):
        model = train_util.init_model_from_config(self.config["model"], print_fn)
        model.set_index(self.tokenizer.bos, self.tokenizer.eos, self.tokenizer.pad)
        return model

    # ------------------------------------------------------------------ forward (run.py:21-47)
    def _forward(self, batch, training=True):
        batch = dict(batch)
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                batch[k] = v.long().to(self.device) if k == "cap" else v.float().to(self.device)
        input_dict = {"mode": "train" if training else "inference"}
        input_dict.update(batch)
        if training:
            input_dict["ss_ratio"] = self.ss_ratio
            input_dict["specaug"] = self.config.get("specaug", False)
            output = self.model(input_dict)
            output["tgt"] = batch["cap"][:, 1:]
            output["tgt_len"] = torch.as_tensor(batch["cap_len"] - 1)
        else:
            input_dict["specaug"] = False
            input_dict.update(self.config["inference_args"])
            output = self.model(input_dict)
        return output

    def _update_ss_ratio(self):
        """run.py:55-65"""
        ss_cfg = self.config.get("scheduled_sampling", {"use": False})
        if not ss_cfg.get("use", False):
            return
        if ss_cfg["mode"] == "exponential":
            self.ss_ratio *= 0.01 ** (1.0 / self.iterations)
        elif ss_cfg["mode"] == "linear":
            self.ss_ratio -= (1.0 - ss_cfg["final_ratio"]) / self.iterations
        else:
            raise Exception(f"mode {ss_cfg['mode']} not supported")

    # ------------------------------------------------------------------ epochs (run.py:77-155)
    def _train_epoch(self):
        total_loss, nsamples = 0.0, 0
        self.model.train()
        losses = []
        def next_batch():
            try:
                return next(self.train_iter)
            except StopIteration:
                self.train_iter = iter(self.train_dataloader)
                return next(self.train_iter)

        for _ in range(self.epoch_length):
            if self.fused and self.look_ahead:
                # one batch of look-ahead: while this iteration's trainable part runs, the NEXT batch is uploaded and goes
                # through the frozen CNN on its own SM set (TrainStep.prefetch)
                if getattr(self, "_staged", None) is None:
                    self._staged = self.train_step.prefetch(next_batch())
                batch, self._staged = self._staged, self.train_step.prefetch(next_batch())
            else:
                batch = next_batch()
            nsample = int(np.sum(batch["cap_len"] - 1))
            if self.fused:
                res = self.train_step.step(batch)                 # schedules, forward, loss, backward, clip, Adam
                self.ss_ratio = self.train_step.ss_ratio
                losses.append((res["loss"], nsample))
            else:
                self._update_ss_ratio()
                self.lr_scheduler.step()
                self.optimizer.zero_grad()
                loss = self.loss_fn(self._forward(batch, training=True))
                if not torch.isnan(loss):
                    loss.backward()
                    torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm)
                    self.optimizer.step()
                    losses.append((loss.detach().reshape(1), nsample))
            self.iteration += 1
        for loss, nsample in losses:                              # one host sync per epoch, not per iteration
            value = float(loss.item())
            if not np.isnan(value):
                total_loss += value * nsample
                nsamples += nsample
        return {"loss": total_loss / max(nsamples, 1)}

    @torch.no_grad()
    def _eval_epoch(self):
        if getattr(self, "train_step", None) is not None:
            self.train_step.wait_look_ahead()                     # the CNN has one workspace: no encoder pass in flight
        if getattr(self, "monitor", "cider") == "cider":
            return self._eval_epoch_cider()
        return self._eval_epoch_loss()

    @torch.no_grad()
    def _eval_epoch_cider(self):
        """run.py:150-155: predictions of the validation split (inference_args: beam search) scored with CIDEr against
        every caption the split holds for the same audio_id."""
        from captioning.metrics.cider import Cider, simple_tokenize
        self.model.eval()
        key2refs, key2pred = {}, {}
        for batch in self.val_dataloader:
            refs = batch["caption"] if "caption" in batch else self.tokenizer.decode(np.asarray(batch["cap"]))
            todo = []
            for i, (aid, ref) in enumerate(zip(batch["audio_id"], refs)):
                key2refs.setdefault(aid, []).append(simple_tokenize(ref))
                if aid not in key2pred:
                    key2pred[aid] = None
                    todo.append(i)
            if todo:                                                     # decode every audio clip once
                wav = torch.as_tensor(batch["wav"])[todo]
                wav_len = np.asarray(batch["wav_len"])[todo]
                output = self._forward({"wav": wav, "wav_len": wav_len}, training=False)
                for i, caption in zip(todo, self.tokenizer.decode(output["seq"].cpu().numpy())):
                    key2pred[batch["audio_id"][i]] = [simple_tokenize(caption)]
        score, _ = Cider().compute_score(key2refs, key2pred)
        return {"score": score}

    @torch.no_grad()
    def _eval_epoch_loss(self):
        from captioning.losses.loss import ls_ce_fwd_bwd
        self.model.eval()
        total, count = 0.0, 0
        for batch in self.val_dataloader:
            wav = torch.as_tensor(batch["wav"]).float().to(self.device)
            cap = torch.as_tensor(batch["cap"]).long().to(self.device)
            enc = self.model.encoder({"wav": wav, "wav_len": batch["wav_len"], "specaug": False})
            out = self.model.seq_forward({"cap": cap, "attn_emb": enc["attn_emb"], "attn_emb_len": enc["attn_emb_len"]})
            tl = torch.as_tensor(batch["cap_len"] - 1).to(self.device)
            loss, _ = ls_ce_fwd_bwd(out["logit"], cap[:, 1:], tl, self.smoothing, want_grad=False)
            n = int(np.sum(batch["cap_len"] - 1))
            total += float(loss.item()) * n
            count += n
        return {"score": -total / max(count, 1)}

    @torch.no_grad()
    def _inference(self, dataloader):
        """python_scripts/train_eval/base.py:212-224"""
        self.model.eval()
        key2pred = {}
        for batch in dataloader:
            output = self._forward({k: v for k, v in batch.items() if k in ("wav", "wav_len")}, training=False)
            for aid, caption in zip(batch["audio_id"], self.tokenizer.decode(output["seq"].cpu().numpy())):
                key2pred[aid] = [caption]
        return key2pred

    # ------------------------------------------------------------------ checkpoints (base.py:231-264, run.py:209-216)
    def save_checkpoint(self, path):
        model_dict = self.model.state_dict()
        ckpt = {"model": {k: model_dict[k].detach().cpu().clone() for k in self.saving_keys}, "epoch": self.epoch,
                "metric_monitor": {"best": self.best_score}, "not_improve_cnt": self.not_improve_cnt,
                "tokenizer": self.tokenizer.state_dict()}
        torch.save(ckpt, path)

    # ------------------------------------------------------------------ train (run.py:158-361)
    def train(self, config, **kwargs):
        from audiocaption_b200.train_step import TrainStep
        self.config = train_util.parse_config_or_kwargs(config, **kwargs)
        self.seed = self.config.get("seed", self.seed)
        train_util.set_seed(self.seed)
        exp_dir = Path(self.config["experiment_path"]) / f"seed_{self.seed}"
        exp_dir.mkdir(parents=True, exist_ok=True)
        log = open(exp_dir / "train.log", "a")

        def info(msg):
            print(msg)
            log.write(msg + "\n")
            log.flush()

        self.tokenizer = self.load_tokenizer()
        self.train_dataloader, self.val_dataloader = self._get_dataloaders()
        self.train_iter = iter(self.train_dataloader)
        trainer = self.config["trainer"]
        self.epochs = trainer["epochs"]
        self.epoch_length = trainer.get("epoch_length", len(self.train_dataloader))
        self.iterations = self.epochs * self.epoch_length
        self.model = self._get_model(info).to(self.device)
        # run.py:209-216: trainable parameters + ALL buffers are saved (the frozen CNN is re-loaded from `pretrained:`)
        self.saving_keys = [k for k, p in self.model.named_parameters() if p.requires_grad] + \
            [k for k, _ in self.model.named_buffers()]
        swa_cfg = self.config.get("swa", {"use": False})
        swa_model = train_util.AveragedModel(self.model) if swa_cfg.get("use", False) else None
        self.max_grad_norm = trainer.get("max_grad_norm", 1.0)
        self.smoothing = self.config["loss"]["args"].get("smoothing", 0.0)
        self.fused = trainer.get("fused", True)
        self.monitor = trainer.get("monitor", "cider")                           # "cider" (run.py:150-155) or "loss"
        self.look_ahead = trainer.get("look_ahead", True)                        # TrainStep.prefetch one batch ahead
        sched_args = dict(self.config["lr_scheduler"]["args"])
        warmup = sched_args.get("warmup_iters", self.iterations // 5)            # run.py:249-251
        opt_args = self.config["optimizer"]["args"]
        ss_cfg = self.config.get("scheduled_sampling", {"use": False})
        self.ss_ratio = 1.0
        if self.fused:
            self.train_step = TrainStep(self.model, total_iters=self.iterations, lr=opt_args["lr"],
                                        weight_decay=opt_args.get("weight_decay", 0.0), max_grad_norm=self.max_grad_norm,
                                        smoothing=self.smoothing, final_lr=sched_args["final_lrs"], warmup_iters=warmup,
                                        ss_mode=ss_cfg.get("mode", "linear"), ss_final_ratio=ss_cfg.get("final_ratio", 1.0),
                                        use_ss=ss_cfg.get("use", False), specaug=self.config.get("specaug", False))
        else:
            self.optimizer = train_util.init_obj_from_dict(
                self.config["optimizer"], params=[p for p in self.model.parameters() if p.requires_grad])
            self.loss_fn = train_util.init_obj_from_dict(self.config["loss"])
            sched_cfg = dict(self.config["lr_scheduler"], args=dict(sched_args, total_iters=self.iterations, warmup_iters=warmup))
            self.lr_scheduler = train_util.init_obj_from_dict(sched_cfg, optimizer=self.optimizer)
        self.config.setdefault("inference_args", {"sample_method": "beam", "beam_size": 3})
        train_util.store_yaml(self.config, exp_dir / "config.yaml")               # read back by inference.py
        self.iteration, self.epoch, self.not_improve_cnt, self.best_score = 0, 1, 0, -float("inf")
        info(f"{sum(p.numel() for p in self.model.parameters())} parameters, "
             f"{sum(p.numel() for p in self.model.parameters() if p.requires_grad)} trainable; {self.iterations} iterations")
        for self.epoch in range(1, self.epochs + 1):
            t0 = time.time()
            train_out = self._train_epoch()
            val_out = self._eval_epoch()
            info(f"epoch {self.epoch}: train loss {train_out['loss']:.4f}  val score {val_out['score']:.4f}  "
                 f"ss_ratio {self.ss_ratio:.4f}  {time.time() - t0:.1f} s")
            if val_out["score"] > self.best_score:
                self.best_score, self.not_improve_cnt = val_out["score"], 0
                self.save_checkpoint(exp_dir / "best.pth")
            else:
                self.not_improve_cnt += 1
            if self.epoch % trainer.get("save_interval", 1) == 0:
                self.save_checkpoint(exp_dir / "last.pth")
            if swa_model is not None and self.epoch >= swa_cfg.get("start", self.epochs + 1):
                swa_model.update_parameters(self.model)
        if swa_model is not None and int(swa_model.n_averaged) > 0:                # run.py:350-355
            sd = swa_model.module.state_dict()
            torch.save({"model": {k: sd[k].detach().cpu() for k in self.saving_keys}, "tokenizer": self.tokenizer.state_dict()},
                       exp_dir / "swa.pth")
        log.close()
        return str(exp_dir)

    def debug(self, config, **kwargs):
        """run.py:363-378: one real batch -> forward -> loss -> backward."""
        self.config = train_util.parse_config_or_kwargs(config, **kwargs)
        train_util.set_seed(self.config.get("seed", self.seed))
        self.tokenizer = self.load_tokenizer()
        self.train_dataloader, self.val_dataloader = self._get_dataloaders()
        self.model = self._get_model().to(self.device).train()
        self.loss_fn = train_util.init_obj_from_dict(self.config["loss"])
        self.ss_ratio = 0.9
        batch = next(iter(self.train_dataloader))
        loss = self.loss_fn(self._forward(batch, training=True))
        loss.backward()
        print(f"forward and backward done, loss {loss.item():.4f}")
        return float(loss.item())


def _parse_cli(argv):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("command", choices=["train", "debug"])
    ap.add_argument("--config", required=True)
    args, extra = ap.parse_known_args(argv)
    overrides = {}
    for item in extra:                                  # --a.b=value, values parsed as YAML scalars (fire-style overrides)
        if not item.startswith("--") or "=" not in item:
            raise SystemExit(f"cannot parse override {item!r} (expected --key.sub=value)")
        key, value = item[2:].split("=", 1)
        import yaml
        overrides[key] = yaml.safe_load(value)
    return args, overrides


if __name__ == "__main__":
    cli, over = _parse_cli(sys.argv[1:])
    getattr(Runner(), cli.command)(cli.config, **over)
