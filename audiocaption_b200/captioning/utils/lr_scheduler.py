"""Mirror of captioning/utils/lr_scheduler.py:5-45 `ExponentialDecayScheduler`: linear warm-up to the base learning rate
over `warmup_iters` scheduler steps, then a geometric decay that reaches `final_lrs` at `total_iters`.

`exponential_decay_lr` is the closed form (a pure function of the scheduler's step count) that the fused train step
(audiocaption_b200/train_step.py) evaluates on the host each iteration; the class wraps it in torch's scheduler protocol
so `lr_scheduler.step()` / `optimizer.param_groups[i]["lr"]` behave as in python_scripts/train_eval/run.py:105-110."""
from torch.optim.lr_scheduler import LRScheduler


def exponential_decay_lr(step_count: int, base_lr: float, final_lr: float, total_iters: int, warmup_iters: int) -> float:
    """Learning rate after the scheduler has been stepped `step_count` times in total (the constructor itself performs
    the first step, so the k-th training iteration -- which steps the scheduler before the optimizer, run.py:105 --
    sees step_count = k + 1).  lr_scheduler.py:22-42."""
    if step_count <= warmup_iters:
        coeff = step_count / warmup_iters if step_count < warmup_iters else 1.0
        return coeff * base_lr
    base = (final_lr / base_lr) ** (1.0 / (total_iters - warmup_iters))
    return base_lr * base ** (step_count - warmup_iters)


class ExponentialDecayScheduler(LRScheduler):

    def __init__(self, optimizer, total_iters, final_lrs, warmup_iters=3000, last_epoch=-1, verbose=False):
        self.total_iters = total_iters
        self.warmup_iters = warmup_iters
        n_groups = len(optimizer.param_groups)
        self.final_lrs = list(final_lrs) if isinstance(final_lrs, (list, tuple)) else [final_lrs] * n_groups
        super().__init__(optimizer, last_epoch)          # performs the first step (step count 1)

    def _get_closed_form_lr(self):
        return [exponential_decay_lr(self._step_count, base_lr, final_lr, self.total_iters, self.warmup_iters)
                for base_lr, final_lr in zip(self.base_lrs, self.final_lrs)]

    def get_lr(self):
        return self._get_closed_form_lr()
