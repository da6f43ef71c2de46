"""Caption metrics the training loop monitors (python_scripts/train_eval/run.py:150-155 of the reference)."""
