"""Host-side mirrors of the reference's ``captioning`` package for the hot path.

Same dotted class names below ``captioning.models`` as the reference, same constructor
kwargs, ``forward(input_dict) -> dict`` keys and ``state_dict`` layout; the compute goes
through the C ABI (audiocaption_b200/_lib.py).  ``audiocaption_b200.install_as_captioning()``
aliases this package as top-level ``captioning``."""
