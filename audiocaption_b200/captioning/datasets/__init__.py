"""Host-side input pipeline pieces the entry points need (captioning/datasets of the reference): the word-level tokenizer,
the padding collate functions and two small datasets.  Pure host logic; the heavy lifting of the reference's pipeline that
belongs on the device (resampling) lives in audiocaption_b200/resample.py."""
