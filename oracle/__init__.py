"""CPU oracle for the AudioCaption hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker or as the timed
CPU baseline -- never as the thing shipped.  The product path
(``audiocaption_b200``) never imports this package and fails loudly when its
CUDA library is missing.

Contents
--------
``audio_frontend``   fp32 restatement of the torchaudio MelSpectrogram +
                     AmplitudeToDB call sites (reference
                     captioning/models/hf_wrapper.py:269-279,292-293 and
                     captioning/models/cnn_encoder.py:338-350,418-419).
``efficientnet_b2``  restatement of the un-vendored third-party dependency
                     ``efficientnet_pytorch==0.7.1`` (requirements.txt:8) as used at
                     hf_wrapper.py:218-241.  PARITY UNPINNED at the third-party
                     boundary: the package is absent from this image, so the
                     restatement is anchored on the reference's own in-repo
                     constructor restatement (captioning/models/eff_latent_encoder.py
                     :76-206, key list :263-290) and on output shapes.
``caption_model``    restatement of the Transformer caption decoder and the
                     greedy / beam decoding loops (hf_wrapper.py:389-726,845-1068),
                     pinned against the imported reference (see gen_golden.py).
``cnn14`` / ``crnn``  Cnn14 body (cnn_encoder.py:32-75,414-464), explicit-loop bi-GRU with packed-sequence
                     semantics (rnn_encoder.py:34-49, model_util.py:10-27), CrnnEncoder and the
                     Cnn14Rnn-Transformer captioner; pinned against the imported classes.
``bah_decoder``      temporal Bahdanau-attention GRU decoder + greedy / beam loops with state re-ordering
                     (hf_wrapper.py:1377-1788); pinned exactly against the imported classes.
``sed``              CNN8 + bi-GRU sound-event tagger and its numpy post-processing (hf_wrapper.py:54-216,
                     1791-1859) restated at frame level; pinned against the imported class.
``ref_import``       imports the real reference from /root/reference with import
                     stubs (build container only; never used on the GPU box).
``gen_golden``       regenerates tests/golden/*.npz from the imported reference.
"""
