"""CPU tests of the drop-in boundary: the YAML reflection factory resolves the reference's `type:` strings
to the B200 mirrors, the reference's own `init_model_from_config` builds the model tree through them, and the resulting
modules load the reference's state_dict strictly."""
import importlib
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch
import yaml

from oracle import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG_DIR = os.path.join(ROOT, "tests", "golden", "eg_configs")          # verbatim model sections of the reference YAMLs


def _drop_captioning_modules():
    for name in [n for n in sys.modules if n == "captioning" or n.startswith("captioning.")]:
        del sys.modules[name]


@pytest.fixture
def alias():
    """`captioning.*` -> the B200 mirrors for the duration of one test (other test files import the REAL reference
    under that name through oracle.ref_import, so the alias must not leak)."""
    _drop_captioning_modules()
    import audiocaption_b200
    yield audiocaption_b200.install_as_captioning()
    _drop_captioning_modules()


def _model_cfg(name):
    with open(os.path.join(CFG_DIR, name)) as f:
        return yaml.safe_load(f)["model"]


def _type_strings(cfg):
    out = []
    if isinstance(cfg, dict):
        if "type" in cfg:
            out.append(cfg["type"])
        for v in cfg.values():
            out += _type_strings(v)
    return out


def test_install_as_captioning_aliases_every_mirror_module(alias):
    aliased = alias
    for must in ("captioning", "captioning.models", "captioning.models.cnn_encoder", "captioning.models.rnn_encoder",
                 "captioning.models.crnn_trm_encoder", "captioning.models.transformer_decoder",
                 "captioning.models.transformer_model", "captioning.models.hf_wrapper", "captioning.models.base",
                 "captioning.utils.train_util", "captioning.utils.lr_scheduler", "captioning.losses.loss"):
        assert must in aliased
        mod = importlib.import_module(must)                      # what train_util.py:63-68 does
        assert mod.__name__.startswith("audiocaption_b200.captioning"), must


@pytest.mark.parametrize("yaml_name", ["audiocaps_cnn14rnn_trm.yaml", "clotho_v2_cnn14rnn_trm.yaml"])
def test_every_type_string_of_the_training_yaml_resolves(yaml_name, alias):
    from audiocaption_b200.captioning.utils import train_util
    cfg = _model_cfg(yaml_name)
    types = _type_strings(cfg)
    assert len(types) == 5
    for t in types:
        cls = train_util.get_cls_from_str(t)
        assert cls.__module__.startswith("audiocaption_b200.captioning"), t


@pytest.mark.parametrize("yaml_name", ["audiocaps_cnn14rnn_trm.yaml", "clotho_v2_cnn14rnn_trm.yaml"])
def test_factory_builds_the_yaml_model_and_loads_reference_state_dict(yaml_name, alias):
    """Our mirror of init_model_from_config on the reference's YAML model section: class tree, frozen CNN, trainable
    parameter count (SURVEY 8c probe: 10.70 M at V=4981), strict state_dict load, `pretrained:` of a missing file skipped."""
    from audiocaption_b200.captioning.utils import train_util
    from oracle import cnn14 as oc, crnn
    cfg = _model_cfg(yaml_name)
    msgs = []
    model = train_util.init_model_from_config(cfg, msgs.append)
    assert any("not exist" in m for m in msgs)                       # the PANNs checkpoint is not shipped
    assert type(model).__name__ == "TransformerModel"
    assert type(model.encoder).__name__ == "CrnnEncoder"
    assert type(model.encoder.cnn).__name__ == "Cnn14Encoder" and type(model.encoder.rnn).__name__ == "RnnEncoder"
    assert model.encoder.freeze_cnn_bn is True
    assert not any(p.requires_grad for p in model.encoder.cnn.parameters())
    V = cfg["decoder"]["args"]["vocab_size"]
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    n_gru = sum(p.numel() for p in model.encoder.rnn.parameters())
    n_dec = sum(p.numel() for n, p in model.decoder.named_parameters() if n != "pos_encoder.pe")
    assert n_train == n_gru + n_dec and n_gru == 5_907_456 and n_dec == 2 * 256 * V + 2_238_720
    sd = {f"encoder.cnn.{k}": v for k, v in oc.build_state_dict(3).items()}
    sd.update({f"encoder.rnn.{k}": v for k, v in crnn.build_gru_state_dict(4).items()})
    sd.update({f"decoder.{k}": v for k, v in crnn.build_decoder(6, 512, V).state_dict().items()})
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    # train() keeps the frozen CNN's BatchNorm holders in eval mode (crnn_trm_encoder.py:195-203)
    model.train()
    assert model.encoder.rnn.training and model.decoder.training
    assert all(not m.training for m in model.encoder.cnn.modules() if "BatchNorm" in type(m).__name__)
    assert model.encoder.cnn.training


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_reference_factory_builds_b200_mirrors_from_the_reference_yaml(alias):
    """The reference's OWN init_model_from_config (train_util.py:83-94), loaded from its file, run on the reference's
    own YAML after install_as_captioning(): every `type:` resolves to a B200 mirror and the key set equals the one the
    reference's classes produce."""
    ref_import.install_stubs()
    path = os.path.join(ref_import.REF_ROOT, "captioning", "utils", "train_util.py")
    spec = importlib.util.spec_from_file_location("_reference_train_util", path)
    ref_tu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_tu)
    cfg_file = os.path.join(ref_import.REF_ROOT, "eg_configs", "audiocaps", "waveform", "cnn14rnn_trm.yaml")
    with open(cfg_file) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)["model"]
    with open(os.path.join(CFG_DIR, "audiocaps_cnn14rnn_trm.yaml")) as f:
        assert yaml.safe_load(f)["model"] == cfg, "tests/golden/eg_configs copy drifted from the reference YAML"
    model = ref_tu.init_model_from_config(cfg, lambda *_: None)
    assert type(model).__module__ == "audiocaption_b200.captioning.models.transformer_model"
    ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    # the reference's classes on the same config (imported under their own package name, in a clean module table)
    _drop_captioning_modules()
    ref_model = ref_import.load("captioning.utils.train_util").init_model_from_config(cfg, lambda *_: None)
    theirs = {k: tuple(v.shape) for k, v in ref_model.state_dict().items()}
    assert ours == theirs
    model.load_state_dict(ref_model.state_dict(), strict=True)


# ------------------------------------------------------------------ CIDEr (captioning/metrics/cider.py; run.py:150-155 monitors it)
def test_cider_properties():
    """pycocoevalcap is absent (parity unpinned): pin what the published definition implies."""
    from audiocaption_b200.captioning.metrics.cider import Cider, simple_tokenize
    assert simple_tokenize("A dog's bark -- loud, isn't it? (Yes) ...") == "a dog's bark loud isn't it yes"
    gts = {"a": ["a dog barks loudly at night", "a dog is barking"], "b": ["rain falls on a tin roof", "heavy rain is falling"],
           "c": ["a man speaks to a crowd", "someone is talking"], "d": ["birds sing in the morning"]}
    exact = {k: [v[0]] for k, v in gts.items()}
    score, per = Cider().compute_score(gts, exact)
    assert Cider().method() == "CIDEr" and per.shape == (4,) and abs(score - per.mean()) < 1e-12
    # known answer by hand: a 3-word candidate equal to one of its two references has cosine 1 for n = 1..3 and no 4-grams,
    # so mean over n = 3/4, over the two references 3/8, x 10 = 3.75 (the other reference shares nothing with it)
    kgts = {"p": ["a man speaks", "someone is talking"], "q": ["rain falls hard"], "r": ["birds sing loudly"]}
    kres = {"p": ["a man speaks"], "q": ["rain falls hard"], "r": ["birds sing loudly"]}
    _, kper = Cider().compute_score(kgts, kres)
    assert abs(kper[0] - 3.75) < 1e-9 and abs(kper[1] - 7.5) < 1e-9, kper
    # a candidate equal to its only reference: cosine 1 for every n-gram order, no length penalty -> 10
    assert abs(per[3] - 10.0) < 1e-9
    # equal to one of two references: half of that reference's weight plus whatever the other shares (here nothing)
    assert 4.9 < per[0] <= 10.0 and 4.9 < per[1] <= 10.0
    # unrelated candidates score 0; permuting the keys permutes the scores
    wrong = {"a": exact["b"], "b": exact["c"], "c": exact["d"], "d": exact["a"]}
    s_wrong, per_wrong = Cider().compute_score(gts, wrong)
    assert s_wrong < 0.5 * score and (per_wrong <= per + 1e-12).all()
    rev = dict(reversed(list(gts.items())))
    _, per_rev = Cider().compute_score(rev, {k: exact[k] for k in rev})
    assert np.allclose(per_rev[::-1], per)
    # the Gaussian length penalty: repeating the caption keeps every cosine but changes the (bigram) length
    longer = {"d": [exact["d"][0] + " " + exact["d"][0]]}
    one = {"d": gts["d"], "a": gts["a"]}
    _, p_same = Cider().compute_score(one, {"d": exact["d"], "a": exact["a"]})
    _, p_long = Cider().compute_score(one, {"d": longer["d"], "a": exact["a"]})
    assert p_long[0] < p_same[0]
    # words that occur in every reference set carry no weight (idf 0): a candidate made of them scores 0
    common = {"x": ["the the the"], "y": ["the the the"]}
    assert Cider().compute_score(common, {"x": ["the the the"], "y": ["the the the"]})[0] == 0.0
    with pytest.raises(ValueError):
        Cider().compute_score(gts, {"a": ["x"]})
