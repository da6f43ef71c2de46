"""Mirror of the model-construction half of captioning/utils/train_util.py: the reflection factory the YAML configs go
through (:63-94), the config loader with single-parent `inherit_from` and CLI overrides (:112-151), the merge-load of
pretrained weights (:188-223), `set_seed` (:225-230) and the SWA `AveragedModel` (:233-253).

Pure host logic (no kernels).  Dataset / HDF5 / logging helpers of the reference file are out of scope."""
import importlib
import os
import random
import sys
from typing import Callable, Dict, Union

import numpy as np
import torch
import yaml
from torch.optim.swa_utils import AveragedModel as _TorchAveragedModel


def pad_sequence(data, pad_value=0):
    """List of 1-D (or [T, D]) arrays -> (padded tensor [N, max_len, ...], lengths) (train_util.py:24-31)."""
    if isinstance(data[0], (np.ndarray, torch.Tensor)):
        data = [torch.as_tensor(arr) for arr in data]
    padded = torch.nn.utils.rnn.pad_sequence(data, batch_first=True, padding_value=pad_value)
    return padded, np.array([x.shape[0] for x in data])


def get_cls_from_str(string, reload=False):
    """'pkg.module.Class' -> the class object (train_util.py:63-68)."""
    module_name, cls_name = string.rsplit(".", 1)
    module = importlib.import_module(module_name)
    if reload:
        module = importlib.reload(module)
    return getattr(module, cls_name)


def init_obj_from_dict(config, **kwargs):
    """Instantiate config["type"](**config["args"], **kwargs); nested dict entries other than type/args are built
    recursively unless the caller already supplied them (train_util.py:70-81)."""
    args = dict(config["args"])
    args.update(kwargs)
    for key, sub in config.items():
        if key in ("type", "args") or key in kwargs or not isinstance(sub, dict):
            continue
        args[key] = init_obj_from_dict(sub)
    try:
        return get_cls_from_str(config["type"])(**args)
    except Exception:
        print(f"Initializing {config} failed, detailed error stack: ")
        raise


def init_model_from_config(config, print_fn=sys.stdout.write):
    """Depth-first construction of a model tree: every sub-key that is not type/args/pretrained is a sub-model built
    first (and merge-loaded from its own `pretrained:` entry), then passed to the parent constructor under the key's
    name (train_util.py:83-94)."""
    subs = {}
    for key, sub in config.items():
        if key in ("type", "args", "pretrained"):
            continue
        sub_model = init_model_from_config(sub, print_fn)
        if "pretrained" in sub:
            load_pretrained_model(sub_model, sub["pretrained"], print_fn)
        subs[key] = sub_model
    return init_obj_from_dict(config, **subs)


def merge_a_into_b(a, b):
    """Deep merge: values of `a` override `b` (train_util.py:112-120)."""
    for key, value in a.items():
        if isinstance(value, dict) and key in b:
            assert isinstance(b[key], dict), "Cannot inherit key '{}' from base!".format(key)
            merge_a_into_b(value, b[key])
        else:
            b[key] = value


def load_config(config_file):
    """YAML with an optional `inherit_from:` path relative to the file itself (train_util.py:122-136)."""
    with open(config_file, "r") as reader:
        config = yaml.load(reader, Loader=yaml.FullLoader)
    if "inherit_from" not in config:
        return config
    base_file = os.path.join(os.path.dirname(config_file), config.pop("inherit_from"))
    assert not os.path.samefile(config_file, base_file), "inherit from itself"
    base = load_config(base_file)
    merge_a_into_b(config, base)
    return base


def parse_config_or_kwargs(config_file, **kwargs):
    """CLI `--a.b=v` overrides merged over the YAML (train_util.py:138-151 goes through TOML to get nested keys; the
    dotted keys are split here directly, same result)."""
    config = load_config(config_file)
    override: Dict = {}
    for dotted, value in kwargs.items():
        node = override
        *parents, leaf = dotted.split(".")
        for part in parents:
            node = node.setdefault(part, {})
        node[leaf] = value
    merge_a_into_b(override, config)
    return config


def store_yaml(config, config_file):
    with open(config_file, "w") as writer:
        yaml.dump(config, writer, indent=4, default_flow_style=False)


def merge_load_state_dict(state_dict, model: torch.nn.Module, output_fn: Callable = sys.stdout.write):
    """Load the entries whose key AND shape match, keep the rest (train_util.py:188-202).  Returns the loaded keys."""
    own = model.state_dict()
    loaded, mismatch = {}, []
    for key, value in state_dict.items():
        if key in own and own[key].shape == value.shape:
            loaded[key] = value
        else:
            mismatch.append(key)
    output_fn(f"Loading pre-trained model, with mismatched keys {mismatch}\n")
    own.update(loaded)
    model.load_state_dict(own, strict=True)
    return loaded.keys()


def load_pretrained_model(model: torch.nn.Module, pretrained: Union[str, Dict],
                          output_fn: Callable = sys.stdout.write):
    """train_util.py:204-223: a missing file is reported and skipped; a model with its own `load_pretrained` gets the
    path; otherwise merge-load the checkpoint's `model` entry."""
    if not isinstance(pretrained, dict) and not os.path.exists(pretrained):
        output_fn(f"pretrained {pretrained} not exist!")
        return
    if hasattr(model, "load_pretrained"):
        model.load_pretrained(pretrained, output_fn)
        return
    state_dict = pretrained if isinstance(pretrained, dict) else torch.load(pretrained, map_location="cpu")
    if "model" in state_dict:
        state_dict = state_dict["model"]
    merge_load_state_dict(state_dict, model, output_fn)


def set_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class AveragedModel(_TorchAveragedModel):
    """SWA copy that averages the buffers as well as the parameters (train_util.py:233-253)."""

    def update_parameters(self, model):
        first = self.n_averaged == 0
        pairs = list(zip(self.parameters(), model.parameters())) + \
            list(zip(list(self.buffers())[1:], model.buffers()))          # buffers()[0] is n_averaged itself
        for avg, cur in pairs:
            cur = cur.detach().to(avg.device)
            if first or not cur.is_floating_point():
                avg.detach().copy_(cur)
            else:                                     # the default avg_fn of swa_utils: equal-weight running mean
                n = self.n_averaged.to(avg.device)
                avg.detach().copy_(avg.detach() + (cur - avg.detach()) / (n + 1))
        self.n_averaged += 1
