"""Config / model factory API -- mirror of the pieces of opencood/hypes_yaml/yaml_utils.py:14-58, 346-379 and
opencood/tools/train_utils.py:171-291 that the inference path uses: ``load_yaml``, ``create_model``,
``load_saved_model``, ``to_device``."""
from __future__ import annotations

import glob
import importlib
import math
import os
import re

import torch
import yaml


def _loader():
    loader = yaml.Loader
    # same float resolver as the reference so that values like 1e-3 parse as floats
    loader.add_implicit_resolver(
        "tag:yaml.org,2002:float",
        re.compile("""^(?:[-+]?(?:[0-9][0-9_]*)\\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                      |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                      |\\.[0-9_]+(?:[eE][-+][0-9]+)?
                      |[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+\\.[0-9_]*
                      |[-+]?\\.(?:inf|Inf|INF)
                      |\\.(?:nan|NaN|NAN))$""", re.X), list("-+0123456789."))
    return loader


def load_general_params(param):
    """Adds the anchor grid extents W / H / D and the voxel grid to the config (reference yaml_utils.py:346-379)."""
    rng = param["preprocess"]["cav_lidar_range"]
    vs = param["preprocess"]["args"]["voxel_size"]
    aa = param["postprocess"]["anchor_args"]
    aa["vw"], aa["vh"], aa["vd"] = vs[0], vs[1], vs[2]
    aa["W"] = math.ceil((rng[3] - rng[0]) / vs[0])
    aa["H"] = math.ceil((rng[4] - rng[1]) / vs[1])
    aa["D"] = math.ceil((rng[5] - rng[2]) / vs[2])
    param["postprocess"].update({"anchor_args": aa})
    return param


def load_yaml(file, opt=None):
    if opt is not None and getattr(opt, "model_dir", None):
        file = os.path.join(opt.model_dir, "config.yaml")
    with open(file, "r") as f:
        param = yaml.load(f, Loader=_loader())
    if "yaml_parser" in param:
        param = {"load_general_params": load_general_params}[param["yaml_parser"]](param)
    return param


def default_config(fusion: str = "att") -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    sub = "Attfuse/lidar_attfuse_stage3.yaml" if fusion == "att" else "Fcooper/lidar_maxfuse_stage3.yaml"
    return os.path.join(here, "hypes_yaml", "v2x_real", "Codebook", sub)


def create_model(hypes):
    """Instantiate hypes['model']['core_method'] (snake_case file name -> CamelCase class) with its args."""
    name = hypes["model"]["core_method"]
    target = name.replace("_", "").lower()
    for mod in ("quantv2x_b200.collab_model", "quantv2x_b200.pyramid_model"):
        lib = importlib.import_module(mod)
        for attr, cls in lib.__dict__.items():
            if attr.lower() == target and isinstance(cls, type):
                return cls(hypes["model"]["args"])
    raise NotImplementedError(f"model {name!r} is not on the B200 path (available: heter_baseline_collab_codebook_mc"
                              "[_encdec], heter_pyramid_collab_codebook_mc[_encdec])")


def load_saved_model(saved_path, model):
    """Load net_epoch_bestval_at*.pth or the newest net_epoch*.pth with strict=False (reference behaviour)."""
    best = glob.glob(os.path.join(saved_path, "net_epoch_bestval_at*.pth"))
    if best:
        path, epoch = best[0], int(re.findall(r"at(\d+)", best[0])[-1])
    else:
        files = glob.glob(os.path.join(saved_path, "net_epoch*.pth"))
        if not files:
            return 0, model
        epochs = [int(re.findall(r"net_epoch(\d+)", f)[-1]) for f in files]
        epoch = max(epochs)
        path = os.path.join(saved_path, f"net_epoch{epoch}.pth")
    missing = model.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
    if missing.missing_keys:
        print("missing keys:", missing.missing_keys)
    if hasattr(model, "codebook"):
        model.codebook.reset_engine()
    return epoch, model


def to_device(inputs, device):
    if isinstance(inputs, list):
        return [to_device(x, device) for x in inputs]
    if isinstance(inputs, dict):
        return {k: to_device(v, device) for k, v in inputs.items()}
    if isinstance(inputs, torch.Tensor):
        return inputs.to(device)
    return inputs
