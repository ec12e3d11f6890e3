"""On-disk formats of the reference's drivers, so that runs started with the reference
continue here and vice versa.  Host-side only; nothing in this file touches the device.

  * checkpoint dict           main.py:32-39     `checkpoint_<fold>.pt` = {state_dict, optimizer, epoch_num, train_loss, val_loss}
  * optimizer state           main.py:251       torch.optim.Adam.state_dict() <-> engine.FlatAdam's flat moment buffers
  * initial weights           model.py:59-60    `initial_weight.pt` = net.state_dict()
  * normalisation statistics  data.py:166-173   `norm.npz` = {mean [N,3], std [N,3]}
  * training history          main.py:282-310   `history<fold>.json`
  * inference reports         inference.py:149-157   `pred.json`, `error_list.json`, `inference.json`
  * OBJ meshes                data.py:20-26 (writer), psbody / open3d readers (`v` / `f` records)
  * configuration             config_parser.py:49-92 (.cfg, RawConfigParser sections)
"""
import configparser
import json
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

SECTIONS = ("Input Output", "ChebModel  Parameters", "Learning Parameters")       # the double space is the reference's

# key -> (section index, type); list types are comma separated
_CFG: Dict[str, Tuple[int, str]] = {
    "root_dir": (0, "str"), "checkpoint_dir": (0, "str"), "template": (0, "str"), "error_file": (0, "str"),
    "log_file": (0, "str"), "type": (0, "str"), "num_classes": (0, "int"), "num_style": (0, "int"), "model": (0, "str"),
    "folds": (0, "int"), "test_size": (0, "float"), "random_seeds": (0, "int"),
    "checkpoint_file": (1, "str"), "n_layers": (1, "int"), "num_hidden": (1, "int"), "downsampling_factors": (1, "ints"),
    "num_conv_filters": (1, "ints"), "workers_thread": (1, "int"), "polygon_order": (1, "ints"),
    "optimizer": (2, "str"), "batch_size": (2, "int"), "learning_rate": (2, "float"), "learning_rates": (2, "floats"),
    "learning_rates_epochs": (2, "floats"), "learning_rate_decay": (2, "float"), "weight_decay": (2, "float"),
    "dropout": (2, "float"), "epoch": (2, "int"),
}
_DEFAULTS = {"error_file": "", "checkpoint_file": "", "log_file": "log.txt", "workers_thread": "0", "learning_rate_decay": "0.99",
             "optimizer": "adam", "model": "optimal_sigma_VAE", "folds": "5", "test_size": "0.3", "random_seeds": "2020"}


def _convert(raw: str, kind: str):
    raw = raw.strip()
    if kind == "str":
        return raw
    if kind == "int":
        return int(raw)
    if kind == "float":
        return float(raw)
    items = [t for t in (u.strip() for u in raw.split(",")) if t]
    return [int(t) for t in items] if kind == "ints" else [float(t) for t in items]


def read_config(fname: str, strict: bool = False) -> Optional[dict]:
    """The dict `config_parser.read_config` returns (config_parser.py:49-92): same keys, same types, `log_file`
    joined onto `checkpoint_dir`.  Tolerant where the reference raises: section names are matched with
    whitespace collapsed, missing optional keys take the reference's defaults (strict=True restores the raise)."""
    if not os.path.exists(fname):
        print("Config not found %s" % fname)
        return None
    cp = configparser.RawConfigParser()
    cp.read(fname)
    by_norm = {" ".join(s.split()).lower(): s for s in cp.sections()}
    out = {}
    for key, (sec, kind) in _CFG.items():
        section = by_norm.get(" ".join(SECTIONS[sec].split()).lower())
        raw = None
        if section is not None and cp.has_option(section, key):
            raw = cp.get(section, key)
        elif not strict:
            # the key may sit in another section of a hand-edited file
            for s in cp.sections():
                if cp.has_option(s, key):
                    raw = cp.get(s, key)
                    break
            if raw is None:
                raw = _DEFAULTS.get(key)
        if raw is None:
            raise KeyError(f"{fname}: missing '{key}' in section [{SECTIONS[sec]}]")
        out[key] = _convert(raw, kind)
    out["log_file"] = os.path.join(out["checkpoint_dir"], out["log_file"])
    return out


def write_config(fname: str, config: dict) -> None:
    cp = configparser.RawConfigParser()
    for s in SECTIONS:
        cp.add_section(s)
    for key, (sec, kind) in _CFG.items():
        if key not in config:
            continue
        v = config[key]
        if key == "log_file":
            v = os.path.basename(v)
        cp.set(SECTIONS[sec], key, ", ".join(str(t) for t in v) if kind in ("ints", "floats") else str(v))
    with open(fname, "w") as fp:
        cp.write(fp)


# ---- OBJ ---------------------------------------------------------------------------------------------
def save_obj(filename: str, vertices, faces) -> None:
    """`v %f %f %f` / 1-based `f %d %d %d` records, as data.py:20-26 writes them (one buffered write)."""
    v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
    f = np.asarray(faces).reshape(-1, 3).astype(np.int64) + 1
    lines = ["v %f %f %f\n" % (a, b, c) for a, b, c in v]
    lines += ["f %d %d %d\n" % (a, b, c) for a, b, c in f]
    with open(filename, "w") as fp:
        fp.write("".join(lines))


def load_obj(filename: str) -> Tuple[np.ndarray, np.ndarray]:
    """(v [N,3] float64, f [F,3] int32, 0-based) of a triangle OBJ file: `v x y z [w|r g b]` and `f a b c` with
    `a`, `a/t`, `a/t/n` or `a//n` corners and negative (relative) indices; other records are skipped - what the
    drivers need from psbody `Mesh(filename=...)` (main.py:217-219) / open3d `read_triangle_mesh` (model.py:36)."""
    verts: List[List[float]] = []
    faces: List[List[int]] = []
    with open(filename) as fp:
        for line in fp:
            if line.startswith("v "):
                t = line.split()
                verts.append([float(t[1]), float(t[2]), float(t[3])])
            elif line.startswith("f "):
                idx = []
                for corner in line.split()[1:]:
                    i = int(corner.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for j in range(1, len(idx) - 1):            # fan-triangulate polygons
                    faces.append([idx[0], idx[j], idx[j + 1]])
    return np.asarray(verts, dtype=np.float64).reshape(-1, 3), np.asarray(faces, dtype=np.int32).reshape(-1, 3)


# ---- norm.npz / history / inference reports ------------------------------------------------------------
def save_norm(checkpoint_dir: str, mean: np.ndarray, std: np.ndarray) -> str:
    path = os.path.join(checkpoint_dir, "norm")
    np.savez(path, mean=mean, std=std)                         # numpy appends .npz (data.py:170)
    return path + ".npz"


def load_norm(checkpoint_dir: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mean, std) as the FloatTensors main.py:56-58 builds"""
    d = np.load(os.path.join(checkpoint_dir, "norm.npz"), allow_pickle=True)
    return torch.FloatTensor(d["mean"]), torch.FloatTensor(d["std"])


def history_entry(epoch: int, begin: float, duration: float, train: Sequence, valid: Sequence) -> dict:
    """one element of history<fold>.json (main.py:282-304) from the return tuples of train() and evaluate()"""
    t_loss, t_kld, t_rec, t_err, t_acc = train
    v_loss, v_kld, v_rec, v_acc, v_errors, v_sex = valid
    num = lambda a: a.item() if hasattr(a, "item") else a      # noqa: E731
    return {"epoch": epoch, "begin": begin, "duration": duration,
            "training": {"loss": num(t_loss), "kld": num(t_kld), "reconstruction_loss": num(t_rec), "accuracy": num(t_acc),
                         "error": num(t_err)},
            "validation": {"loss": num(v_loss), "kld": num(v_kld), "reconstruction_loss": num(v_rec), "accuracy": num(v_acc),
                           "error": float(np.asarray(v_errors).mean()), "sex_change_success_rate": num(v_sex)}}


def save_history(checkpoint_dir: str, fold: int, history: List[dict]) -> str:
    path = os.path.join(checkpoint_dir, "history" + str(fold) + ".json")
    with open(path, "w") as fp:
        json.dump(history, fp)
    return path


def save_inference_reports(output_path: str, names: Iterable[str], sex: Iterable[int], mean_err: Iterable[float],
                           max_err: Iterable[float]) -> None:
    """pred.json / error_list.json / inference.json exactly as inference.py:77-80, 120-122, 149-157 fills them:
    `names` are the full paths the loader yields; inference.json is keyed by the base name."""
    results, pred, errors = {}, {}, {}
    for name, sx, e_mean, e_max in zip(names, sex, mean_err, max_err):
        base = name.split("/").pop()
        results[base] = {"sex": int(sx), "reconstruction_error": {"mean": float(str(np.float32(e_mean))),
                                                                   "max": float(str(np.float32(e_max)))}}
        pred[name] = str(int(sx))
        errors[name] = format(float(e_mean), ".4f")
    for fname, obj in (("pred.json", pred), ("error_list.json", errors), ("inference.json", results)):
        with open(os.path.join(output_path, fname), "w") as fp:
            json.dump(obj, fp)


# ---- checkpoints ---------------------------------------------------------------------------------------
def adam_state_dict(opt, net: torch.nn.Module) -> dict:
    """engine.FlatAdam -> the dict torch.optim.Adam(net.parameters(), ...).state_dict() would hold after the same
    steps (main.py:35): parameter indices follow net.parameters(); parameters that never received a gradient
    (dec_lin_1, quirk 7) have no state entry, as in torch."""
    index = {id(p): i for i, p in enumerate(net.parameters())}
    step = float(opt.step_count.item())
    state = {}
    for p, o in zip(opt.params, opt.offsets):
        k = p.numel()
        if step > 0:
            state[index[id(p)]] = {"step": torch.tensor(step), "exp_avg": opt.m[o:o + k].view_as(p).clone(),
                                   "exp_avg_sq": opt.v[o:o + k].view_as(p).clone()}
    group = {"lr": opt.lr, "betas": tuple(opt.betas), "eps": opt.eps, "weight_decay": opt.weight_decay, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "decoupled_weight_decay": False, "params": list(range(len(index)))}
    return {"state": dict(sorted(state.items())), "param_groups": [group]}


def load_adam_state_dict(opt, net: torch.nn.Module, sd: dict) -> None:
    """the inverse: a torch.optim.Adam state dict (checkpoint['optimizer'], main.py:242) into FlatAdam's buffers"""
    index = {id(p): i for i, p in enumerate(net.parameters())}
    g = sd["param_groups"][0]
    opt.lr, opt.betas, opt.eps, opt.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
    steps = set()
    opt.m.zero_()
    opt.v.zero_()
    for p, o in zip(opt.params, opt.offsets):
        st = sd["state"].get(index[id(p)])
        if st is None:
            continue
        k = p.numel()
        opt.m[o:o + k].copy_(st["exp_avg"].reshape(-1))
        opt.v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
        steps.add(int(float(st["step"])))
    if len(steps) > 1:
        raise ValueError(f"optimizer state with different step counts per parameter ({sorted(steps)}): not representable "
                         "with one fused step counter")
    opt.step_count.fill_(steps.pop() if steps else 0)


def save_model(net, optimizer, epoch, train_loss, val_loss, checkpoint_dir) -> str:
    """main.py:32-39.  `optimizer`: a torch optimizer or an engine.FlatAdam (stored in torch.optim.Adam's format)."""
    opt_sd = optimizer.state_dict() if hasattr(optimizer, "state_dict") else adam_state_dict(optimizer, net)
    ck = {"state_dict": {k: v.detach().clone() for k, v in net.state_dict().items()}, "optimizer": opt_sd, "epoch_num": epoch,
          "train_loss": train_loss, "val_loss": val_loss}
    path = os.path.join(checkpoint_dir, "checkpoint_" + str(epoch) + ".pt")
    torch.save(ck, path)
    return path


def load_model(net, path: str, optimizer=None, map_location=None) -> dict:
    """main.py:237-247 / inference.py:211-213: restores the weights (in place - parameters that live in an engine's
    flat buffer stay there) and, if given, the optimizer."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    net.load_state_dict(ck["state_dict"])
    if optimizer is not None and "optimizer" in ck:
        if hasattr(optimizer, "load_state_dict"):
            optimizer.load_state_dict(ck["optimizer"])
        else:
            load_adam_state_dict(optimizer, net, ck["optimizer"])
    return ck


def save_initial_weight(net, checkpoint_dir: str) -> str:
    path = os.path.join(checkpoint_dir, "initial_weight.pt")             # model.py:59-60
    torch.save(net.state_dict(), path)
    return path
