"""Real-weights import: Keras `metaseg.h5` -> the `.npz` layout of ecseg_b200.weights.

The reference loads its U-Net with tf.keras.models.load_model('models/metaseg.h5') (src/utils.py:27-33); the file is
a Mendeley download that is not available offline, and its exact architecture is only known through the topology
template src/model_layers/models.py:17-136 (SURVEY.md finding 0.2).  This module therefore

  1. DISCOVERS the architecture from the checkpoint's own `model_config` JSON (Functional / Model / Sequential
     graphs: Conv2D, Conv2DTranspose, BatchNormalization, MaxPooling2D, Concatenate, Activation / ReLU / Softmax,
     Dropout), walking the graph from the input and matching it against ecseg_b200.spec.UNET_LAYERS -- channel
     counts, kernel sizes, strides, where the pools / skips / transposed convolutions sit, which convolutions carry a
     ReLU -- and reports the first deviation in words (ArchitectureMismatch) instead of importing garbage;
  2. MAPS the Keras weights (Conv2D kernel (kh,kw,Cin,Cout), Conv2DTranspose kernel (kh,kw,Cout,Cin), bias,
     BatchNorm gamma/beta/moving_mean/moving_variance) onto the spec's layer names, un-swapping the input-channel
     halves when a Concatenate lists [up, skip] instead of the template's [skip, up] (models.py:100,112,124);
  3. reads them from an HDF5 checkpoint through h5py WHEN IT IS INSTALLED (it is not in the offline image: the
     converter is meant to run once on any machine that has it), or from a `model.to_json()` + `np.savez(*model.
     get_weights())` pair that needs nothing but numpy.

    python -m ecseg_b200.keras_import models/metaseg.h5 models/metaseg.npz
    python -m ecseg_b200.keras_import --json model.json --weights weights.npz models/metaseg.npz
"""
from __future__ import annotations

import json
import sys

import numpy as np

from . import spec, weights as wmod


class ArchitectureMismatch(ValueError):
    """The checkpoint's graph is not the metaseg U-Net this framework implements."""


_IGNORED = {"Dropout", "SpatialDropout2D", "GaussianNoise", "Lambda_identity"}


def _layers_of(model_config) -> list:
    cfg = json.loads(model_config) if isinstance(model_config, (str, bytes)) else model_config
    if "class_name" not in cfg or "config" not in cfg:
        raise ArchitectureMismatch("model_config has no class_name/config: not a Keras model description")
    body = cfg["config"]
    layers = body["layers"] if isinstance(body, dict) else body      # old Sequential configs are a bare list
    seq = cfg["class_name"] == "Sequential"
    out = []
    prev = None
    for l in layers:
        name = l.get("name") or l["config"]["name"]
        if seq:
            inbound = [prev] if prev is not None else []
        else:
            nodes = l.get("inbound_nodes") or []
            inbound = []
            if nodes:
                node = nodes[0]
                if isinstance(node, dict):            # Keras 3 style {"args": [...]} is not what TF 2.8 wrote
                    raise ArchitectureMismatch("Keras 3 model_config is not supported; re-save with TF 2.x")
                for ref in node:
                    inbound.append(ref[0])
        out.append({"name": name, "cls": l["class_name"], "cfg": l["config"], "in": inbound})
        prev = name
    return out


def _pair(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v)


def discover(model_config) -> list:
    """model_config (dict or JSON text) -> 23 records in spec.UNET_LAYERS order:
    {"spec": name, "keras": conv layer name, "bn": BatchNormalization layer name or None, "bn_cfg": its config,
     "use_bias": bool, "swap_halves": bool (the layer reads a Concatenate([up, skip]))}."""
    layers = _layers_of(model_config)
    by_name = {l["name"]: l for l in layers}
    # tensor state per layer output: channels, level, and for concat bookkeeping the producing chain
    state = {}
    convs = []          # conv-like layers in graph order with what follows them
    pending = {}        # layer name -> index into convs of the conv whose BN / activation may still follow

    def src(l):
        if len(l["in"]) != 1:
            raise ArchitectureMismatch(f"layer {l['name']} ({l['cls']}) has {len(l['in'])} inputs, expected 1")
        return l["in"][0]

    for l in layers:
        cls, cfg, name = l["cls"], l["cfg"], l["name"]
        if cls == "InputLayer":
            shp = cfg.get("batch_input_shape") or cfg.get("batch_shape")
            state[name] = {"ch": shp[-1] if shp else None, "level": 0, "conv": None, "origin": "input"}
            continue
        if not l["in"]:
            raise ArchitectureMismatch(f"layer {name} ({cls}) is not connected to anything")
        if cls in ("Conv2D", "Conv2DTranspose"):
            s = state[src(l)]
            ks, st = _pair(cfg["kernel_size"]), _pair(cfg.get("strides", 1))
            if ks != (3, 3):
                raise ArchitectureMismatch(f"{name}: kernel_size {ks}, the metaseg U-Net is 3x3 throughout")
            if cfg.get("padding", "valid") != "same":
                raise ArchitectureMismatch(f"{name}: padding '{cfg.get('padding')}', expected 'same'")
            if _pair(cfg.get("dilation_rate", 1)) != (1, 1):
                raise ArchitectureMismatch(f"{name}: dilated convolutions are not part of the metaseg U-Net")
            want_st = (2, 2) if cls == "Conv2DTranspose" else (1, 1)
            if st != want_st:
                raise ArchitectureMismatch(f"{name}: strides {st}, expected {want_st}")
            act = cfg.get("activation") or "linear"
            rec = {"keras": name, "kind": "convT" if cls == "Conv2DTranspose" else "conv", "cin": s["ch"],
                   "cout": int(cfg["filters"]), "use_bias": bool(cfg.get("use_bias", True)), "act": act, "bn": None,
                   "bn_cfg": None, "level": s["level"] - (1 if cls == "Conv2DTranspose" else 0),
                   "swap_halves": bool(s.get("swapped")), "in_origin": s.get("origin")}
            if act not in ("linear", "relu", "softmax"):
                raise ArchitectureMismatch(f"{name}: activation '{act}' is not used by the metaseg U-Net")
            convs.append(rec)
            state[name] = {"ch": rec["cout"], "level": rec["level"], "conv": len(convs) - 1, "origin": name,
                           "open": act == "linear"}
        elif cls == "BatchNormalization":
            s = dict(state[src(l)])
            ci = s.get("conv")
            if ci is None or not s.get("open") or convs[ci]["bn"] is not None:
                raise ArchitectureMismatch(f"{name}: BatchNormalization must sit between a convolution and its activation "
                                           "(post-activation normalisation cannot be folded into that convolution)")
            ax = cfg.get("axis", -1)
            ax = ax[0] if isinstance(ax, (list, tuple)) else ax
            if ax not in (-1, 3):
                raise ArchitectureMismatch(f"{name}: BatchNormalization over axis {ax}, expected the channel axis")
            convs[ci]["bn"] = name
            convs[ci]["bn_cfg"] = cfg
            state[name] = s
        elif cls in ("Activation", "ReLU", "Softmax", "LeakyReLU"):
            s = dict(state[src(l)])
            kind = {"ReLU": "relu", "Softmax": "softmax"}.get(cls) or cfg.get("activation")
            if cls == "LeakyReLU" or kind not in ("relu", "softmax", "linear"):
                raise ArchitectureMismatch(f"{name}: activation '{kind or cls}' is not used by the metaseg U-Net")
            if cls == "ReLU" and (cfg.get("max_value") is not None or cfg.get("negative_slope", 0) or cfg.get("threshold", 0)):
                raise ArchitectureMismatch(f"{name}: only the plain ReLU is supported")
            ci = s.get("conv")
            if kind != "linear":
                if ci is None or not s.get("open"):
                    raise ArchitectureMismatch(f"{name}: activation without a preceding convolution")
                convs[ci]["act"] = kind
                s["open"] = False
            state[name] = s
        elif cls == "MaxPooling2D":
            s = dict(state[src(l)])
            if _pair(cfg.get("pool_size", 2)) != (2, 2) or _pair(cfg.get("strides") or cfg.get("pool_size", 2)) != (2, 2):
                raise ArchitectureMismatch(f"{name}: expected a 2x2 / stride-2 max pool")
            s.update(level=s["level"] + 1, conv=None, origin="pool:" + str(s.get("origin")))
            state[name] = s
        elif cls == "Concatenate":
            if len(l["in"]) != 2:
                raise ArchitectureMismatch(f"{name}: Concatenate of {len(l['in'])} tensors, expected [skip, up]")
            ax = cfg.get("axis", -1)
            if ax not in (-1, 3):
                raise ArchitectureMismatch(f"{name}: Concatenate along axis {ax}, expected the channel axis")
            a, b = state[l["in"][0]], state[l["in"][1]]
            if a["level"] != b["level"]:
                raise ArchitectureMismatch(f"{name}: concatenated tensors live on different resolutions")
            ka = convs[a["conv"]]["kind"] if a.get("conv") is not None else None
            kb = convs[b["conv"]]["kind"] if b.get("conv") is not None else None
            if (ka == "convT") == (kb == "convT"):
                raise ArchitectureMismatch(f"{name}: expected one encoder skip and one transposed convolution")
            swapped = ka == "convT"                       # [up, skip] instead of the template's [skip, up]
            skip, up = (b, a) if swapped else (a, b)
            state[name] = {"ch": a["ch"] + b["ch"], "level": a["level"], "conv": None, "origin": "cat",
                           "swapped": swapped, "halves": (skip["ch"], up["ch"]), "skip_origin": skip.get("origin")}
        elif cls in _IGNORED:
            state[name] = dict(state[src(l)])
        else:
            raise ArchitectureMismatch(f"layer {name}: class {cls} is not part of the metaseg U-Net")

    if len(convs) != len(spec.UNET_LAYERS):
        raise ArchitectureMismatch(f"{len(convs)} convolution layers, the metaseg U-Net has {len(spec.UNET_LAYERS)}")
    out = []
    skip_src = {"conv3-3": "conv3-2", "conv2-3": "conv2-2", "conv1-3": "conv1-2"}
    spec_to_keras = {}
    for rec, (sname, kind, cin, cout, relu, bias, level) in zip(convs, spec.UNET_LAYERS):
        where = f"{rec['keras']} (position of {sname})"
        if rec["kind"] != kind:
            raise ArchitectureMismatch(f"{where}: is a {rec['kind']}, expected a {kind}")
        if rec["cin"] is not None and rec["cin"] != cin:
            hint = " (a level-4 skip concat? the template leaves it out, models.py:87)" if sname == "conv4-3" else ""
            hint = " (RGB input? metaseg feeds one channel, utils.py:113)" if sname == "conv1-1" else hint
            raise ArchitectureMismatch(f"{where}: {rec['cin']} input channels, expected {cin}{hint}")
        if rec["cout"] != cout:
            raise ArchitectureMismatch(f"{where}: {rec['cout']} filters, expected {cout}")
        if rec["level"] != level:
            raise ArchitectureMismatch(f"{where}: runs at 1/{2 ** rec['level']} resolution, expected 1/{2 ** level}")
        want_act = "softmax" if sname == "final" else ("relu" if relu else "linear")
        got_act = rec["act"]
        if sname == "final" and got_act == "linear":
            got_act = "softmax"        # logits head: the softmax is applied by the framework (utils.py:115 reads probabilities)
        if got_act != want_act:
            raise ArchitectureMismatch(f"{where}: activation '{rec['act']}', expected '{want_act}'")
        if sname in skip_src:          # the skip must come from the encoder layer the template names
            cat = state[by_name[rec["keras"]]["in"][0]]
            if cat.get("origin") != "cat" or spec_to_keras.get(skip_src[sname]) != cat.get("skip_origin"):
                raise ArchitectureMismatch(f"{where}: expected the concat of {skip_src[sname]} and the up-sampled tensor")
        elif state[by_name[rec["keras"]]["in"][0]].get("origin") == "cat":
            raise ArchitectureMismatch(f"{where}: reads a concatenation the metaseg U-Net does not have")
        if rec["bn"] and kind == "convT":
            raise ArchitectureMismatch(f"{where}: BatchNormalization on a transposed convolution is not supported")
        spec_to_keras[sname] = rec["keras"]
        out.append({"spec": sname, "keras": rec["keras"], "bn": rec["bn"], "bn_cfg": rec["bn_cfg"],
                    "use_bias": rec["use_bias"], "swap_halves": rec["swap_halves"],
                    "halves": state[by_name[rec["keras"]]["in"][0]].get("halves")})
    return out


def map_weights(records: list, get) -> dict:
    """`get(keras_layer_name) -> [arrays in Keras order]`; returns the dict layout of ecseg_b200.weights."""
    w = {}
    for rec, (sname, kind, cin, cout, _relu, _bias, _level) in zip(records, spec.UNET_LAYERS):
        arrs = [np.asarray(a) for a in get(rec["keras"])]
        want = (3, 3, cin, cout) if kind == "conv" else (3, 3, cout, cin)
        if not arrs or arrs[0].shape != want:
            raise ArchitectureMismatch(f"{rec['keras']}: kernel shape {arrs[0].shape if arrs else None}, expected {want}")
        k = np.array(arrs[0], np.float32)
        if rec["swap_halves"]:
            n_skip, n_up = rec["halves"]
            k = np.concatenate([k[:, :, n_up:n_up + n_skip, :], k[:, :, :n_up, :]], axis=2)   # [up, skip] -> [skip, up]
        w[f"{sname}/kernel"] = np.ascontiguousarray(k)
        if rec["use_bias"]:
            if len(arrs) < 2 or arrs[1].shape != (cout,):
                raise ArchitectureMismatch(f"{rec['keras']}: bias missing or of the wrong shape")
            w[f"{sname}/bias"] = np.array(arrs[1], np.float32)
        elif sname != "final":
            w[f"{sname}/bias"] = np.zeros(cout, np.float32)
        if sname == "final" and rec["use_bias"]:
            raise ArchitectureMismatch("final: the head has no bias in the template (models.py:134); a biased head would "
                                       "need the constant-channel trick of ecseg_b200.weights and is not imported silently")
        if rec["bn"]:
            cfg = rec["bn_cfg"] or {}
            eps = float(cfg.get("epsilon", spec.BN_EPS))
            b = [np.asarray(a, np.float32) for a in get(rec["bn"])]
            it = iter(b)
            gamma = next(it) if cfg.get("scale", True) else np.ones(cout, np.float32)
            beta = next(it) if cfg.get("center", True) else np.zeros(cout, np.float32)
            mean, var = next(it), next(it)
            for a in (gamma, beta, mean, var):
                if a.shape != (cout,):
                    raise ArchitectureMismatch(f"{rec['bn']}: BatchNormalization vector of shape {a.shape}, expected ({cout},)")
            # the library folds with the Keras default epsilon: re-express other epsilons through the variance
            w[f"{sname}/bn_gamma"], w[f"{sname}/bn_beta"], w[f"{sname}/bn_mean"] = gamma, beta, mean
            w[f"{sname}/bn_var"] = (var + np.float32(eps - spec.BN_EPS)).astype(np.float32)
    return w


def from_weight_list(model_config, arrays: list) -> dict:
    """model.to_json() + model.get_weights(): the flat list follows the model's layer order, each layer contributing
    kernel[, bias] or gamma, beta, moving_mean, moving_variance."""
    records = discover(model_config)
    order = [l["name"] for l in _layers_of(model_config)]
    n_of = {}
    for r in records:
        n_of[r["keras"]] = 2 if r["use_bias"] else 1
        if r["bn"]:
            cfg = r["bn_cfg"] or {}
            n_of[r["bn"]] = 2 + int(bool(cfg.get("scale", True))) + int(bool(cfg.get("center", True)))
    table, pos = {}, 0
    for name in order:
        if name in n_of:
            table[name] = arrays[pos:pos + n_of[name]]
            pos += n_of[name]
    if pos != len(arrays):
        raise ArchitectureMismatch(f"{len(arrays)} weight arrays, the discovered graph accounts for {pos}")
    return map_weights(records, lambda n: table[n])


def load_keras_h5(path: str, h5py_module=None) -> dict:
    """Read a Keras HDF5 checkpoint (model.save('x.h5')): attrs['model_config'] + the model_weights group."""
    if h5py_module is None:
        try:
            import h5py as h5py_module
        except ImportError as e:
            raise ImportError("h5py is needed to read a Keras .h5 checkpoint; run this converter once on a machine that has it, "
                              "or export `model.to_json()` and `np.savez('w.npz', *model.get_weights())` and use "
                              "--json / --weights") from e
    with h5py_module.File(path, "r") as f:
        cfg = f.attrs.get("model_config")
        if cfg is None:
            raise ArchitectureMismatch(f"{path}: no model_config attribute (weights-only file): pass the JSON with --json")
        cfg = cfg.decode() if isinstance(cfg, bytes) else cfg
        records = discover(cfg)
        g = f["model_weights"] if "model_weights" in f else f

        def get(layer):
            grp = g[layer]
            names = [n.decode() if isinstance(n, bytes) else n for n in grp.attrs["weight_names"]]
            return [np.asarray(grp[n][()]) for n in names]

        return map_weights(records, get)


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="Keras metaseg checkpoint -> ecseg_b200 .npz weights")
    ap.add_argument("src", nargs="?", help="metaseg.h5 (needs h5py)")
    ap.add_argument("dst", help="output .npz (default location: models/metaseg.npz)")
    ap.add_argument("--json", help="model.to_json() text file")
    ap.add_argument("--weights", help="np.savez(*model.get_weights()) file")
    a = ap.parse_args(argv)
    try:
        if a.json and a.weights:
            with np.load(a.weights) as z:
                arrays = [z[k] for k in sorted(z.files, key=lambda s: int(s.split("_")[-1]))]
            w = from_weight_list(open(a.json).read(), arrays)
        elif a.src:
            w = load_keras_h5(a.src)
        else:
            ap.error("give metaseg.h5, or --json and --weights")
    except ArchitectureMismatch as e:
        print("architecture mismatch:", e, file=sys.stderr)
        return 3
    wmod.pack_blob(w)            # shape check of everything against the spec
    wmod.save_npz(a.dst, w)
    print(f"wrote {a.dst}: {sum(v.size for v in w.values())} parameters, batch norm: {wmod.has_bn(w)}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
