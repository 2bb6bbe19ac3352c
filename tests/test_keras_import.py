"""CPU-only: Keras checkpoint import.  The checkpoint's model_config is generated here in
the schema TF 2.8 writes (class_name / config / inbound_nodes) for the topology of the reference template
src/model_layers/models.py:17-136; weights are the seeded random-init set, so a round trip must be exact."""
import json

import numpy as np
import pytest

from ecseg_b200 import keras_import as ki, spec, weights as wmod


def keras_model(w, with_bn=True, act_in_conv=False, swap_concat=False, dropout=False, level4_skip=False, in_ch=1,
                leaky=False, bn_eps=1e-3):
    """(model_config dict, get_weights() list) of a Functional model with the metaseg topology."""
    layers, arrays = [], []
    n = [0]

    def add(cls, cfg, inbound, name=None):
        n[0] += 1
        name = name or f"{cls.lower()}_{n[0]}"
        cfg = dict(cfg, name=name)
        layers.append({"class_name": cls, "name": name, "config": cfg,
                       "inbound_nodes": [[[i, 0, 0, {}] for i in inbound]] if inbound else []})
        return name

    x = add("InputLayer", {"batch_input_shape": [None, 256, 256, in_ch], "dtype": "float32"}, [], "input_1")
    skips = {}
    for sname, kind, cin, cout, relu, bias, level in spec.UNET_LAYERS:
        if sname in ("conv2-1", "conv3-1", "conv4-1", "conv5-1"):
            x = add("MaxPooling2D", {"pool_size": [2, 2], "strides": [2, 2], "padding": "same"}, [x])
        if sname in ("conv3-3", "conv2-3", "conv1-3") or (sname == "conv4-3" and level4_skip):
            sk = skips[{"conv4-3": "conv4-2", "conv3-3": "conv3-2", "conv2-3": "conv2-2", "conv1-3": "conv1-2"}[sname]]
            x = add("Concatenate", {"axis": -1}, [x, sk] if swap_concat else [sk, x])
            if dropout:
                x = add("Dropout", {"rate": 0.5}, [x])
        bn = with_bn and f"{sname}/bn_gamma" in w
        final = sname == "final"
        inline = ("relu" if relu else "linear") if (act_in_conv and not bn) else "linear"
        if final and act_in_conv:
            inline = "softmax"
        kcin = cin if not (sname == "conv4-3" and level4_skip) else 2 * cin
        x = add("Conv2DTranspose" if kind == "convT" else "Conv2D",
                {"filters": cout, "kernel_size": [3, 3], "strides": [2, 2] if kind == "convT" else [1, 1], "padding": "same",
                 "dilation_rate": [1, 1], "activation": inline, "use_bias": bool(bias)}, [x],
                name=sname if kind == "conv" else None)
        k = w[f"{sname}/kernel"]
        if kcin != cin:
            k = np.concatenate([k, k], axis=2)
        if swap_concat and sname in ("conv3-3", "conv2-3", "conv1-3"):
            h = k.shape[2] // 2
            k = np.concatenate([k[:, :, h:], k[:, :, :h]], axis=2)        # the model was trained on [up, skip]
        arrays.append(k)
        if bias:
            arrays.append(w[f"{sname}/bias"])
        if bn:
            x = add("BatchNormalization", {"axis": [3], "epsilon": bn_eps, "center": True, "scale": True}, [x])
            var = w[f"{sname}/bn_var"] - np.float32(bn_eps - 1e-3)
            arrays += [w[f"{sname}/bn_gamma"], w[f"{sname}/bn_beta"], w[f"{sname}/bn_mean"], var.astype(np.float32)]
        if relu and inline == "linear":
            x = add("LeakyReLU", {"alpha": 0.1}, [x]) if leaky else add("Activation", {"activation": "relu"}, [x])
        if final and inline != "softmax":
            x = add("Softmax", {"axis": -1}, [x])
        if sname in ("conv1-2", "conv2-2", "conv3-2", "conv4-2"):
            skips[sname] = x
    cfg = {"class_name": "Functional", "config": {"name": "model", "layers": layers, "input_layers": [["input_1", 0, 0]],
                                                  "output_layers": [[x, 0, 0]]}, "keras_version": "2.8.0", "backend": "tensorflow"}
    return cfg, arrays


def same(a, b):
    assert set(a) == set(b), set(a) ^ set(b)
    for k in a:
        assert a[k].dtype == np.float32 and np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("with_bn", [True, False])
@pytest.mark.parametrize("act_in_conv,swap_concat,dropout", [(False, False, False), (True, False, True), (False, True, False)])
def test_round_trip_exact(with_bn, act_in_conv, swap_concat, dropout):
    w = wmod.make_weights(0, with_bn=with_bn)
    cfg, arrays = keras_model(w, with_bn, act_in_conv, swap_concat, dropout)
    recs = ki.discover(json.dumps(cfg))
    assert [r["spec"] for r in recs] == [l[0] for l in spec.UNET_LAYERS]
    assert [r["swap_halves"] for r in recs if r["spec"] in ("conv3-3", "conv2-3", "conv1-3")] == [swap_concat] * 3
    got = ki.from_weight_list(cfg, arrays)
    same(got, w)
    assert np.array_equal(wmod.pack_blob(got), wmod.pack_blob(w))        # what ecseg_load_weights receives


def test_other_bn_epsilon_is_re_expressed_through_the_variance():
    w = wmod.make_weights(0, with_bn=True)
    cfg, arrays = keras_model(w, bn_eps=1e-5)
    got = ki.from_weight_list(cfg, arrays)
    for k in w:
        assert np.allclose(got[k], w[k], rtol=0, atol=1e-6), k


@pytest.mark.parametrize("kw,needle", [
    ({"level4_skip": True}, "level-4 skip"),          # the classic U-Net the template deliberately is not (models.py:87)
    ({"in_ch": 3}, "RGB input"),
    ({"leaky": True}, "not used by the metaseg U-Net"),
])
def test_deviations_are_named(kw, needle):
    w = wmod.make_weights(0, with_bn=False)
    cfg, arrays = keras_model(w, with_bn=False, **kw)
    with pytest.raises(ki.ArchitectureMismatch) as e:
        ki.from_weight_list(cfg, arrays)
    assert needle in str(e.value)


def test_wrong_kernel_shape_and_count():
    w = wmod.make_weights(0, with_bn=False)
    cfg, arrays = keras_model(w, with_bn=False)
    bad = list(arrays)
    bad[2] = bad[2][:, :, :, :32]
    with pytest.raises(ki.ArchitectureMismatch):
        ki.from_weight_list(cfg, bad)
    with pytest.raises(ki.ArchitectureMismatch):
        ki.from_weight_list(cfg, arrays[:-1])


class _FakeH5:
    """The slice of the h5py API load_keras_h5 uses, over an in-memory Keras-style layout."""

    class Node(dict):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.attrs = {}

    class DS:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, key):
            assert key == ()
            return self.a

    def __init__(self, cfg, arrays, records, weights_only=False):
        self.root = self.Node()
        if not weights_only:
            self.root.attrs["model_config"] = json.dumps(cfg).encode()
        mw = self.root["model_weights"] = self.Node()
        it = iter(arrays)
        order = [l["name"] for l in cfg["config"]["layers"]]
        n_of = {}
        for r in records:
            n_of[r["keras"]] = 2 if r["use_bias"] else 1
            if r["bn"]:
                n_of[r["bn"]] = 4
        for name in order:
            g = mw[name] = self.Node()
            names = [f"{name}/w{i}:0".encode() for i in range(n_of.get(name, 0))]
            g.attrs["weight_names"] = names
            for nm in names:
                g[nm.decode()] = self.DS(next(it))

    def File(self, path, mode):
        root = self.root

        class Ctx:
            def __enter__(self_inner):
                return root

            def __exit__(self_inner, *a):
                return False

        return Ctx()


def test_h5_layout_through_an_injected_h5py(tmp_path):
    w = wmod.make_weights(0, with_bn=True)
    cfg, arrays = keras_model(w)
    fake = _FakeH5(cfg, arrays, ki.discover(cfg))
    same(ki.load_keras_h5("metaseg.h5", h5py_module=fake), w)
    with pytest.raises(ki.ArchitectureMismatch):
        ki.load_keras_h5("weights_only.h5", h5py_module=_FakeH5(cfg, arrays, ki.discover(cfg), weights_only=True))
    # without h5py the failure is loud and says what to do instead
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError) as e:
            ki.load_keras_h5("metaseg.h5")
        assert "--json" in str(e.value)


def test_cli_json_plus_weights(tmp_path):
    w = wmod.make_weights(0, with_bn=True)
    cfg, arrays = keras_model(w)
    (tmp_path / "m.json").write_text(json.dumps(cfg))
    np.savez(tmp_path / "w.npz", *arrays)
    out = tmp_path / "metaseg.npz"
    assert ki.main(["--json", str(tmp_path / "m.json"), "--weights", str(tmp_path / "w.npz"), str(out)]) == 0
    same(wmod.load_npz(str(out)), w)
    cfg3, arrays3 = keras_model(wmod.make_weights(0, with_bn=False), with_bn=False, in_ch=3)
    (tmp_path / "m3.json").write_text(json.dumps(cfg3))
    np.savez(tmp_path / "w3.npz", *arrays3)
    assert ki.main(["--json", str(tmp_path / "m3.json"), "--weights", str(tmp_path / "w3.npz"), str(out)]) == 3
