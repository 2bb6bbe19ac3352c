"""GPU parity: U-Net forward (fp32 parity mode and tcgen05 fp16/bf16 modes) vs the torch-CPU
oracle on identical weights.  Floating point: tolerances from BASELINE.json's north_star --
fp32 logits within 1e-3 relative, label agreement >= 99.9 % excluding quantised top-2 ties."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_REL_TOL_FP32 = 1e-3
LABEL_AGREEMENT = 0.999
BF16_FLOOR = 0.995
# GPU against the CPU restatement of its own arithmetic (UNetOracle16), rms error relative to the layer's scale, for the
# first tensor-core layer (fused conv1-1 + conv1-2), where the two can only differ by rounding flips caused by the
# order of the fp32 accumulation: measured 3.0e-5 in fp16 (a tenth of the distance to the fp32 oracle).  Further down
# the rounding decisions of the two decorrelate (9e-5, 1.9e-4, ... 7e-4) and the general bars apply.
BAR16_FIRST = {"fp16": 1e-4, "bf16": 4e-4}        # measured 3.0e-5 and 1.06e-4


@pytest.fixture(scope="module")
def setup():
    from ecseg_b200 import synth, weights as wmod
    from ecseg_b200.engine import Engine
    from oracle import metaseg_oracle as mo
    from oracle.unet_oracle import UNetOracle
    w = wmod.make_weights(0)
    img = synth.synth_dapi(31, 462, 470)
    pre = mo.meta_preprocess(img)
    pos, tiles = mo.im2patches_overlap(pre[..., None])
    oracle = UNetOracle(w, batch=3)
    z_ref = oracle.predict_logits(tiles)
    p_ref = torch.softmax(torch.from_numpy(z_ref), -1).numpy()
    eng = Engine(0, 512, 512)
    yield dict(w=w, img=img, pre=pre, pos=pos, tiles=tiles, z_ref=z_ref, p_ref=p_ref, eng=eng, mo=mo)
    eng.close()


def _agreement(p_gpu, p_ref):
    q_ref = np.clip(np.rint(p_ref.astype(np.float64) * 255), 0, 255)
    q_gpu = np.clip(np.rint(p_gpu.astype(np.float64) * 255), 0, 255)
    srt = np.sort(q_ref, -1)
    notie = srt[..., 3] != srt[..., 2]
    return float((np.argmax(q_gpu, -1) == np.argmax(q_ref, -1))[notie].mean()), float(1 - notie.mean())


def test_fp32_parity_mode(setup):
    s = setup
    eng = s["eng"]
    eng.load_weights(s["w"], "fp32")
    probs, logits = eng.unet_forward(s["tiles"][..., 0], want_logits=True)
    z = logits.cpu().numpy()
    rel = np.abs(z - s["z_ref"]).max() / np.abs(s["z_ref"]).max()
    assert rel <= LOGIT_REL_TOL_FP32, rel
    agree, ties = _agreement(probs.cpu().numpy(), s["p_ref"])
    assert agree >= LABEL_AGREEMENT, (agree, ties)
    assert np.abs(probs.cpu().numpy() - s["p_ref"]).max() < 1e-3


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_tensor_core_modes(setup, prec):
    s = setup
    eng = s["eng"]
    eng.load_weights(s["w"], prec)
    probs, logits = eng.unet_forward(s["tiles"][..., 0], want_logits=True)
    assert eng.device_error() == 0
    z = logits.cpu().numpy()
    rel = np.abs(z - s["z_ref"]).max() / np.abs(s["z_ref"]).max()
    agree, ties = _agreement(probs.cpu().numpy(), s["p_ref"])
    print(f"{prec}: logits max rel err {rel:.3e}, label agreement {agree * 100:.4f}% (ties {ties * 100:.3f}%)")
    assert eng.activation_overflow() == -1
    assert rel <= (6e-2 if prec == "bf16" else 1e-2), rel
    # fp16 is the throughput dtype and carries the 99.9 % bar.  bf16 (8-bit mantissa) cannot reach it on this network
    # (SURVEY section 7 probe: 99.87 %, measured here ~99.8 %): it stays selectable for checkpoints whose activations
    # leave the fp16 range, with a regression floor here and the bar itself recorded as an expected failure below.
    assert agree >= (BF16_FLOOR if prec == "bf16" else LABEL_AGREEMENT), (agree, ties)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_every_layer_against_the_oracle(setup, prec):
    """Per-layer parity of the tensor-core modes.  With random-init weights the label map is decided by levels 0-2 and
    their skip connections -- replacing the whole output of up3 by zeros still leaves 99.90 % of the labels and moves
    the logits by 2.7e-3 (profiles/r02_probe_numerics.txt) -- so the label and logit bars above say little about the
    deep layers.  Here every layer's activation (stop_after: the activation buffers are reused further down the
    network) is compared with the fp32 oracle's: maximum and rms error relative to the layer's own scale.  Bars = 3x
    / 2x what 16-bit operands with fp32 accumulation cost in an emulation on the CPU (fp16: max <= 1.4e-3, rms
    <= 9.2e-4; bf16: 1.1e-2, 7.9e-3); a wrong block, tap or channel chunk in any layer is an O(1) error."""
    from ecseg_b200 import spec
    from oracle.unet_oracle import UNetOracle, UNetOracle16
    s = setup
    eng = s["eng"]
    tiles = s["tiles"]
    n = len(tiles)
    x = torch.from_numpy(np.ascontiguousarray(tiles)).float().permute(0, 3, 1, 2)
    taps, taps16 = {}, {}
    with torch.no_grad():
        UNetOracle(s["w"], batch=n).logits(x, taps)
        UNetOracle16(s["w"], torch.float16 if prec == "fp16" else torch.bfloat16).logits(x, taps16)
    bar_max, bar_rms = (4e-3, 2e-3) if prec == "fp16" else (3e-2, 1.6e-2)
    eng.load_weights(s["w"], prec)
    worst, worst16 = [0.0, 0.0], [0.0, 0.0]

    def errs(got, ref):
        return (float(np.abs(got - ref).max() / np.abs(ref).max()),
                float(np.sqrt(np.mean((got - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))))

    try:
        for li in range(1, 22):                  # conv1-1 is fused into conv1-2 (its activation never exists); 22 = logits
            name = spec.UNET_LAYERS[li][0]
            eng.debug_set(stop_after=li)
            eng.unet_forward(tiles[..., 0])
            got = eng.layer_output(li, n).cpu().numpy()
            ref = taps[name].permute(0, 2, 3, 1).numpy()
            assert got.shape == ref.shape
            e_max, e_rms = errs(got, ref)
            assert e_max <= bar_max and e_rms <= bar_rms, (name, e_max, e_rms)
            # against the same arithmetic restated on the CPU: only the fp32 accumulation order differs, which shows
            # where a stored activation sits on a rounding boundary of the 16-bit format
            f_max, f_rms = errs(got, taps16[name].permute(0, 2, 3, 1).numpy())
            assert f_max <= bar_max and f_rms <= (BAR16_FIRST[prec] if li == 1 else bar_rms), (name, f_max, f_rms)
            worst = [max(worst[0], e_max), max(worst[1], e_rms)]
            worst16 = [max(worst16[0], f_max), max(worst16[1], f_rms)]
            print(f"  {name:8s} vs fp32 oracle {e_max:.2e} / {e_rms:.2e}   vs 16-bit-arithmetic oracle {f_max:.2e} / {f_rms:.2e}")
        assert eng.device_error() == 0
    finally:
        eng.debug_set(stop_after=-1)
    print(f"{prec}: worst layer against the fp32 oracle: max err {worst[0]:.3e}, rms err {worst[1]:.3e}; against the "
          f"16-bit-arithmetic oracle: max err {worst16[0]:.3e}, rms err {worst16[1]:.3e} (relative to the layer's scale)")
    _probs, logits = eng.unet_forward(tiles[..., 0], want_logits=True)
    z16 = taps16["final"].permute(0, 2, 3, 1).numpy()
    l_max, l_rms = errs(logits.cpu().numpy(), z16)
    print(f"{prec}: logits against the 16-bit-arithmetic oracle: max err {l_max:.3e}, rms err {l_rms:.3e}")
    assert l_max <= 2 * bar_max and l_rms <= 2 * bar_rms, (l_max, l_rms)


@pytest.mark.xfail(reason="bf16 operands (8-bit mantissa) measure ~99.8 % < the 99.9 % label bar; fp16 -- same tcgen05 "
                          "kind::f16 rate -- is the throughput dtype (DESIGN.md section 2)", strict=False)
def test_bf16_at_the_label_bar(setup):
    s = setup
    eng = s["eng"]
    eng.load_weights(s["w"], "bf16")
    agree, _ = _agreement(eng.unet_forward(s["tiles"][..., 0]).cpu().numpy(), s["p_ref"])
    print(f"bf16 label agreement {agree * 100:.4f}%")
    assert agree >= LABEL_AGREEMENT, agree


def test_fp16_range_guard_fires_on_overflow(setup):
    """With a real checkpoint nothing bounds the activations; an fp16 overflow in an epilogue conversion would become
    inf -> NaN logits -> label 0 silently.  Scaled-up weights must raise ECSEG_E_RANGE (ValueError) and name the
    layer, in the staged and in the whole-image call; bf16 and fp32 run the same weights without complaint; conv1-1
    (bounded input) is rejected at load time."""
    import copy
    s = setup
    eng = s["eng"]
    big = copy.copy(s["w"])
    big["conv1-2/kernel"] = np.asarray(s["w"]["conv1-2/kernel"]) * np.float32(3e4)
    eng.load_weights(big, "fp16")
    eng.unet_forward(s["tiles"][..., 0])
    with pytest.raises(ValueError, match="layer 1 "):
        eng.activation_overflow()
    with pytest.raises(ValueError, match="16-bit operand range"):
        eng.segment_host(s["img"])
    eng.load_weights(big, "bf16")
    eng.unet_forward(s["tiles"][..., 0])
    assert eng.activation_overflow() == -1
    first = copy.copy(s["w"])
    first["conv1-1/kernel"] = np.asarray(s["w"]["conv1-1/kernel"]) * np.float32(1e4)
    with pytest.raises(ValueError, match="conv1-1"):
        eng.load_weights(first, "fp16")
    eng.load_weights(first, "bf16")            # bf16 has fp32's exponent range
    eng.load_weights(s["w"], "fp16")
    eng.unet_forward(s["tiles"][..., 0])
    assert eng.activation_overflow() == -1     # the flag does not stick


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_fused_whole_image_equals_staged_and_oracle_postprocess(setup, prec):
    """ecseg_segment_image_host == staged C-ABI calls; post-processing + count bit-exact vs the
    oracle given the same (GPU) label map; conv1-1 reading through the tile grid == materialised tiles."""
    s = setup
    eng, mo = s["eng"], s["mo"]
    eng.load_weights(s["w"], prec)
    h, w = s["img"].shape
    dapi = np.empty((h, w), np.uint8)
    labels, n_ec, ec_px = eng.segment_host(s["img"], dapi_out=dapi)
    assert np.array_equal(dapi, 255 - s["pre"])
    pre, _ = eng.preprocess(s["img"])
    raw = eng.stitch_argmax(eng.unet_forward(eng.tile(pre)), h, w).cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = mo.meta_inference(raw.astype(np.int64).copy())
    assert np.array_equal(labels, want)
    assert (n_ec, ec_px) == mo.count_cc(want == 3)


def test_batch_larger_than_one_image_and_determinism(setup):
    s = setup
    eng = s["eng"]
    eng.load_weights(s["w"], "fp16")
    a = eng.unet_forward(s["tiles"][..., 0]).cpu().numpy()
    b = eng.unet_forward(s["tiles"][..., 0]).cpu().numpy()
    assert np.array_equal(a, b)
    one = eng.unet_forward(s["tiles"][2:3, ..., 0]).cpu().numpy()
    assert np.array_equal(one[0], a[2])      # tiles are independent: batch composition must not matter


def test_fused_first_layer_pair_equals_single_cta(monkeypatch):
    """The CTA-pair variant of the fused conv1-1 -> conv1-2 kernel hands its generated halo stages to the MMA issuer
    with plain remote mbarrier arrives (no cluster-scope release); the single-CTA variant has no cross-CTA step and
    the same arithmetic.  Bit-identical outputs over a whole image's 100 tiles, repeated, is the race check."""
    from ecseg_b200 import weights as wmod
    from ecseg_b200.engine import Engine
    rng = np.random.default_rng(5)
    tiles = rng.integers(0, 256, (100, 256, 256), dtype=np.uint8)
    tiles[::7] = 0                      # some all-zero tiles: any stale data would show
    eng = Engine(0, 2048, 2048)
    try:
        eng.load_weights(wmod.make_weights(0), "fp16")
        dev = eng._dev(tiles, torch.uint8)

        def run():
            probs = eng.unet_forward(dev)
            return eng.layer_output(1, 100).clone(), eng.layer_output(2, 100).clone(), probs

        monkeypatch.setenv("ECSEG_FUSE1_SINGLE", "1")
        ref = run()
        monkeypatch.delenv("ECSEG_FUSE1_SINGLE")
        for _ in range(4):
            got = run()
            assert eng.device_error() == 0
            for a, b in zip(got, ref):
                assert torch.equal(a, b)
    finally:
        eng.close()


@pytest.mark.parametrize("shape", [(256, 256), (256, 700), (700, 512), (300, 300), (513, 777), (462, 1100), (1040, 600)])
def test_owned_block_skipping_equals_full_computation(shape):
    """The whole-image call computes only the blocks of the last four layers that the stitcher can take from a tile
    (work lists, unet.cu); the staged calls compute every tile in full.  Same labels, bit for bit, on tile grids with a
    single row / column, pulled-back last tiles and tiles that own only a sliver."""
    from ecseg_b200 import synth, weights as wmod
    from ecseg_b200.engine import Engine, unet_work
    h, w = shape
    img = synth.synth_dapi(90 + h % 7 + w % 5, h, w)
    eng = Engine(0, max(h, 256), max(w, 256))
    try:
        eng.load_weights(wmod.make_weights(0), "fp16")
        other = synth.synth_dapi(7 + h % 5 + w % 3, h, w)
        for rep in range(2):
            if rep:      # second pass: every activation buffer holds ANOTHER image's values in full (first pass: zeros),
                eng.unet_forward(eng.tile(eng.preprocess(other)[0]))     # so a block skipped by mistake reads wrong data
            labels, n_ec, ec_px = eng.segment_host(img)
            pre, _ = eng.preprocess(img)
            raw = eng.stitch_argmax(eng.unet_forward(eng.tile(pre)), h, w)
            want, n2, px2 = eng.postprocess(raw)
            assert np.array_equal(labels, want.cpu().numpy()), (shape, rep)
            assert (n_ec, ec_px) == (n2, px2)
            assert eng.activation_overflow() == -1
        ref, ex = unet_work(h, w)
        assert ex <= ref and (ex < ref or shape == (256, 256))
    finally:
        eng.close()
