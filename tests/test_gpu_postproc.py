"""GPU parity: CCL / morphology / rule kernels vs the golden vectors and the CPU oracle.
Bit-exact (integer work).  All calls go through the C ABI (ecseg_b200.engine -> ctypes)."""
import warnings

import numpy as np
import pytest
from scipy import ndimage as ndi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ecseg_b200.engine import Engine
    e = Engine(0, 2048, 2048, max_tiles=0)
    yield e
    e.close()


def _oracle():
    from oracle import metaseg_oracle as mo
    return mo


def test_golden_full_pipeline(eng, golden):
    g = golden("postproc")
    for i in range(int(g["n_cases"])):
        for faithful in (False, True):
            out, n, px = eng.postprocess(g[f"in_{i}"], faithful_merge=faithful)
            assert np.array_equal(out.cpu().numpy(), g[f"out_{i}"]), (i, faithful)
            assert (n, px) == tuple(int(v) for v in g[f"cnt_{i}"]), (i, faithful)


def test_golden_single_steps(eng, golden):
    g = golden("postproc")
    for i in range(int(g["n_cases"])):
        m = g[f"in_{i}"]
        assert np.array_equal(eng.fill_holes(m, 1).cpu().numpy(), g[f"fill1_{i}"]), i
        assert np.array_equal(eng.fill_holes(m, 2).cpu().numpy(), g[f"fill2_{i}"]), i
        assert np.array_equal(eng.size_thresh(m).cpu().numpy(), g[f"size_{i}"]), i
        assert np.array_equal(eng.merge_comp(m, 1).cpu().numpy(), g[f"merge1_{i}"]), i
        assert np.array_equal(eng.merge_comp(m, 2).cpu().numpy(), g[f"merge2_{i}"]), i
        assert eng.count_cc(m == 3) == tuple(int(v) for v in g[f"cnt_in_{i}"]), i


@pytest.mark.parametrize("conn", [4, 8])
def test_label_partition_and_root_is_first_pixel(eng, conn):
    from ecseg_b200 import synth
    mo = _oracle()
    st = ndi.generate_binary_structure(2, 1 if conn == 4 else 2)
    for seed, shape in [(0, (97, 131)), (1, (256, 256)), (2, (300, 1000)), (3, (513, 33))]:
        m = synth.synth_noise_label_map(seed, *shape, block=1 + seed % 3)
        got = eng.label(m, conn).cpu().numpy()
        ref = np.zeros(m.shape, np.int64)
        off = 0
        for v in (1, 2, 3):
            l, k = ndi.label(m == v, structure=st)
            ref[l > 0] = l[l > 0] + off
            off += k
        assert mo.labels_equal_up_to_permutation(got, ref), (seed, conn)
        fg = got > 0
        first = np.full(int(got.max()) + 1, got.size, np.int64)
        np.minimum.at(first, got.ravel(), np.arange(got.size))
        assert np.array_equal(first[got[fg]] + 1, got[fg]), "label must be 1 + index of the first pixel"


def test_full_size_vs_oracle(eng):
    """BASELINE config 4 shape: 2048x2048 synthetic label maps, bit-exact map + count."""
    from ecseg_b200 import synth
    mo = _oracle()
    for seed in (0, 1, 2):
        m = synth.synth_label_map(seed, 2048, 2048)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = mo.meta_inference(m.astype(np.int64).copy())
        out, n, px = eng.postprocess(m)
        assert np.array_equal(out.cpu().numpy(), want), seed
        assert (n, px) == mo.count_cc(want == 3), seed


def test_adversarial_noise_and_ragged_shapes(eng):
    from ecseg_b200 import synth
    mo = _oracle()
    for seed, shape, block in [(7, (1040, 1392), 1), (8, (777, 1291), 2), (9, (1, 500), 1), (10, (500, 1), 1),
                               (11, (31, 33), 1), (12, (1024, 1024), 5)]:
        m = synth.synth_noise_label_map(seed, *shape, block=block)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = mo.meta_inference(m.astype(np.int64).copy())
        out, n, px = eng.postprocess(m, faithful_merge=bool(seed % 2))
        assert np.array_equal(out.cpu().numpy(), want), (seed, shape)
        assert (n, px) == mo.count_cc(want == 3), (seed, shape)


def test_postprocess_idempotent_count(eng):
    """Size-independent property: counting the ecDNA of the final map again gives the same tuple."""
    from ecseg_b200 import synth
    m = synth.synth_label_map(5, 2048, 2048)
    out, n, px = eng.postprocess(m)
    assert eng.count_cc(out == 3) == (n, px)
    assert int((out == 3).sum().item()) == px


def test_graph_replay_same_buffers_different_maps(eng, golden):
    """On a non-default stream ecseg_postprocess captures its launch sequence once per argument set and replays it
    (postproc.cu pp_postprocess).  Same device buffers, changing contents and both labelling parities: every replay
    must equal the oracle bit for bit, and account for the same number of launches (21) as the captured sequence."""
    import torch
    from ctypes import c_void_p
    from ecseg_b200 import synth
    mo = _oracle()
    dev = eng.device
    s = torch.cuda.Stream(device=dev)
    h, w = 300, 420
    buf = torch.empty((h, w), dtype=torch.uint8, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    px = torch.zeros(1, dtype=torch.int64, device=dev)
    per_call = []
    for seed in range(7):
        m = synth.synth_label_map(40 + seed, h, w) if seed % 2 else synth.synth_noise_label_map(40 + seed, h, w, block=2)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = mo.meta_inference(m.astype(np.int64).copy())
        src = torch.from_numpy(m).to(dev)
        l0 = eng.launch_count()
        with torch.cuda.stream(s):
            buf.copy_(src, non_blocking=True)
            eng._chk(eng.lib.ecseg_postprocess(eng.ctx, buf.data_ptr(), h, w, 0, n.data_ptr(), px.data_ptr(), c_void_p(s.cuda_stream)))
        s.synchronize()
        per_call.append(eng.launch_count() - l0)
        assert np.array_equal(buf.cpu().numpy(), want), seed
        assert (int(n.item()), int(px.item())) == mo.count_cc(want == 3), seed
    assert len(set(per_call)) == 1 and per_call[0] >= 15, per_call       # replays account for the same launches
    # golden cases through the graph path too (fresh buffers: captured, launched once, evicted from the bounded cache)
    g = golden("postproc")
    with torch.cuda.stream(s):
        for i in range(int(g["n_cases"])):
            out, cnt, cpx = eng.postprocess(g[f"in_{i}"], faithful_merge=bool(i % 2))
            assert np.array_equal(out.cpu().numpy(), g[f"out_{i}"]), i
            assert (cnt, cpx) == tuple(int(v) for v in g[f"cnt_{i}"]), i


def test_config4_maps_vs_reference_run(eng, golden):
    """BASELINE config 4 on the maps bench.py cycles: the 64 synthetic 2048x2048 label maps (seeds 0..63) against what
    the reference's own meta_inference + count_cc returned for them (tests/golden/config4.npz, frozen by
    oracle/make_golden.py gen_config4): count tuple, class histogram and SHA-256 of the final map, all 64 bit-exact."""
    import hashlib
    import torch
    from ecseg_b200 import synth
    g = golden("config4")
    n_maps = len(g["cnt"])
    assert n_maps >= 64
    s = torch.cuda.Stream(device=eng.device)
    with torch.cuda.stream(s):          # non-default stream: the graph-replay path the bench uses
        for seed in range(n_maps):
            out, n, px = eng.postprocess(synth.synth_label_map(seed, 2048, 2048))
            o = out.cpu().numpy()
            assert (n, px) == tuple(int(v) for v in g["cnt"][seed]), seed
            assert np.array_equal(np.bincount(o.ravel(), minlength=4), g["hist"][seed]), seed
            assert hashlib.sha256(o.tobytes()).digest() == g["sha"][seed].tobytes(), seed
