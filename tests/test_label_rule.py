"""Host-side check of the head kernel's label shortcut (ecseg_b200/csrc/stitch.cuh: label_from_logits).

The fused head turns 4 logits into the label np.argmax(img_as_ubyte(softmax(z))) of the reference (src/utils.py:117-118).
It skips the divisions and the float64 quantisation when the largest softmax term leads every other by more than
2/255 of the sum; this test restates both paths in float32 numpy and checks, on random and adversarial near-tie
logits, that the shortcut never disagrees with the full rule (it may only decline to fire)."""
import numpy as np


def _both_paths(z):
    z = z.astype(np.float32)
    m = z.max(1, keepdims=True)
    e = np.exp((z - m).astype(np.float32)).astype(np.float32)
    s = np.zeros(len(z), np.float32)
    for c in range(4):                       # the kernel's summation order
        s = (s + e[:, c]).astype(np.float32)
    best = np.argmax(e, 1)                   # first maximum, as the kernel's strict '>' scan
    top = e[np.arange(len(z)), best]
    masked = e.copy()
    masked[np.arange(len(z)), best] = 0
    second = masked.max(1)
    fires = ((top - second).astype(np.float32) * np.float32(255.0)).astype(np.float32) > (np.float32(2.0) * s).astype(np.float32)
    p = (e / s[:, None]).astype(np.float32)
    q = np.clip(np.rint(p.astype(np.float64) * 255.0), 0, 255)
    full = np.argmax(q, 1)                   # first maximum of the quantised probabilities
    return fires, best, full


def test_shortcut_agrees_with_full_rule_whenever_it_fires():
    rng = np.random.default_rng(0)
    n = 2_000_000
    z = rng.normal(0, 3, (n, 4))
    # adversarial: the two largest logits within a few quantisation steps of each other, and exact ties
    near = rng.normal(0, 3, (n, 4))
    near[:, 1] = near[:, 0] + rng.uniform(-0.05, 0.05, n)
    near[:, 2:] -= 2
    ties = rng.normal(0, 1, (1000, 4))
    ties[:, 3] = ties[:, 1]
    for block in (z, near, ties, np.zeros((4, 4))):
        fires, best, full = _both_paths(block)
        assert np.array_equal(best[fires], full[fires])
    fires, _, _ = _both_paths(z)
    assert fires.mean() > 0.9                # and it does fire for the bulk of ordinary pixels
    fires, _, _ = _both_paths(np.zeros((4, 4)))
    assert not fires.any()                   # exact ties always take the full path
