"""Live-TensorFlow pin of the floating-point oracle.

The network / loss / optimizer oracle (oracle/network_oracle.py, oracle/heatmap_oracle.py) restates
model/hourglass.py:5-206, loss.py:2-36 and Keras' legacy Adam (trainer.py:31).  TensorFlow is not installable in the
build container, so there these tests SKIP and the restatement stays "parity unpinned"; on any machine where
`import tensorflow` works and the reference tree is present (HGB_REFERENCE_DIR, default /root/reference) they load
ONE set of numpy weights into the real Keras model and into the restatement, by layer name, and compare:

  * architecture: layer names, shapes, creation order, parameter counts (model/hourglass.py:5-32)
  * forward heat maps, training-mode and inference-mode BatchNorm, fp32 (<= 1e-4 relative)
  * the four losses on identical tensors (loss.py:2-36 + tf.keras.losses.mean_squared_error) and the Keras reduction
  * parameter gradients of one weighted-MSE step (cosine > 0.9999 per tensor) and two Keras-Adam updates

CPU only (no `gpu` marker): it pins the ORACLE; the CUDA path is pinned against the oracle by the -m gpu tests.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("HGB_REFERENCE_DIR", "/root/reference")
_HAVE_TF = importlib.util.find_spec("tensorflow") is not None
_HAVE_REF = os.path.isfile(os.path.join(REF, "model", "hourglass.py"))

pytestmark = pytest.mark.skipif(not (_HAVE_TF and _HAVE_REF),
                                reason="needs an importable tensorflow and the reference tree (parity stays unpinned here)")


def _load_reference():
    """Import the reference's own model/hourglass.py and loss.py (unmodified, from where they lie)."""
    import tensorflow as tf
    tf.config.set_visible_devices([], "GPU")            # the pin is an fp32 CPU comparison
    saved = list(sys.path)
    sys.path.insert(0, REF)
    try:
        for name in [n for n in sys.modules if n == "model" or n.startswith("model.") or n == "loss"]:
            del sys.modules[name]
        spec = importlib.util.spec_from_file_location("ref_hourglass", os.path.join(REF, "model", "hourglass.py"))
        hg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(hg)
        spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(REF, "loss.py"))
        loss = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(loss)
    finally:
        sys.path[:] = saved
    return tf, hg, loss


def _keras_weight_names(model):
    """[(oracle name, keras variable)] in Keras creation order: '<layer>/<kernel|bias|gamma|...>'."""
    out = []
    for layer in model.layers:
        for v in layer.weights:
            leaf = v.name.split("/")[-1].split(":")[0]
            out.append((f"{layer.name}/{leaf}", v))
    return out


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def ref():
    return _load_reference()


@pytest.mark.parametrize("stacks", [1, 2])
def test_architecture_names_shapes_counts(ref, stacks):
    from oracle import network_oracle as norc
    tf, hg, _loss = ref
    model = hg.create_hourglass_model(17, stacks, 256, (256, 256, 3), "sigmoid")
    spec = norc.param_spec(17, stacks, 256)
    named = _keras_weight_names(model)
    assert {n for n, _ in named} == set(spec), "layer/variable names differ from the reference"
    for n, v in named:
        assert tuple(v.shape) == tuple(spec[n]), n
    assert model.count_params() == norc.count_params(spec)[0]
    # creation order of the trainable kernels (what Keras checkpoint layer numbering follows)
    assert [n for n, _ in named if n.endswith("/kernel")] == [n for n in spec if n.endswith("/kernel")]


def test_forward_loss_gradients_and_adam_match_live_tf(ref):
    from oracle import heatmap_oracle as horc
    from oracle import network_oracle as norc
    tf, hg, loss = ref
    S, B = 2, 2
    model = hg.create_hourglass_model(17, S, 256, (256, 256, 3), "sigmoid")
    weights = norc.init_params(norc.param_spec(17, S, 256), seed=2, perturb_bn=True)
    named = _keras_weight_names(model)
    for n, v in named:
        v.assign(weights[n])
    rng = np.random.default_rng(0)
    images = rng.random((B, 256, 256, 3), dtype=np.float32)
    kx = rng.uniform(-4, 68, (B, 17)).astype(np.float32)
    ky = rng.uniform(-4, 68, (B, 17)).astype(np.float32)
    kv = rng.integers(0, 3, (B, 17))
    targets = horc.render_targets(kx, ky, kv, 64, 64)

    # ---- forward, both BatchNorm modes
    for training in (False, True):
        got = model(images, training=training)
        want, _ = norc.forward(weights, images, 17, S, 256, training=training)
        for s in range(S):
            g, w = np.asarray(got[s]), want[s].detach().numpy()
            assert np.abs(g - w).max() <= 1e-4 * max(np.abs(w).max(), 1e-12) + 1e-6, (training, s)
        if training:      # Keras updated its moving statistics in place: put the shared values back
            for n, v in named:
                v.assign(weights[n])

    # ---- the four losses + Keras reduction on identical tensors
    pred = rng.random(targets.shape, dtype=np.float32)
    np.testing.assert_allclose(np.asarray(loss.weighted_mse(targets, pred)), horc.weighted_mse_map(targets, pred), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(np.asarray(loss.weighed_keypoint_mse(targets, pred)), horc.keypoint_mse_map(targets, pred), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(np.asarray(loss.IOU(targets, pred)), horc.iou_vec(targets, pred), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(np.asarray(tf.keras.losses.mean_squared_error(targets, pred)), horc.mse_map(targets, pred), rtol=1e-5, atol=1e-8)
    for kind, fn in (("weighted_mse", loss.weighted_mse), ("iou", loss.IOU), ("weighted_keypoint_mse", loss.weighed_keypoint_mse)):
        ref_scalar = float(tf.reduce_mean(fn(targets, pred)))
        assert abs(horc.loss_and_grad(kind, targets, pred)[0] - ref_scalar) <= 1e-5 * abs(ref_scalar)

    # ---- one training step: summed loss (trainer.py:35 compile semantics), gradients, two Adam updates (trainer.py:31)
    train_vars = [v for _n, v in named if v.trainable]
    train_names = [n for n, v in named if v.trainable]
    with tf.GradientTape() as tape:
        outs = model(images, training=True)
        per = [tf.reduce_mean(loss.weighted_mse(targets, o)) for o in outs]
        total = tf.add_n(per)
    tf_grads = tape.gradient(total, train_vars)
    _o, o_losses, o_grads = norc.loss_and_grads(weights, images, targets, "weighted_mse", 17, S, 256)
    for s in range(S):
        assert abs(float(per[s]) - o_losses[s]) <= 1e-4 * abs(o_losses[s])
    worst = min(_cos(np.asarray(g), o_grads[n]) for n, g in zip(train_names, tf_grads))
    assert worst > 0.9999, f"lowest per-tensor gradient cosine vs live TF: {worst}"

    try:
        opt = tf.keras.optimizers.legacy.Adam(learning_rate=1e-3)      # what trainer.py:31 resolved to in TF 2.8-2.10
    except AttributeError:
        opt = tf.keras.optimizers.Adam(learning_rate=1e-3)
    probe = [i for i, n in enumerate(train_names) if n in ("front_conv_1x1_1/kernel", "hg0_conv_1x1_predict/bias",
                                                           "batch_normalization_20/gamma")]
    state = {i: (weights[train_names[i]].copy(), np.zeros_like(weights[train_names[i]]), np.zeros_like(weights[train_names[i]]))
             for i in probe}
    for t in (1, 2):
        opt.apply_gradients(zip(tf_grads, train_vars))
        for i in probe:
            w, m, v = state[i]
            norc.adam_step(w, o_grads[train_names[i]], m, v, t)
    for i in probe:
        np.testing.assert_allclose(train_vars[i].numpy(), state[i][0], rtol=1e-4, atol=1e-6)
