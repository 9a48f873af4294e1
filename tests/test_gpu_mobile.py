"""mobile=True: the SeparableConv2D bottleneck variant (model/hourglass.py:9-11, 209-231) end to end against the oracle's
restatement of it (oracle/network_oracle.py, `mobile=True`): depthwise k x k stencil kernels + the pointwise tcgen05 GEMMs.
Per-op parity (depthwise forward / input gradient / weight gradient vs fp32 torch, cosine > 0.999) is in
test_gpu_ops_replay.py; here: the loss gate (2e-2) at BASELINE config 1's shape, the backward wiring in the tame regime,
inference with moving statistics, the factory switch and an optimizer run."""
import numpy as np
import pytest

from oracle import network_oracle as norc
from tests.test_gpu_network import _inputs, _rel, _run_train_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hgb():
    import hgb200
    return hgb200


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def test_mobile_config1_loss_gate(hgb, torch):
    """1-stack, 256 ch, batch 8, weighted_MSE (BASELINE config 1's shape) with separable bottlenecks: loss within 2e-2."""
    _run_train_case(hgb, torch, S=1, B=8, kind="weighted_mse", perturb=False, mobile=True)


def test_mobile_two_stack_backward_wiring_in_tame_regime(hgb, torch):
    """Every parameter tensor's gradient (depthwise kernels included) tracks the bf16-emulating oracle tensor by tensor."""
    _run_train_case(hgb, torch, S=2, B=2, kind="weighted_mse", perturb=True, tame=True, loss_tol=4e-2, mobile=True)


def test_mobile_inference_matches_oracle(hgb, torch):
    images, _ = _inputs(3)
    weights = norc.init_params(norc.param_spec(17, 2, 256, mobile=True), seed=5, perturb_bn=True)
    model = hgb.create_hourglass_model(17, 2, 256, (256, 256, 3), "sigmoid", mobile=True)
    model.set_weights_dict(weights)
    got = model.predict(images, batch_size=2)
    ref, _ = norc.forward(weights, images, 17, 2, 256, training=False, emulate_bf16=True, mobile=True)
    ref32, _ = norc.forward(weights, images, 17, 2, 256, training=False, mobile=True)
    assert len(got) == 2 and got[0].shape == (3, 64, 64, 17)
    for s in range(2):
        r32, e32 = ref32[s].detach().numpy(), ref[s].detach().numpy()
        d = float(np.linalg.norm(got[s] - r32) / np.linalg.norm(r32))
        e = float(np.linalg.norm(e32 - r32) / np.linalg.norm(r32))
        print(f"mobile inference stack {s}: rel-L2 from fp32 oracle: CUDA {d:.4g}, bf16-emulating oracle {e:.4g}; "
              f"max-rel CUDA {_rel(got[s], r32):.4g}")
        assert d <= 2.0 * e + 2e-2


def test_mobile_training_reduces_loss_and_roundtrips_checkpoint(hgb, torch, tmp_path):
    images, targets = _inputs(4)
    model = hgb.HourglassModel(17, 2, 256, (256, 256, 3), "sigmoid", seed=3, mobile=True)
    model.compile(optimizer=hgb.Adam(1e-3), loss=hgb.loss.weighted_mse)
    first = model.train_on_batch(images, targets)
    for _ in range(6):
        last = model.train_on_batch(images, targets)
    assert last[0] < first[0]
    before = model.get_weights_dict()
    assert not np.array_equal(before["hg0_downsample_f1_conv_3x3_2/depthwise_kernel"],
                              hgb.HourglassModel(17, 2, 256, (256, 256, 3), "sigmoid", seed=3, mobile=True)
                              .get_weights_dict()["hg0_downsample_f1_conv_3x3_2/depthwise_kernel"])     # Adam moved it
    model.save_weights(str(tmp_path / "m.ckpt"))
    other = hgb.HourglassModel(17, 2, 256, (256, 256, 3), "sigmoid", seed=9, mobile=True)
    other.load_weights(str(tmp_path / "m.ckpt"))
    a, b = model.predict(images[:2]), other.predict(images[:2])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
