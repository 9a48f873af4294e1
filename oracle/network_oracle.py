"""fp32 CPU restatement of the reference network, losses-on-outputs and Keras Adam --
TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs).

Follows model/hourglass.py:5-231 line by line (mobile=True: the SeparableConv2D bottleneck of :209-231), with the Keras defaults the reference relies on
(SURVEY.md section 7 appendix): Conv2D use_bias, 'same' padding (TF: extra pixel after),
activation inside the conv, BatchNormalization(momentum .99, eps 1e-3) in training mode
(biased batch variance) or inference mode (moving statistics), MaxPool2D 2x2/2,
UpSampling2D nearest 2x, Add.  Parameters are a dict keyed by Keras names holding HWIO kernels,
so the same arrays can be loaded into the CUDA model by name.

Parity pin: TensorFlow is not installable here, so the numerics of this restatement are
"parity unpinned" against live TF; the ARCHITECTURE is pinned by the reference's saved
model.summary() parameter counts (3,659,665 / 7,034,530 / 13,784,260), checked in
tests/test_cpu_host.py (test_parameter_counts_match_reference_summaries); wherever TensorFlow IS importable,
tests/test_tf_crosscheck.py loads shared weights into the real model/hourglass.py + loss.py + Keras Adam
and pins this file against them.

`device`: the restatement runs on the CPU by default; the GPU parity tests of the large BASELINE
configurations (4-stack batch 64, 8-stack batch 32 / 128) pass device="cuda" so the fp32 reference
finishes in seconds -- still plain fp32 torch (TF32 disabled), still test infrastructure only.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
BN_MOMENTUM = 0.99


class _Builder:
    """Walks model/hourglass.py once; in 'spec' mode it records parameter names/shapes, in 'run' mode it computes."""

    def __init__(self, params=None, training=True, update_moving=False, emulate_bf16=False, mobile=False):
        self.mobile = mobile              # bottleneck_block_mobile instead of bottleneck_block (hourglass.py:9-11)
        self.params = params
        self.training = training
        self.update_moving = update_moving
        # emulate_bf16: round to bfloat16 at exactly the points where the CUDA pipeline stores bf16
        # (conv operands, conv outputs, BN(+residual) outputs, merges); straight-through in backward.
        self.emulate = emulate_bf16
        self.spec = OrderedDict()
        self.bn_count = 0
        self.taps = {}

    def q(self, t):
        if not self.emulate:
            return t
        return t + (t.to(torch.bfloat16).to(torch.float32) - t).detach()

    # ---- layers
    def conv(self, x, name, k, cin, cout, activation, stride=1, residual=None):
        if self.params is None:
            self.spec[name + "/kernel"] = (k, k, cin, cout)
            self.spec[name + "/bias"] = (cout,)
            return x
        w = self.q(self.params[name + "/kernel"]).permute(3, 2, 0, 1)  # HWIO -> OIHW
        b = self.params[name + "/bias"]
        if stride == 2:  # TF 'same': total pad = k - stride, extra pixel after (hourglass.py:59)
            x = F.pad(x, (2, 3, 2, 3))
            y = F.conv2d(x, w, b, stride=2)
        else:
            y = F.conv2d(x, w, b, padding=k // 2)
        if residual is not None:
            y = y + residual
        if activation == "relu":
            y = torch.relu(y)
        y = self.q(y)                      # the conv epilogue stores bf16 (logits included)
        if activation == "sigmoid":
            y = torch.sigmoid(y)           # fp32 heat map computed from the stored logits
        self.taps[name] = y
        return y

    def sepconv(self, x, name, k, cin, cout, activation):
        """SeparableConv2D (hourglass.py:216-226): depthwise k x k 'same' (depth_multiplier 1, no bias, no activation), then
        pointwise 1x1 + bias + activation.  Keras variables: depthwise_kernel (k,k,cin,1), pointwise_kernel (1,1,cin,cout), bias."""
        if self.params is None:
            self.spec[name + "/depthwise_kernel"] = (k, k, cin, 1)
            self.spec[name + "/pointwise_kernel"] = (1, 1, cin, cout)
            self.spec[name + "/bias"] = (cout,)
            return x
        dw = self.params[name + "/depthwise_kernel"].permute(2, 3, 0, 1)       # (k,k,cin,1) -> (cin,1,k,k); fp32 on the device too
        t = self.q(F.conv2d(x, dw, None, padding=k // 2, groups=cin))          # the stencil kernel stores bf16
        w = self.q(self.params[name + "/pointwise_kernel"]).permute(3, 2, 0, 1)
        y = F.conv2d(t, w, self.params[name + "/bias"])
        if activation == "relu":
            y = torch.relu(y)
        y = self.q(y)
        self.taps[name + "/depthwise"] = t
        self.taps[name] = y
        return y

    def bn(self, x, c, residual=None):
        name = "batch_normalization" if self.bn_count == 0 else f"batch_normalization_{self.bn_count}"
        self.bn_count += 1
        if self.params is None:
            for s in ("gamma", "beta", "moving_mean", "moving_variance"):
                self.spec[f"{name}/{s}"] = (c,)
            return x
        g = self.params[name + "/gamma"].view(1, -1, 1, 1)
        b = self.params[name + "/beta"].view(1, -1, 1, 1)
        if self.training and self.emulate:  # same algebra as the kernels: E[y^2] - E[y]^2, scale/shift form
            mean = x.mean(dim=(0, 2, 3), keepdim=True)
            var = ((x * x).mean(dim=(0, 2, 3), keepdim=True) - mean * mean).clamp_min(0)
        elif self.training:
            mean = x.mean(dim=(0, 2, 3), keepdim=True)
            var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
        else:
            mean = self.params[name + "/moving_mean"].view(1, -1, 1, 1)
            var = self.params[name + "/moving_variance"].view(1, -1, 1, 1)
        if self.training and self.update_moving:  # TF fused BN feeds the unbiased variance to the moving average
            n = x.numel() / x.shape[1]
            with torch.no_grad():
                mm = self.params[name + "/moving_mean"]
                mv = self.params[name + "/moving_variance"]
                mm.mul_(BN_MOMENTUM).add_(mean.flatten() * (1 - BN_MOMENTUM))
                mv.mul_(BN_MOMENTUM).add_(var.flatten() * (n / max(n - 1, 1)) * (1 - BN_MOMENTUM))
        if self.emulate:
            sc = g * torch.rsqrt(var + BN_EPS)
            z = x * sc + (b - mean * sc)
        else:
            z = (x - mean) / torch.sqrt(var + BN_EPS) * g + b
        if residual is not None:
            z = z + residual
        return self.q(z)

    def pool(self, x):
        return x if self.params is None else F.max_pool2d(x, 2, 2)

    def up(self, x):
        return x if self.params is None else F.interpolate(x, scale_factor=2, mode="nearest")

    # ---- blocks (hourglass.py:184-206, 160-181, 127-157, 71-93, 54-68)
    def bottleneck(self, x, cin, cout, name):
        conv = self.sepconv if self.mobile else self.conv      # hourglass.py:209-231 vs :184-206: same wiring, other layer class
        skip = x
        if cin != cout:
            skip = conv(x, name + "_skip", 1, cin, cout, "relu")
        y = conv(x, name + "_conv_1x1_1", 1, cin, cout // 2, "relu")
        y = self.bn(y, cout // 2)
        y = conv(y, name + "_conv_3x3_2", 3, cout // 2, cout // 2, "relu")
        y = self.bn(y, cout // 2)
        y = conv(y, name + "_conv_1x1_3", 1, cout // 2, cout, "relu")
        y = self.bn(y, cout, residual=skip if self.params is not None else None)  # Add (hourglass.py:204)
        return y if self.params is not None else x

    def front(self, x, C):
        if self.params is not None:
            x = self.q(x)                  # the stem reads bf16 patches of the f32 image
        x = self.conv(x, "front_conv_1x1_1", 7, 3, 64, "relu", stride=2)
        x = self.bn(x, 64)
        x = self.bottleneck(x, 64, C // 2, "front_bottleneck_1")
        x = self.pool(x)
        x = self.bottleneck(x, C // 2, C // 2, "front_bottleneck_2")
        x = self.bottleneck(x, C // 2, C, "front_bottleneck_3")
        return x

    def hourglass(self, x, K, C, i, activation, last):
        hg = f"hg{i}"
        f1 = self.bottleneck(x, C, C, hg + "_downsample_f1")
        f2 = self.bottleneck(self.pool(f1), C, C, hg + "_downsample_f2")
        f4 = self.bottleneck(self.pool(f2), C, C, hg + "_downsample_f4")
        f8 = self.bottleneck(self.pool(f4), C, C, hg + "_downsample_f8")
        b = self.pool(f8)
        for j in (1, 2, 3):
            b = self.bottleneck(b, C, C, f"{hg}_downsample_f8_{j}")
        cur = b
        for f, nm in ((f8, "f8"), (f4, "f4"), (f2, "f2"), (f1, "f1")):
            s = self.bottleneck(f, C, C, f"{hg}_upsample_{nm}_short")
            a = self.q(s + self.up(cur)) if self.params is not None else s
            cur = self.bottleneck(a, C, C, f"{hg}_upsample_{nm}_merged")
        head = self.conv(cur, hg + "_conv_1x1_1", 1, C, C, "relu")
        head = self.bn(head, C)
        predict = self.conv(head, hg + "_conv_1x1_predict", 1, C, K, activation)
        nxt = None
        if not last:  # Keras prunes the last stack's re-injection branch (not on a path to an output)
            # Add()([head, head_m, x]) (hourglass.py:91), evaluated as two GEMM epilogues with residuals
            h2 = self.conv(head, hg + "_conv_1x1_2", 1, C, C, "linear", residual=x)
            pin = self.q(predict) if self.params is not None else predict
            nxt = self.conv(pin, hg + "_conv_1x1_3", 1, K, C, "linear", residual=h2)
        return nxt, predict

    def model(self, x, K, S, C, activation):
        x = self.front(x, C)
        outs = []
        for i in range(S):
            x, p = self.hourglass(x, K, C, i, activation, last=(i == S - 1))
            outs.append(p)
        return outs


def param_spec(num_classes=17, num_stacks=1, num_channels=256, mobile=False):
    """OrderedDict name -> shape in Keras creation order (kernels HWIO)."""
    b = _Builder(None, mobile=mobile)
    b.model(None, num_classes, num_stacks, num_channels, "sigmoid")
    return b.spec


def count_params(spec):
    tot = sum(int(np.prod(s)) for s in spec.values())
    non = sum(int(np.prod(s)) for n, s in spec.items() if "moving_" in n)
    return tot, tot - non, non


def init_params(spec, seed=2, perturb_bn=False):
    """Keras initialisers: glorot_uniform kernels, zero bias, gamma 1, beta 0, moving mean 0 / var 1.
    perturb_bn=True draws gamma/beta/moving stats at random instead (stronger parity test)."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in spec.items():
        if name.endswith("kernel"):       # kernel / depthwise_kernel / pointwise_kernel: glorot_uniform over Keras' fans
            k1, k2, cin, cout = shape
            limit = np.sqrt(6.0 / (k1 * k2 * cin + k1 * k2 * cout))
            out[name] = rng.uniform(-limit, limit, size=shape).astype(np.float32)
        elif name.endswith("/bias"):
            out[name] = (rng.normal(0, 0.05, size=shape) if perturb_bn else np.zeros(shape)).astype(np.float32)
        elif name.endswith("/gamma"):
            out[name] = (rng.uniform(0.5, 1.5, size=shape) if perturb_bn else np.ones(shape)).astype(np.float32)
        elif name.endswith("/beta"):
            out[name] = (rng.normal(0, 0.2, size=shape) if perturb_bn else np.zeros(shape)).astype(np.float32)
        elif name.endswith("/moving_mean"):
            out[name] = (rng.normal(0.3, 0.2, size=shape) if perturb_bn else np.zeros(shape)).astype(np.float32)
        elif name.endswith("/moving_variance"):
            out[name] = (rng.uniform(0.5, 1.5, size=shape) if perturb_bn else np.ones(shape)).astype(np.float32)
        else:
            raise KeyError(name)
    return out


def _fp32_exact(device):
    """fp32 means fp32: no TF32 in cuDNN / cuBLAS when the restatement runs on a GPU."""
    if str(device) != "cpu":
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False


def forward(params_np, images_nhwc, num_classes, num_stacks, num_channels, activation="sigmoid", training=True,
            requires_grad=False, update_moving=False, return_taps=False, emulate_bf16=False, device="cpu", mobile=False):
    """images (B,H,W,3) f32 -> list of S tensors (B,h,w,K) f32 (NHWC).  Returns (outputs, torch params)."""
    _fp32_exact(device)
    params = OrderedDict((k, torch.tensor(v, dtype=torch.float32, device=device, requires_grad=requires_grad and "moving_" not in k))
                         for k, v in params_np.items())
    x = torch.as_tensor(np.asarray(images_nhwc), dtype=torch.float32).to(device).permute(0, 3, 1, 2)
    b = _Builder(params, training=training, update_moving=update_moving, emulate_bf16=emulate_bf16, mobile=mobile)
    outs = b.model(x, num_classes, num_stacks, num_channels, activation)
    outs = [o.permute(0, 2, 3, 1) for o in outs]
    if return_taps:
        return outs, params, b.taps
    return outs, params


def torch_loss(kind, y_true, y_pred):
    """Keras-reduced scalar of one output (mean over everything the reference loss fn returns)."""
    t, p = y_true, y_pred
    if kind == "weighted_mse":
        w = (t > 0).float() * 81 + 1
        return ((t - p) ** 2 * w).mean()
    if kind == "mse":
        return ((t - p) ** 2).mean()
    if kind == "weighted_keypoint_mse":
        kw = 1.0 - (t.sum(dim=(1, 2), keepdim=True) == 0).float()
        return ((t - p) ** 2 * kw).mean()
    if kind == "iou":
        eps = 1e-7
        inter = (t * p).sum(dim=(1, 2))
        union = (t * t).sum(dim=(1, 2)) + (p * p).sum(dim=(1, 2)) - inter
        return (1 - ((inter + eps) / (union + eps)).mean(dim=-1)).mean()
    raise ValueError(kind)


def loss_and_grads(params_np, images, y_true, kind, num_classes, num_stacks, num_channels, activation="sigmoid",
                   emulate_bf16=False, device="cpu", mobile=False):
    """Training-mode forward, sum of per-stack losses (Keras compile with one loss fn), backward."""
    outs, params = forward(params_np, images, num_classes, num_stacks, num_channels, activation, training=True,
                           requires_grad=True, emulate_bf16=emulate_bf16, device=device, mobile=mobile)
    t = torch.as_tensor(np.asarray(y_true), dtype=torch.float32).to(device)
    losses = [torch_loss(kind, t, o) for o in outs]
    total = sum(losses)
    total.backward()
    grads = OrderedDict((k, v.grad.cpu().numpy()) for k, v in params.items() if v.requires_grad)
    return [o.detach().cpu().numpy() for o in outs], [float(l.detach()) for l in losses], grads


def adam_step(w, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """Keras legacy OptimizerV2 Adam (epsilon outside the bias correction). In-place on numpy arrays."""
    m *= b1
    m += (1 - b1) * g
    v *= b2
    v += (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    w -= (lr_t * m / (np.sqrt(v) + eps)).astype(w.dtype)
