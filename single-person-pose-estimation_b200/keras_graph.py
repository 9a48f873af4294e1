"""The Keras layer graph of the reference's hourglass and the order Keras puts its layers in.

Why this exists: the reference saves TF-format checkpoints (`model.save_weights('...ckpt')`, trainer.py:63-64,141; SavedModel
variables, save_model.ipynb).  In those files a weight is addressed as `layer_with_weights-<N>/<attr>/.ATTRIBUTES/VARIABLE_VALUE`,
where N counts the layers that own weights in `model.layers` order -- and for a functional model that order is NOT creation
order: Keras sorts layers by their depth from the outputs and breaks ties by the order a depth-first walk from the outputs
first meets them.  This module rebuilds the graph from model/hourglass.py:5-206 (explicit and auto-generated layer names
included) and restates that ordering rule, so `tf_checkpoint.py` can translate between the two addressings.

Pinned by the reference itself: `tests/golden/keras_summary_1stack.json` is the `model.summary()` the reference saved in
dev/making_hourglass.ipynb (147 layers with their connections, in `model.layers` order); tests/test_cpu_checkpoint.py requires
this module to reproduce it exactly.
"""
from __future__ import annotations

from collections import OrderedDict, defaultdict


class Layer:
    __slots__ = ("name", "cls", "inbound", "weights", "channels")

    def __init__(self, name, cls, inbound, weights, channels):
        self.name, self.cls, self.inbound, self.weights, self.channels = name, cls, list(inbound), weights, channels

    def param_count(self):
        n = 0
        for _attr, shape in self.weights:
            c = 1
            for s in shape:
                c *= s
            n += c
        return n


class _Builder:
    """Creates layers the way the reference's calls do, with Keras' per-class name counters (reset by K.clear_session, :7)."""

    def __init__(self, mobile=False):
        self.layers, self.counters, self.mobile = [], defaultdict(int), mobile

    def _auto(self, base):
        k = self.counters[base]
        self.counters[base] += 1
        return base if k == 0 else f"{base}_{k}"

    def _add(self, layer):
        self.layers.append(layer)
        return layer

    def input(self, channels):
        self.counters["input"] += 1
        return self._add(Layer(f"input_{self.counters['input']}", "InputLayer", [], [], channels))

    def conv(self, x, filters, k, name):
        return self._add(Layer(name, "Conv2D", [x], [("kernel", (k, k, x.channels, filters)), ("bias", (filters,))], filters))

    def sepconv(self, x, filters, k, name):
        # SeparableConv2D (model/hourglass.py:216-226): variables depthwise_kernel, pointwise_kernel, bias in that order
        return self._add(Layer(name, "SeparableConv2D", [x], [("depthwise_kernel", (k, k, x.channels, 1)),
                                                               ("pointwise_kernel", (1, 1, x.channels, filters)),
                                                               ("bias", (filters,))], filters))

    def bn(self, x):
        c = x.channels
        return self._add(Layer(self._auto("batch_normalization"), "BatchNormalization", [x],
                               [("gamma", (c,)), ("beta", (c,)), ("moving_mean", (c,)), ("moving_variance", (c,))], c))

    def pool(self, x):
        return self._add(Layer(self._auto("max_pooling2d"), "MaxPooling2D", [x], [], x.channels))

    def up(self, x):
        return self._add(Layer(self._auto("up_sampling2d"), "UpSampling2D", [x], [], x.channels))

    def add(self, xs, name=None):
        return self._add(Layer(name or self._auto("add"), "Add", xs, [], xs[0].channels))

    # model/hourglass.py:184-206 (bottleneck_block) / :209-231 (bottleneck_block_mobile)
    def bottleneck(self, x, out, name):
        conv = self.sepconv if self.mobile else self.conv
        skip = x if x.channels == out else conv(x, out, 1, name + "_skip")
        y = self.bn(conv(x, out // 2, 1, name + "_conv_1x1_1"))
        y = self.bn(conv(y, out // 2, 3, name + "_conv_3x3_2"))
        y = self.bn(conv(y, out, 1, name + "_conv_1x1_3"))
        return self.add([skip, y], name + "_add")


def build_hourglass_graph(num_classes=17, num_stacks=1, num_channels=256, in_channels=3, mobile=False):
    """-> (layers in creation order, output layers).  Mirrors create_hourglass_model (model/hourglass.py:5-32)."""
    b = _Builder(mobile)
    c = num_channels
    x = b.input(in_channels)
    # front module (:54-68)
    x = b.bn(b.conv(x, 64, 7, "front_conv_1x1_1"))
    x = b.bottleneck(x, c // 2, "front_bottleneck_1")
    x = b.pool(x)
    x = b.bottleneck(x, c // 2, "front_bottleneck_2")
    x = b.bottleneck(x, c, "front_bottleneck_3")
    outputs = []
    for i in range(num_stacks):
        name = f"hg{i}"
        # downsample (:160-181)
        f1 = b.bottleneck(x, c, name + "_downsample_f1")
        f2 = b.bottleneck(b.pool(f1), c, name + "_downsample_f2")
        f4 = b.bottleneck(b.pool(f2), c, name + "_downsample_f4")
        f8 = b.bottleneck(b.pool(f4), c, name + "_downsample_f8")
        # bottom (:127-140)
        y = b.pool(f8)
        for j in (1, 2, 3):
            y = b.bottleneck(y, c, f"{name}_downsample_f8_{j}")
        # upsample + merge (:143-157, :96-124)
        for feat, tag in ((f8, "f8"), (f4, "f4"), (f2, "f2"), (f1, "f1")):
            short = b.bottleneck(feat, c, f"{name}_upsample_{tag}_short")
            up = b.up(y)
            y = b.bottleneck(b.add([short, up]), c, f"{name}_upsample_{tag}_merged")
        # heads (:71-93)
        head = b.bn(b.conv(y, c, 1, name + "_conv_1x1_1"))
        predict = b.conv(head, num_classes, 1, name + "_conv_1x1_predict")
        head2 = b.conv(head, c, 1, name + "_conv_1x1_2")
        head_m = b.conv(predict, c, 1, name + "_conv_1x1_3")
        x = b.add([head2, head_m, x])
        outputs.append(predict)
    return b.layers, outputs


def model_layers(outputs):
    """`keras.Model(inputs, outputs).layers`: the functional-API ordering rule (every layer here is called once, so a layer
    is its single node).  (1) depth-first walk from each output through the inbound layers in call order: first visit
    numbers a layer (traversal index), completion appends it (post-order); (2) walking that list backwards, depth(layer) is
    kept and every inbound layer gets max(own, depth + 1); (3) layers are listed by decreasing depth, ties by traversal index.
    Layers that no output depends on (the last stack's re-injection convolutions) are not part of the model."""
    index, post, done = {}, [], set()
    for out in outputs:
        if out in done:
            continue
        stack = [(out, 0)]
        if out not in index:
            index[out] = len(index)
        while stack:
            layer, k = stack[-1]
            if k < len(layer.inbound):
                stack[-1] = (layer, k + 1)
                nxt = layer.inbound[k]
                if nxt not in done:
                    if nxt not in index:
                        index[nxt] = len(index)
                    stack.append((nxt, 0))
            else:
                stack.pop()
                if layer not in done:
                    done.add(layer)
                    post.append(layer)
    depth = {}
    for layer in reversed(post):
        d = depth.setdefault(layer, 0)
        for parent in layer.inbound:
            depth[parent] = max(d + 1, depth.get(parent, 0))
    by_depth = defaultdict(list)
    for layer, d in depth.items():
        by_depth[d].append(layer)
    ordered = []
    for d in sorted(by_depth, reverse=True):
        ordered.extend(sorted(by_depth[d], key=lambda l: index[l]))
    return ordered


def checkpoint_keys(num_classes=17, num_stacks=1, num_channels=256, mobile=False):
    """OrderedDict {'<layer name>/<attr>' (the names of hgb_model's parameter table) ->
    'layer_with_weights-<N>/<attr>/.ATTRIBUTES/VARIABLE_VALUE'} in `model.layers` order."""
    _layers, outputs = build_hourglass_graph(num_classes, num_stacks, num_channels, mobile=mobile)
    keys, n = OrderedDict(), 0
    for layer in model_layers(outputs):
        if not layer.weights:
            continue
        for attr, _shape in layer.weights:
            keys[f"{layer.name}/{attr}"] = f"layer_with_weights-{n}/{attr}/.ATTRIBUTES/VARIABLE_VALUE"
        n += 1
    return keys
