"""Keras layer ordering (keras_graph.py) pinned by the reference's own saved model.summary(), and TensorFlow checkpoint
files (tf_checkpoint.py): table format, bundle entries, Keras object-graph addressing, model.save_weights / load_weights.
CPU only (weights live on the host until a GPU is used)."""
import json
import os
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def hgb():
    import hgb200
    return hgb200


def test_layer_order_reproduces_the_reference_summary(hgb, golden_dir):
    """dev/making_hourglass.ipynb cell 3: all 147 layers -- names (explicit and auto-generated), classes, parameter counts,
    inbound connections -- in `model.layers` order."""
    from hgb200 import keras_graph as kg
    rows = json.load(open(os.path.join(golden_dir, "keras_summary_1stack.json")))
    _layers, outputs = kg.build_hourglass_graph(17, 1, 256)
    got = [[l.name, l.cls, l.param_count(), [p.name for p in l.inbound]] for l in kg.model_layers(outputs)]
    assert got == rows


@pytest.mark.parametrize("stacks,total", [(1, 3659665), (2, 7034530), (4, 13784260), (8, 27283720)])
def test_checkpoint_keys_cover_the_parameter_table(hgb, stacks, total):
    from hgb200 import keras_graph as kg
    layers, outputs = kg.build_hourglass_graph(17, stacks, 256)
    ordered = kg.model_layers(outputs)
    assert sum(l.param_count() for l in ordered) == total          # counts the reference printed (Train.ipynb / SURVEY 8a)
    pruned = {l.name for l in layers} - {l.name for l in ordered}
    assert pruned == {f"hg{stacks - 1}_conv_1x1_2", f"hg{stacks - 1}_conv_1x1_3", f"add_{5 * stacks - 1}"}
    keys = kg.checkpoint_keys(17, stacks, 256)
    model = hgb.create_hourglass_model(17, stacks, 256, (256, 256, 3), "sigmoid")
    assert set(keys) == set(model._table)
    for name, key in keys.items():
        assert key.endswith("/" + name.rsplit("/", 1)[1] + "/.ATTRIBUTES/VARIABLE_VALUE")
    idx = [int(k.split("/")[0].split("-")[1]) for k in keys.values()]
    assert idx == sorted(idx) and idx[0] == 0 and idx[-1] == len({k.rsplit('/', 1)[0] for k in keys}) - 1
    # depth ordering interleaves the skip branch with the main chain: not creation order
    if stacks == 1:
        names = [k.rsplit("/", 1)[0] for k in keys]
        assert names.index("front_bottleneck_1_skip") > names.index("front_bottleneck_1_conv_1x1_3")


def _bitwise_crc32c(data):
    c = 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    return c ^ 0xFFFFFFFF


def test_table_format_structure_and_round_trip(hgb, tmp_path):
    from hgb200 import tf_checkpoint as tc
    rng = np.random.default_rng(0)
    tensors = {f"layer_with_weights-{i}/{a}/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal((3, i + 1)).astype(np.float32)
               for i in range(40) for a in ("kernel", "bias")}
    tensors["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"] = np.asarray(12345, np.int64)
    tensors["_CHECKPOINTABLE_OBJECT_GRAPH"] = b"\x0a\x00graph bytes"
    prefix = str(tmp_path / "ckpt")
    tc.write_checkpoint(prefix, tensors)
    raw = open(prefix + ".index", "rb").read()
    assert raw[-8:] == bytes.fromhex("57fb808b247547db")                       # LevelDB table magic, little endian
    # first data block: first entry is key "" (the bundle header) with no shared prefix; its trailer CRC is the masked CRC-32C
    assert raw[0] == 0 and raw[1] == 0
    back = tc.read_checkpoint(prefix)
    assert list(back) == sorted(tensors, key=lambda s: s.encode())
    for k, v in tensors.items():
        if isinstance(v, bytes):
            assert back[k] == v
        else:
            assert back[k].dtype == v.dtype and back[k].shape == v.shape
            np.testing.assert_array_equal(back[k], v)
    # tensors are laid out in key order in the data shard
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    first = sorted(tensors, key=lambda s: s.encode())[1]
    assert data[len(b"\x0c") + 4 + len(tensors["_CHECKPOINTABLE_OBJECT_GRAPH"]):][:tensors[first].nbytes] == tensors[first].tobytes()
    # several small blocks: prefix compression + restart arrays + index lookups
    entries = [(b"", b"h")] + [(f"key{i:04d}".encode(), bytes([i % 251]) * (i % 7)) for i in range(300)]
    tc._write_table(str(tmp_path / "small.index"), entries, block_size=256)
    got = tc._read_table(str(tmp_path / "small.index"))
    assert list(got.items()) == entries
    small = open(tmp_path / "small.index", "rb").read()
    footer = small[-48:]
    _o, at = tc._varint(footer, 0)
    _s, at = tc._varint(footer, at)
    ioff, at = tc._varint(footer, at)
    isize, at = tc._varint(footer, at)
    handles = tc._read_block(small, ioff, isize)
    assert len(handles) > 5                                                    # really split into blocks
    off, a2 = tc._varint(handles[0][1], 0)
    size, _ = tc._varint(handles[0][1], a2)
    stored = struct.unpack("<I", small[off + size + 1:off + size + 5])[0]
    crc = _bitwise_crc32c(small[off:off + size + 1])
    assert stored == (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    # second entry of a block shares the "key0" prefix with the first
    blk = small[off:off + size]
    p = 0
    sh0, p = tc._varint(blk, p)
    ns0, p = tc._varint(blk, p)
    vl0, p = tc._varint(blk, p)
    p += ns0 + vl0
    sh1, p = tc._varint(blk, p)
    assert sh0 == 0 and sh1 == 0                                               # "" then "key0000": nothing shared
    # corruption is detected
    bad = bytearray(raw)
    bad[10] ^= 1
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        tc.read_checkpoint(prefix)
    open(prefix + ".index", "wb").write(raw)
    d = bytearray(data)
    d[-3] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(d))
    with pytest.raises(ValueError, match="checksum mismatch in tensor"):
        tc.read_checkpoint(prefix)
    with pytest.raises(ValueError, match="magic"):
        open(tmp_path / "junk.index", "wb").write(bytes(64))
        tc.read_checkpoint(str(tmp_path / "junk"))


def test_model_weights_round_trip_through_tf_checkpoint(hgb, tmp_path):
    from hgb200 import tf_checkpoint as tc
    a = hgb.create_hourglass_model(17, 2, 256, (256, 256, 3), "sigmoid")
    rng = np.random.default_rng(1)
    w = {k: rng.standard_normal(sh).astype(np.float32) for k, (sh, _o, _t) in a._table.items()}
    a.set_weights_dict(w)
    prefix = str(tmp_path / "checkpoints" / "best_val_loss_weights.ckpt")
    a.save_weights(prefix)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")      # trainer.py:150-170 globs these
    bundle = tc.read_checkpoint(prefix)
    assert bundle["layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"].shape == (7, 7, 3, 64)      # Keras HWIO layout
    np.testing.assert_array_equal(bundle["layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"], w["front_conv_1x1_1/kernel"])
    named = tc.parse_object_graph(bundle["_CHECKPOINTABLE_OBJECT_GRAPH"])
    assert named["layer_with_weights-1/moving_mean/.ATTRIBUTES/VARIABLE_VALUE"] == "batch_normalization/moving_mean"
    assert len(named) == len(a._table)
    b = hgb.create_hourglass_model(17, 2, 256, (256, 256, 3), "sigmoid")
    b.load_weights(prefix)
    got = b.get_weights_dict()
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])
    # a checkpoint of another architecture is rejected, by missing keys or by the names in its object graph
    with pytest.raises((KeyError, ValueError)):
        hgb.create_hourglass_model(17, 4, 256, (256, 256, 3), "sigmoid").load_weights(prefix)
    # the 1-stack model is a prefix of the 2-stack one in Keras' layer order (its head comes before the re-injection convs):
    # like Keras, loading takes the matching objects and ignores the rest of the file
    one = hgb.create_hourglass_model(17, 1, 256, (256, 256, 3), "sigmoid").load_weights(prefix).get_weights_dict()
    np.testing.assert_array_equal(one["hg0_conv_1x1_predict/kernel"], w["hg0_conv_1x1_predict/kernel"])
    # Adam slots and the step counter travel in Keras' slot-variable keys
    m = {k: rng.standard_normal(sh).astype(np.float32) for k, (sh, _o, tr) in a._table.items() if tr}
    v = {k: np.abs(rng.standard_normal(sh)).astype(np.float32) for k, (sh, _o, tr) in a._table.items() if tr}
    tc.save_keras_weights(a, prefix, adam=(777, m, v))
    weights, adam = tc.load_keras_weights(b, prefix)
    assert adam[0] == 777 and set(adam[1]) == set(m)
    np.testing.assert_array_equal(adam[2]["hg1_conv_1x1_predict/bias"], v["hg1_conv_1x1_predict/bias"])
    keys = tc.read_checkpoint(prefix)
    assert "layer_with_weights-0/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE" in keys
    assert not any("moving_mean/.OPTIMIZER_SLOT" in k for k in keys)
    # a checkpoint whose root object holds the model as an attribute (tf.train.Checkpoint(model=...)): keys carry the path
    nested = {("model/" + k if k != "_CHECKPOINTABLE_OBJECT_GRAPH" else k): v for k, v in tc.read_checkpoint(prefix).items()}
    nested["save_counter/.ATTRIBUTES/VARIABLE_VALUE"] = np.asarray(1, np.int64)
    tc.write_checkpoint(str(tmp_path / "ckpt-1"), nested)
    w2, adam2 = tc.load_keras_weights(b, str(tmp_path / "ckpt-1"))
    np.testing.assert_array_equal(w2["hg1_conv_1x1_predict/kernel"], w["hg1_conv_1x1_predict/kernel"])
    assert adam2 is None or adam2[0] == 777
    # the npz/json payload of earlier checkpoints still loads
    a.save_weights(prefix, save_format="hgb")
    c = hgb.create_hourglass_model(17, 2, 256, (256, 256, 3), "sigmoid")
    c.load_weights(prefix)
    np.testing.assert_array_equal(c.get_weights_dict()["hg0_conv_1x1_2/kernel"], w["hg0_conv_1x1_2/kernel"])


def test_model_utils_checkpoint_helpers(hgb, tmp_path):
    from hgb200.utilities import model_utils
    a = hgb.create_hourglass_model(17, 1, 256, (256, 256, 3), "sigmoid")
    d = tmp_path / "checkpoints"
    for name in ("E5_2022-03-01_cont.ckpt", "E12_2022-03-02_cont.ckpt", "best_val_loss_weights.ckpt"):
        a.save_weights(str(d / name))
    names, epochs = model_utils.get_epochs_from_ckpt_path(str(d))
    assert epochs == [12, 5, -1]                                   # name order, as the reference sorts them
    assert names[-1].endswith("best_val_loss_weights.ckpt") and names[0].endswith("E12_2022-03-02_cont.ckpt")
    b = hgb.create_hourglass_model(17, 1, 256, (256, 256, 3), "sigmoid")
    out = model_utils.compile_model_from_checkpoint(b, names[1], hgb.Adam(1e-3), hgb.loss.weighted_mse)
    assert out is b and b.optimizer is not None
    np.testing.assert_array_equal(b.get_weights_dict()["front_conv_1x1_1/kernel"], a.get_weights_dict()["front_conv_1x1_1/kernel"])
