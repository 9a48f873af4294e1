"""Independent pins for the TensorFlow container formats restated in tfrecord.py / tf_checkpoint.py.

TensorFlow itself is absent, but the image ships (a) tensorboard, which carries TensorFlow's own Python implementation of the
TFRecord framing (RecordWriter / PyRecordReader_New / masked_crc32c) and its compiled protos (TrackableObjectGraph,
TensorShapeProto, DataType, VersionDef), and (b) the protobuf runtime, with which the public schemas of tf.train.Example
and of the tensor-bundle protos are declared below and serialized / parsed by Google's encoder instead of ours.
Every check crosses implementations: ours writes -> theirs reads, theirs writes -> ours reads."""
import struct

import numpy as np
import pytest

tb_stub = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402


@pytest.fixture(scope="module")
def hgb():
    import hgb200
    return hgb200


# ------------------------------------------------------------------ TFRecord framing
def test_crc_and_record_framing_against_tensorflows_python_implementation(hgb, tmp_path):
    from hgb200 import tfrecord
    from tensorboard.summary.writer.record_writer import RecordWriter
    rng = np.random.default_rng(0)
    payloads = [bytes(rng.integers(0, 256, int(n), dtype=np.uint8)) for n in (0, 1, 7, 8, 9, 63, 64, 65, 1000, 70001)]
    for p in payloads:
        assert tfrecord.masked_crc32c(p) == tb_stub.masked_crc32c(p)
        assert hgb._lib.lib.hgb_crc32c(p, len(p)) == tb_stub.crc32c(p)
    theirs, ours = str(tmp_path / "theirs.tfrec"), str(tmp_path / "ours.tfrec")
    with open(theirs, "wb") as f:
        w = RecordWriter(f)
        for p in payloads:
            w.write(p)
    assert list(tfrecord.read_records(theirs)) == payloads                     # theirs -> ours (CRCs verified)
    tfrecord.write_records(ours, payloads)
    assert open(ours, "rb").read() == open(theirs, "rb").read()                 # byte-identical files
    reader, back = tb_stub.PyRecordReader_New(ours), []
    while True:
        try:
            reader.GetNext()                                                    # ours -> theirs (it checks both CRCs)
        except Exception as ex:  # tensorboard's OutOfRangeError at end of file
            assert "No more events" in str(ex) or type(ex).__name__ == "OutOfRangeError", ex
            break
        back.append(reader.record())
    assert back == payloads


# ------------------------------------------------------------------ tf.train.Example through the protobuf runtime
def _example_messages():
    """example.proto / feature.proto as published: BytesList{repeated bytes value=1}, FloatList{repeated float value=1 [packed]},
    Int64List{repeated int64 value=1 [packed]}, Feature{oneof kind: bytes_list=1, float_list=2, int64_list=3},
    Features{map<string, Feature> feature=1}, Example{Features features=1}."""
    fd = descriptor_pb2.FileDescriptorProto(name="hgb_test_example.proto", package="hgbtest", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    for name, ftype in (("BytesList", T.TYPE_BYTES), ("FloatList", T.TYPE_FLOAT), ("Int64List", T.TYPE_INT64)):
        m = msg(name)
        f = m.field.add(name="value", number=1, type=ftype, label=T.LABEL_REPEATED)
        if ftype != T.TYPE_BYTES:
            f.options.packed = True
    feature = msg("Feature")
    feature.oneof_decl.add(name="kind")
    for n, (fname, tname) in enumerate((("bytes_list", "BytesList"), ("float_list", "FloatList"), ("int64_list", "Int64List")), 1):
        feature.field.add(name=fname, number=n, type=T.TYPE_MESSAGE, type_name=f".hgbtest.{tname}", label=T.LABEL_OPTIONAL, oneof_index=0)
    features = msg("Features")
    entry = features.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=T.TYPE_MESSAGE, type_name=".hgbtest.Feature", label=T.LABEL_OPTIONAL)
    features.field.add(name="feature", number=1, type=T.TYPE_MESSAGE, type_name=".hgbtest.Features.FeatureEntry", label=T.LABEL_REPEATED)
    msg("Example").field.add(name="features", number=1, type=T.TYPE_MESSAGE, type_name=".hgbtest.Features", label=T.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName(f"hgbtest.{n}")) for n in ("Example", "Feature")}


def test_example_wire_format_against_the_protobuf_runtime(hgb):
    from hgb200 import tfrecord
    M = _example_messages()
    rng = np.random.default_rng(1)
    for trial in range(25):
        values = {"ann_id": np.array([int(rng.integers(0, 2 ** 40))]), "neg": rng.integers(-2 ** 62, 2 ** 62, int(rng.integers(0, 6))),
                  "keypoints/x": rng.random(int(rng.integers(0, 20))).astype(np.float32), "bbox_x": np.array([np.float32(-3.25)]),
                  "image": bytes(rng.integers(0, 256, int(rng.integers(0, 500)), dtype=np.uint8)), "coco_url": b"http://x/y.jpg"}
        ex = M["Example"]()
        for name, v in values.items():
            f = ex.features.feature[name]
            if isinstance(v, bytes):
                f.bytes_list.value.append(v)
            elif v.dtype.kind == "f":
                f.float_list.value.extend(float(x) for x in v)
                if not len(v):
                    f.float_list.SetInParent()
            else:
                f.int64_list.value.extend(int(x) for x in v)
                if not len(v):
                    f.int64_list.SetInParent()
        theirs = ex.SerializeToString(deterministic=True)
        assert tfrecord.build_example(values) == theirs, trial                 # ours == Google's encoder, byte for byte
        for parser in (tfrecord.parse_example, tfrecord._parse_example_py):
            got = parser(ex.SerializeToString())                               # theirs (any map order) -> ours
            assert set(got) == set(values)
            for name, v in values.items():
                if isinstance(v, bytes):
                    assert got[name] == [v]
                else:
                    np.testing.assert_array_equal(got[name], v)
        back = M["Example"].FromString(tfrecord.build_example(values))         # ours -> theirs
        assert list(back.features.feature["neg"].int64_list.value) == values["neg"].tolist()
        assert back.features.feature["image"].bytes_list.value[0] == values["image"]
        assert back.features.feature["keypoints/x"].float_list.value == pytest.approx(values["keypoints/x"].tolist())


# ------------------------------------------------------------------ tensor-bundle protos and the Keras object graph
def _bundle_messages():
    """tensor_bundle.proto as published, on top of TensorFlow's own compiled TensorShapeProto / DataType / VersionDef:
    BundleHeaderProto{int32 num_shards=1; Endianness endianness=2; VersionDef version=3},
    BundleEntryProto{DataType dtype=1; TensorShapeProto shape=2; int32 shard_id=3; int64 offset=4; int64 size=5; fixed32 crc32c=6}."""
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2
    pool = descriptor_pool.DescriptorPool()
    for mod in (tensor_shape_pb2, types_pb2, versions_pb2):
        proto = descriptor_pb2.FileDescriptorProto()
        mod.DESCRIPTOR.CopyToProto(proto)
        pool.Add(proto)
    T = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="hgb_test_bundle.proto", package="hgbtest", syntax="proto3",
                                            dependency=[tensor_shape_pb2.DESCRIPTOR.name, types_pb2.DESCRIPTOR.name, versions_pb2.DESCRIPTOR.name])
    header = fd.message_type.add(name="BundleHeaderProto")
    header.field.add(name="num_shards", number=1, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    header.field.add(name="endianness", number=2, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    header.field.add(name="version", number=3, type=T.TYPE_MESSAGE, type_name=".tensorboard.VersionDef", label=T.LABEL_OPTIONAL)
    entry = fd.message_type.add(name="BundleEntryProto")
    entry.field.add(name="dtype", number=1, type=T.TYPE_ENUM, type_name=".tensorboard.DataType", label=T.LABEL_OPTIONAL)
    entry.field.add(name="shape", number=2, type=T.TYPE_MESSAGE, type_name=".tensorboard.TensorShapeProto", label=T.LABEL_OPTIONAL)
    entry.field.add(name="shard_id", number=3, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    entry.field.add(name="offset", number=4, type=T.TYPE_INT64, label=T.LABEL_OPTIONAL)
    entry.field.add(name="size", number=5, type=T.TYPE_INT64, label=T.LABEL_OPTIONAL)
    entry.field.add(name="crc32c", number=6, type=T.TYPE_FIXED32, label=T.LABEL_OPTIONAL)
    pool.Add(fd)
    return (message_factory.GetMessageClass(pool.FindMessageTypeByName("hgbtest.BundleHeaderProto")),
            message_factory.GetMessageClass(pool.FindMessageTypeByName("hgbtest.BundleEntryProto")))


def test_bundle_entries_against_tensorflows_compiled_protos(hgb, tmp_path):
    from hgb200 import tf_checkpoint as tc
    from tensorboard.compat.proto import types_pb2
    assert (tc.DT_FLOAT, tc.DT_DOUBLE, tc.DT_INT32, tc.DT_STRING, tc.DT_INT64) == (
        types_pb2.DT_FLOAT, types_pb2.DT_DOUBLE, types_pb2.DT_INT32, types_pb2.DT_STRING, types_pb2.DT_INT64)
    Header, Entry = _bundle_messages()
    rng = np.random.default_rng(2)
    tensors = {"layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal((7, 7, 3, 64)).astype(np.float32),
               "layer_with_weights-0/bias/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal(64).astype(np.float32),
               "optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE": np.asarray(4321, np.int64)}
    prefix = str(tmp_path / "w.ckpt")
    tc.write_checkpoint(prefix, tensors)
    table = tc._read_table(prefix + ".index")
    header = Header.FromString(table[b""])                                      # ours -> theirs
    assert header.num_shards == 1 and header.endianness == 0 and header.version.producer == 1
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    for name, arr in tensors.items():
        e = Entry.FromString(table[name.encode()])
        assert e.dtype == (types_pb2.DT_INT64 if arr.dtype == np.int64 else types_pb2.DT_FLOAT)
        assert [d.size for d in e.shape.dim] == list(arr.shape) and e.shard_id == 0 and e.size == arr.nbytes
        blob = data[e.offset:e.offset + e.size]
        assert blob == arr.tobytes() and e.crc32c == tb_stub.masked_crc32c(blob)   # TensorFlow's own CRC over the tensor bytes
    # theirs -> ours: entries serialized by the protobuf runtime inside a table written by us parse back to the same tensors
    entries, blob, offset = [(b"", Header(num_shards=1).SerializeToString())], b"", 0
    for name in sorted(tensors, key=lambda s: s.encode()):
        arr = tensors[name]
        e = Entry(dtype=types_pb2.DT_INT64 if arr.dtype == np.int64 else types_pb2.DT_FLOAT, offset=offset, size=arr.nbytes,
                  crc32c=tb_stub.masked_crc32c(arr.tobytes()))
        e.shape.SetInParent()
        for d in arr.shape:
            e.shape.dim.add(size=d)
        entries.append((name.encode(), e.SerializeToString()))
        blob += arr.tobytes()
        offset += arr.nbytes
    other = str(tmp_path / "theirs.ckpt")
    tc._write_table(other + ".index", entries)
    open(other + ".data-00000-of-00001", "wb").write(blob)
    back = tc.read_checkpoint(other)
    for name, arr in tensors.items():
        np.testing.assert_array_equal(back[name], arr)
        assert back[name].shape == arr.shape


def test_object_graph_against_tensorflows_compiled_proto(hgb, tmp_path):
    from hgb200 import tf_checkpoint as tc
    from tensorboard.compat.proto.trackable_object_graph_pb2 import TrackableObjectGraph
    model = hgb.create_hourglass_model(17, 1, 256, (256, 256, 3), "sigmoid")
    keys, layer_weights = tc._layer_weights(model)
    graph = TrackableObjectGraph.FromString(tc._object_graph(layer_weights, with_optimizer=True))     # ours -> theirs
    root = graph.nodes[0]
    by_name = {c.local_name: c.node_id for c in root.children}
    assert len([n for n in by_name if n.startswith("layer_with_weights-")]) == len(layer_weights) and "optimizer" in by_name
    first = graph.nodes[by_name["layer_with_weights-0"]]
    assert [c.local_name for c in first.children] == ["kernel", "bias"]
    kernel = graph.nodes[first.children[0].node_id]
    assert (kernel.attributes[0].name, kernel.attributes[0].full_name, kernel.attributes[0].checkpoint_key) == (
        "VARIABLE_VALUE", "front_conv_1x1_1/kernel", "layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE")
    opt = graph.nodes[by_name["optimizer"]]
    assert [c.local_name for c in opt.children] == ["iter"]
    trainable = [k for k, (_s, _o, tr) in model._table.items() if tr]
    assert len(opt.slot_variables) == 2 * len(trainable)
    slot = opt.slot_variables[0]
    assert slot.slot_name == "m" and slot.original_variable_node_id == first.children[0].node_id
    assert graph.nodes[slot.slot_variable_node_id].attributes[0].checkpoint_key == \
        "layer_with_weights-0/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"
    # theirs -> ours: a graph assembled with TensorFlow's proto classes the way Keras lays it out
    g = TrackableObjectGraph()
    g.nodes.add()
    for n, (layer, attrs) in enumerate(layer_weights[:3]):
        ln = len(g.nodes)
        g.nodes.add()
        g.nodes[0].children.add(node_id=ln, local_name=f"layer_with_weights-{n}")
        for attr in attrs:
            vn = len(g.nodes)
            node = g.nodes.add()
            node.attributes.add(name="VARIABLE_VALUE", full_name=f"{layer}/{attr}", checkpoint_key=f"layer_with_weights-{n}/{attr}/.ATTRIBUTES/VARIABLE_VALUE")
            g.nodes[ln].children.add(node_id=vn, local_name=attr)
    named = tc.parse_object_graph(g.SerializeToString())
    assert named["layer_with_weights-1/moving_variance/.ATTRIBUTES/VARIABLE_VALUE"] == "batch_normalization/moving_variance"
    assert len(named) == sum(len(a) for _l, a in layer_weights[:3])
    # and the string-tensor encoding that carries the graph inside the bundle: varint length | masked CRC of the length | bytes
    blob, crc = tc._encode_string_scalar(b"abc")
    length32 = struct.pack("<I", 3)
    assert blob == b"\x03" + struct.pack("<I", tb_stub.masked_crc32c(length32)) + b"abc"
    assert crc == tb_stub.masked_crc32c(length32 + struct.pack("<I", tb_stub.masked_crc32c(length32)) + b"abc")
