"""TensorFlow checkpoint files (tensor bundles) without TensorFlow: read the weights the reference saved, write ones it can find.

The reference stores weights with `model.save_weights('<dir>/best_val_loss_weights.ckpt')` / `E{n}_{date}_cont.ckpt`
(trainer.py:40-64,141) and as SavedModel variables (save_model.ipynb; utilities/model_utils.py:5-44).  All of these are TF
"tensor bundles": `<prefix>.index` + `<prefix>.data-00000-of-00001`.

  .index  : an SSTable in the public LevelDB table format -- prefix-compressed key/value blocks with restart arrays, each
            block followed by a type byte and a masked CRC-32C, an index block, a 48-byte footer ending in the magic
            0xdb4775248b80fb57.  Key "" -> BundleHeaderProto, every other key -> BundleEntryProto {dtype, shape, shard_id,
            offset, size, crc32c}.
  .data-* : the raw little-endian tensor bytes at those offsets.

Keras addresses a weight as `layer_with_weights-<N>/<attr>/.ATTRIBUTES/VARIABLE_VALUE` (N from `model.layers` order, see
keras_graph.py), Adam slots as `<weight>/.OPTIMIZER_SLOT/optimizer/{m,v}/.ATTRIBUTES/VARIABLE_VALUE`, the step counter as
`optimizer/iter/...`; `_CHECKPOINTABLE_OBJECT_GRAPH` holds a TrackableObjectGraph proto that names every variable
(`full_name`, e.g. "hg0_conv_1x1_1/kernel").  The reader resolves weights through that graph when it is present and through
the layer order otherwise, and insists that both agree.

TensorFlow is absent from the build image, so no file written by real Keras is available.  What IS pinned
(tests/test_cpu_format_pins.py, against TensorFlow's own code shipped inside tensorboard): the bundle header / entry protos
(compiled TensorShapeProto, DataType, VersionDef), the masked CRC-32C, the string-tensor encoding and the TrackableObjectGraph
proto in both directions; the key naming is pinned through the reference's own saved `model.summary()` (keras_graph.py).
The LevelDB table container is restated from the published format only.  CRCs run in libhgb200 (`hgb_crc32c`).
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

from . import keras_graph
from .tfrecord import _fields, _len_delimited, _put_varint, _varint
from ._lib import lib

_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_STRING, DT_INT64 = 1, 2, 3, 7, 9
_NP = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DT = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64}
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def _mask(crc):
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def _crc(data):
    return lib.hgb_crc32c(bytes(data), len(data))


# ------------------------------------------------------------------ LevelDB table: reading
def _read_block(buf, offset, size):
    block = buf[offset:offset + size]
    kind = buf[offset + size]
    (stored,) = struct.unpack("<I", buf[offset + size + 1:offset + size + 5])
    if _mask(_crc(buf[offset:offset + size + 1])) != stored:
        raise ValueError("checkpoint index: block checksum mismatch")
    if kind != 0:
        raise ValueError("checkpoint index: compressed table blocks are not supported (TF writes bundles uncompressed)")
    (num_restarts,) = struct.unpack("<I", block[-4:])
    end = len(block) - 4 * (num_restarts + 1)
    at, key, out = 0, b"", []
    while at < end:
        shared, at = _varint(block, at)
        non_shared, at = _varint(block, at)
        value_len, at = _varint(block, at)
        key = key[:shared] + bytes(block[at:at + non_shared])
        at += non_shared
        out.append((key, bytes(block[at:at + value_len])))
        at += value_len
    return out


def _read_table(path):
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != _MAGIC:
        raise ValueError(f"{path}: not a TensorFlow checkpoint index (bad table magic)")
    footer = buf[-48:]
    _meta_off, at = _varint(footer, 0)
    _meta_size, at = _varint(footer, at)
    index_off, at = _varint(footer, at)
    index_size, at = _varint(footer, at)
    entries = OrderedDict()
    for _sep, handle in _read_block(buf, index_off, index_size):
        off, at = _varint(handle, 0)
        size, at = _varint(handle, at)
        for k, v in _read_block(buf, off, size):
            entries[k] = v
    return entries


# ------------------------------------------------------------------ LevelDB table: writing
class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", restart_interval

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            limit = min(len(key), len(self.last))
            while shared < limit and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _write_table(path, entries, block_size=262144):
    """entries: iterable of (key bytes, value bytes) in ascending key order."""
    out, index, block, last_key = bytearray(), _BlockBuilder(1), _BlockBuilder(), None

    def emit(contents):
        handle = _put_varint(len(out)) + _put_varint(len(contents))
        out.extend(contents + b"\x00" + struct.pack("<I", _mask(_crc(contents + b"\x00"))))
        return handle

    for key, value in entries:
        if last_key is not None and key <= last_key:
            raise ValueError("table keys must be strictly ascending")
        block.add(key, value)
        last_key = key
        if block.size() >= block_size:
            index.add(last_key, emit(block.finish()))
            block = _BlockBuilder()
    if block.buf or last_key is None:
        index.add(last_key or b"", emit(block.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    out.extend(footer + bytes(40 - len(footer)) + struct.pack("<Q", _MAGIC))
    with open(path, "wb") as f:
        f.write(bytes(out))


# ------------------------------------------------------------------ bundle protos
def _parse_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0}
    for n, w, v in _fields(memoryview(buf)):
        if n == 1:
            e["dtype"] = v
        elif n == 2:
            for sn, _sw, sv in _fields(v):
                if sn == 2:
                    dim = 0
                    for dn, _dw, dv in _fields(sv):
                        if dn == 1:
                            dim = dv
                    e["shape"].append(dim)
        elif n == 3:
            e["shard_id"] = v
        elif n == 4:
            e["offset"] = v
        elif n == 5:
            e["size"] = v
        elif n == 6:
            (e["crc32c"],) = struct.unpack("<I", v)
        elif n == 7:
            raise ValueError("sliced (partitioned) checkpoint tensors are not supported")
    return e


def _build_entry(dtype, shape, offset, size, crc):
    dims = b"".join(_len_delimited(2, _put_varint(1 << 3) + _put_varint(int(d))) for d in shape)
    out = _put_varint(1 << 3) + _put_varint(dtype) + _len_delimited(2, dims)
    if offset:
        out += _put_varint(4 << 3) + _put_varint(offset)
    out += _put_varint(5 << 3) + _put_varint(size) + _put_varint((6 << 3) | 5) + struct.pack("<I", crc)
    return out


def _header(num_shards=1):
    version = _put_varint(1 << 3) + _put_varint(1)                       # VersionDef.producer = 1 (kTensorBundleVersion)
    return _put_varint(1 << 3) + _put_varint(num_shards) + _len_delimited(3, version)   # endianness LITTLE (0) is the default


def _encode_string_scalar(value: bytes):
    """DT_STRING tensor with one element: varint length | masked CRC of the fixed-width length | bytes."""
    length32 = struct.pack("<I", len(value))
    checksum = struct.pack("<I", _mask(_crc(length32)))
    data = _put_varint(len(value)) + checksum + value
    return data, _mask(_crc(length32 + checksum + value))


def _decode_strings(data, count):
    at, lens = 0, []
    for _ in range(count):
        n, at = _varint(data, at)
        lens.append(n)
    at += 4
    out = []
    for n in lens:
        out.append(bytes(data[at:at + n]))
        at += n
    return out


# ------------------------------------------------------------------ checkpoint read / write
def read_checkpoint(prefix, verify=True):
    """-> OrderedDict {key: ndarray (numeric) | bytes / list[bytes] (DT_STRING)} of every tensor in the bundle."""
    table = _read_table(prefix + ".index")
    header = table.get(b"")
    num_shards = 1
    if header is not None:
        for n, _w, v in _fields(memoryview(header)):
            if n == 1:
                num_shards = v
            if n == 2 and v != 0:
                raise ValueError("big-endian checkpoints are not supported")
    shards = {}
    out = OrderedDict()
    for key, value in table.items():
        if key == b"":
            continue
        e = _parse_entry(value)
        if e["shard_id"] not in shards:
            shards[e["shard_id"]] = np.memmap(f"{prefix}.data-{e['shard_id']:05d}-of-{num_shards:05d}", dtype=np.uint8, mode="r")
        raw = shards[e["shard_id"]][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise ValueError(f"{prefix}: tensor {key.decode()} extends past the end of its data shard")
        name = key.decode("utf-8")
        if e["dtype"] == DT_STRING:
            count = int(np.prod(e["shape"])) if e["shape"] else 1
            strings = _decode_strings(bytes(raw), count)
            out[name] = strings[0] if not e["shape"] else strings
            continue
        if e["dtype"] not in _NP:
            raise ValueError(f"{prefix}: tensor {name} has unsupported dtype enum {e['dtype']}")
        if verify and _mask(_crc(raw.tobytes())) != e["crc32c"]:
            raise ValueError(f"{prefix}: checksum mismatch in tensor {name}")
        out[name] = np.frombuffer(raw.tobytes(), dtype=_NP[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def write_checkpoint(prefix, tensors):
    """tensors: {key: ndarray | bytes}.  Writes `<prefix>.index` and `<prefix>.data-00000-of-00001` (tensors laid out in key
    order, as BundleWriter does)."""
    data, entries = bytearray(), [(b"", _header())]
    for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
        v = tensors[name]
        if isinstance(v, (bytes, bytearray)):
            blob, crc = _encode_string_scalar(bytes(v))
            entry = _build_entry(DT_STRING, (), len(data), len(blob), crc)
        else:
            arr = np.asarray(v)
            if not arr.flags.c_contiguous:
                arr = arr.copy()
            if arr.dtype not in _DT:
                raise TypeError(f"{name}: unsupported dtype {arr.dtype}")
            blob = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
            entry = _build_entry(_DT[arr.dtype], arr.shape, len(data), len(blob), _mask(_crc(blob)))
        entries.append((name.encode("utf-8"), entry))
        data += blob
    # like BundleWriter: both files are written under temporary names and renamed, data first, so a crash or a concurrent
    # reader (ModelCheckpoint runs every epoch) never sees an index that points into a different data file
    with open(prefix + ".data-00000-of-00001.tmp", "wb") as f:
        f.write(bytes(data))
    _write_table(prefix + ".index.tmp", entries)
    os.replace(prefix + ".data-00000-of-00001.tmp", prefix + ".data-00000-of-00001")
    os.replace(prefix + ".index.tmp", prefix + ".index")


# ------------------------------------------------------------------ TrackableObjectGraph
def _object_graph(layer_weights, with_optimizer):
    """Minimal TrackableObjectGraph: root -> layer_with_weights-N -> <attr> variable nodes carrying
    (name='VARIABLE_VALUE', full_name='<layer>/<attr>', checkpoint_key); optimizer node with iter and the m / v slots."""
    nodes = [[[], [], []]]                                              # per node: children, attributes, slot refs

    def new_node():
        nodes.append([[], [], []])
        return len(nodes) - 1

    var_node = {}
    for n, (layer_name, attrs) in enumerate(layer_weights):
        ln = new_node()
        nodes[0][0].append((ln, f"layer_with_weights-{n}"))
        for attr in attrs:
            vn = new_node()
            nodes[ln][0].append((vn, attr))
            nodes[vn][1].append(("VARIABLE_VALUE", f"{layer_name}/{attr}", f"layer_with_weights-{n}/{attr}{_SUFFIX}"))
            var_node[f"layer_with_weights-{n}/{attr}"] = (vn, f"{layer_name}/{attr}")
    if with_optimizer:
        on = new_node()
        nodes[0][0].append((on, "optimizer"))
        it = new_node()
        nodes[on][0].append((it, "iter"))
        nodes[it][1].append(("VARIABLE_VALUE", "Adam/iter", f"optimizer/iter{_SUFFIX}"))
        for key, (vn, full) in var_node.items():
            if key.rsplit("/", 1)[1].startswith("moving_"):
                continue
            for slot in ("m", "v"):
                sn = new_node()
                nodes[sn][1].append(("VARIABLE_VALUE", f"Adam/{full}/{slot}", f"{key}/.OPTIMIZER_SLOT/optimizer/{slot}{_SUFFIX}"))
                nodes[on][2].append((vn, slot, sn))
    out = b""
    for children, attributes, slots in nodes:
        body = b""
        for node_id, local_name in children:
            body += _len_delimited(1, _put_varint(1 << 3) + _put_varint(node_id) + _len_delimited(2, local_name.encode()))
        for name, full_name, key in attributes:
            body += _len_delimited(2, _len_delimited(1, name.encode()) + _len_delimited(2, full_name.encode()) + _len_delimited(3, key.encode()))
        for orig, slot_name, slot_node in slots:
            body += _len_delimited(3, _put_varint(1 << 3) + _put_varint(orig) + _len_delimited(2, slot_name.encode()) +
                                   _put_varint(3 << 3) + _put_varint(slot_node))
        out += _len_delimited(1, body)
    return out


def parse_object_graph(blob):
    """-> {checkpoint_key: full_name} of every serialized variable in a TrackableObjectGraph."""
    out = {}
    for n, _w, node in _fields(memoryview(blob)):
        if n != 1:
            continue
        for fn, _fw, fv in _fields(node):
            if fn != 2:
                continue
            full, key = None, None
            for an, _aw, av in _fields(fv):
                if an == 2:
                    full = bytes(av).decode()
                elif an == 3:
                    key = bytes(av).decode()
            if key is not None:
                out[key] = full
    return out


# ------------------------------------------------------------------ Keras weights <-> checkpoint
def _layer_weights(model):
    keys = keras_graph.checkpoint_keys(model.num_classes, model.num_stacks, model.num_channels, mobile=getattr(model, "mobile", False))
    if set(keys) != set(model._table):
        raise RuntimeError("keras_graph and the library's parameter table disagree on the weight names")
    layers = OrderedDict()
    for ours, key in keys.items():
        layer, attr = ours.rsplit("/", 1)
        layers.setdefault(layer, []).append(attr)
    return keys, list(layers.items())


def save_keras_weights(model, prefix, weights=None, adam=None):
    """`model.save_weights(prefix)` in TF format: weights under Keras' object-graph keys (+ Adam step / m / v slots when
    `adam = (iterations, m_dict, v_dict)` is given) and the object graph that names them."""
    keys, layer_weights = _layer_weights(model)
    weights = weights if weights is not None else model.get_weights_dict()
    tensors = {key: np.asarray(weights[ours], np.float32) for ours, key in keys.items()}
    if adam is not None:
        iterations, m, v = adam
        tensors[f"optimizer/iter{_SUFFIX}"] = np.asarray(iterations, np.int64)
        for ours, key in keys.items():
            if ours in m:
                base = key[:-len(_SUFFIX)]
                tensors[f"{base}/.OPTIMIZER_SLOT/optimizer/m{_SUFFIX}"] = np.asarray(m[ours], np.float32)
                tensors[f"{base}/.OPTIMIZER_SLOT/optimizer/v{_SUFFIX}"] = np.asarray(v[ours], np.float32)
    tensors[OBJECT_GRAPH_KEY] = _object_graph(layer_weights, adam is not None)
    write_checkpoint(prefix, tensors)


def load_keras_weights(model, prefix):
    """Read a checkpoint written by the reference (Keras `save_weights` in TF format, or a SavedModel's variables/variables)
    -> (weights {our name: array}, adam (iterations, m, v) or None).  Every weight of the model must be present with the
    right shape; the object graph's variable names (when present) must agree with the layer-order addressing."""
    keys, _ = _layer_weights(model)
    bundle = read_checkpoint(prefix)
    named = parse_object_graph(bundle[OBJECT_GRAPH_KEY]) if isinstance(bundle.get(OBJECT_GRAPH_KEY), bytes) else {}
    first = next(iter(keys.values()))
    if first not in bundle:
        # the model saved as an attribute of another object (tf.train.Checkpoint(model=m).save(...), wrappers): every key
        # carries that attribute path in front, e.g. "model/layer_with_weights-0/kernel/..."
        roots = sorted({k[:-len(first)] for k in bundle if k.endswith("/" + first)})
        if len(roots) == 1:
            root = roots[0]
            bundle = OrderedDict((k[len(root):] if k.startswith(root) else k, v) for k, v in bundle.items())
            named = {(k[len(root):] if k.startswith(root) else k): v for k, v in named.items()}
    weights, m, v = OrderedDict(), {}, {}
    for ours, key in keys.items():
        if key not in bundle:
            raise KeyError(f"{prefix}: missing {key} ({ours}); the checkpoint was written by a different architecture")
        full = named.get(key)
        if full is not None and full.split(":")[0] != ours:
            raise ValueError(f"{prefix}: {key} is variable {full!r} in the checkpoint's object graph but {ours!r} in this model")
        arr = bundle[key]
        shape = tuple(model._table[ours][0])
        if tuple(arr.shape) != shape:
            raise ValueError(f"{prefix}: {ours} has shape {tuple(arr.shape)}, expected {shape}")
        weights[ours] = arr.astype(np.float32)
        base = key[:-len(_SUFFIX)]
        mk, vk = f"{base}/.OPTIMIZER_SLOT/optimizer/m{_SUFFIX}", f"{base}/.OPTIMIZER_SLOT/optimizer/v{_SUFFIX}"
        if mk in bundle and vk in bundle:
            m[ours], v[ours] = bundle[mk].astype(np.float32), bundle[vk].astype(np.float32)
    adam = None
    it_key = f"optimizer/iter{_SUFFIX}"
    if it_key in bundle and m:
        # a checkpoint may hold slots for only some variables (a frozen layer, an optimizer rebuilt mid-run): the missing
        # moments start at zero, exactly what Keras does when it creates a slot
        for ours, (shape, _off, trainable) in model._table.items():
            if trainable and ours not in m:
                m[ours], v[ours] = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
        adam = (int(bundle[it_key]), m, v)
    return weights, adam
