"""TFRecord files and tf.train.Example messages without TensorFlow.

The reference stores its dataset as TFRecords of tf.train.Example (gen_tfrecords.py:12-86 writes them,
dataset_builder.py:241-268 parses them).  Both formats are public and tiny:

  record  := uint64 length | uint32 masked_crc32c(length) | payload | uint32 masked_crc32c(payload)        (little endian)
  Example := { 1: Features { 1: map<string, Feature> } },  Feature := { 1: BytesList | 2: FloatList | 3: Int64List },
             each list := { 1: repeated value } (floats / int64s packed, unpacked accepted)

Pinned against TensorFlow's own Python record writer / reader (shipped in tensorboard) and against the protobuf runtime's
serialization of the published schema: files and messages are byte-identical (tests/test_cpu_format_pins.py).
The CRC and the Example walk run in libhgb200 (`hgb_crc32c`, `hgb_example_parse`, host code); the rest is byte bookkeeping.  JPEG payloads are decoded on
the GPU (`decode_jpeg_batch` -> `hgb_jpeg_decode`).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import threading

import numpy as np

from . import _lib
from ._lib import check, lib

_MASK_DELTA = 0xA282EAD8
_lib_ERR_STATE = -3


def masked_crc32c(data: bytes) -> int:
    crc = lib.hgb_crc32c(data, len(data))
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------ record framing
def read_records(path, verify=True):
    """Yield the payload of every record of one .tfrec file; corrupt framing raises ValueError like TF's DataLossError.
    The file is memory-mapped and its framing walked (and checksummed) in libhgb200, 1024 records per call."""
    import mmap
    if os.path.getsize(path) == 0:
        return
    with open(path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as mm:
        view = np.frombuffer(mm, dtype=np.uint8)
        try:
            offsets, lengths = np.empty(1024, np.int64), np.empty(1024, np.int64)
            nxt, at, total = C.c_int64(0), 0, len(view)
            while at < total:
                n = lib.hgb_tfrecord_scan(view.ctypes.data, total, at, int(bool(verify)), offsets.ctypes.data, lengths.ctypes.data,
                                          offsets.size, C.byref(nxt))
                if n < 0:
                    raise ValueError(f"{path}: {lib.hgb_last_error().decode('utf-8', 'replace')}")
                for off, size in zip(offsets[:n].tolist(), lengths[:n].tolist()):
                    yield mm[off:off + size]
                at = nxt.value
        finally:
            del view                                         # release the buffer export before the map closes


def write_records(path, payloads):
    """tf.io.TFRecordWriter (gen_tfrecords.py:106-114)."""
    with open(path, "wb") as f:
        for payload in payloads:
            head = struct.pack("<Q", len(payload))
            f.write(head + struct.pack("<I", masked_crc32c(head)) + payload + struct.pack("<I", masked_crc32c(payload)))


# ------------------------------------------------------------------ protobuf wire format (the subset Example uses)
def _varint(buf, at):
    value = shift = 0
    while True:
        b = buf[at]
        at += 1
        value |= (b & 0x7F) << shift
        if not b & 0x80:
            return value, at
        shift += 7


def _put_varint(value):
    value &= 0xFFFFFFFFFFFFFFFF
    out = bytearray()
    while True:
        b = value & 0x7F
        value >>= 7
        out.append(b | (0x80 if value else 0))
        if not value:
            return bytes(out)


def _fields(buf):
    at, end = 0, len(buf)
    while at < end:
        key, at = _varint(buf, at)
        number, wire = key >> 3, key & 7
        if wire == 0:
            value, at = _varint(buf, at)
        elif wire == 1:
            value, at = buf[at:at + 8], at + 8
        elif wire == 2:
            size, at = _varint(buf, at)
            value, at = buf[at:at + size], at + size
        elif wire == 5:
            value, at = buf[at:at + 4], at + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wire}")
        if at > end:
            raise ValueError("truncated protobuf field")
        yield number, wire, value


def _len_delimited(number, payload):
    return _put_varint((number << 3) | 2) + _put_varint(len(payload)) + payload


def _parse_feature(buf):
    for number, wire, value in _fields(buf):
        if wire != 2:
            continue
        if number == 1:                                                    # BytesList
            return [bytes(v) for n, w, v in _fields(value) if n == 1 and w == 2]
        if number == 2:                                                    # FloatList
            out = []
            for n, w, v in _fields(value):
                if n == 1 and w in (2, 5):
                    if len(v) % 4:
                        raise ValueError("malformed FloatList")
                    out.append(np.frombuffer(v, dtype="<f4"))
            return np.concatenate(out).astype(np.float32) if out else np.zeros(0, np.float32)
        if number == 3:                                                    # Int64List
            out = []
            for n, w, v in _fields(value):
                if n != 1:
                    continue
                if w == 0:
                    out.append(v)
                elif w == 2:
                    at = 0
                    while at < len(v):
                        x, at = _varint(v, at)
                        out.append(x)
            arr = np.array(out, dtype=np.uint64).astype(np.int64) if out else np.zeros(0, np.int64)
            return arr
    return None                                                            # Feature with no list set


def _parse_example_py(payload: bytes) -> dict:
    """Pure-Python walk of the wire format (the definition the native parser is tested against)."""
    out = {}
    try:
        for number, wire, features in _fields(memoryview(payload)):
            if number != 1 or wire != 2:
                continue
            for n, w, entry in _fields(features):
                if n != 1 or w != 2:
                    continue
                name, feature = None, None
                for en, ew, ev in _fields(entry):
                    if en == 1 and ew == 2:
                        name = bytes(ev).decode("utf-8")
                    elif en == 2 and ew == 2:
                        feature = _parse_feature(ev)
                if name is not None:
                    out[name] = feature
    except (IndexError, TypeError, struct.error) as ex:
        raise ValueError("malformed tf.train.Example") from ex
    return out


_scratch = threading.local()


def parse_example(payload: bytes) -> dict:
    """Serialized tf.train.Example -> {name: list[bytes] | float32 array | int64 array | None}.  The walk runs in libhgb200
    (`hgb_example_parse`); examples beyond its fixed capacities or with multi-valued bytes features take the Python path."""
    sc = getattr(_scratch, "bufs", None)
    if sc is None:
        sc = _scratch.bufs = (np.empty((64, 6), np.int64), np.empty(4096, np.float32), np.empty(4096, np.int64))
    table, fvals, ivals = sc
    n = lib.hgb_example_parse(payload, len(payload), table.shape[0], table.ctypes.data, fvals.ctypes.data, fvals.size,
                              ivals.ctypes.data, ivals.size)
    if n == _lib_ERR_STATE:
        return _parse_example_py(payload)
    check(n if n < 0 else 0)
    out = {}
    for name_off, name_len, kind, start, count, extra in table[:n].tolist():
        name = payload[name_off:name_off + name_len].decode("utf-8")
        if kind == 1:
            if count != 1:
                return _parse_example_py(payload)
            out[name] = [payload[start:start + extra]]
        elif kind == 2:
            out[name] = fvals[start:start + count].copy()
        elif kind == 3:
            out[name] = ivals[start:start + count].copy()
        else:
            out[name] = None
    return out


def build_example(features: dict) -> bytes:
    """{name: bytes | str | list[bytes] | float array | int array} -> serialized tf.train.Example (map entries sorted by
    name, the deterministic order protobuf serialisation uses)."""
    entries = b""
    for name in sorted(features):
        v = features[name]
        if isinstance(v, str):
            v = v.encode()
        if isinstance(v, (bytes, bytearray)):
            v = [bytes(v)]
        if isinstance(v, list) and v and isinstance(v[0], (bytes, bytearray)):
            feature = _len_delimited(1, b"".join(_len_delimited(1, bytes(b)) for b in v))
        else:
            arr = np.atleast_1d(np.asarray(v))
            if arr.dtype.kind == "f":
                body = _len_delimited(1, arr.astype("<f4").tobytes()) if arr.size else b""
                feature = _len_delimited(2, body)
            elif arr.dtype.kind in "iub":
                body = _len_delimited(1, b"".join(_put_varint(int(x)) for x in arr)) if arr.size else b""
                feature = _len_delimited(3, body)
            else:
                raise TypeError(f"feature {name!r}: unsupported value type {arr.dtype}")
        entries += _len_delimited(1, _len_delimited(1, name.encode()) + _len_delimited(2, feature))
    return _len_delimited(1, entries)


# ------------------------------------------------------------------ the reference's schema
FIXED_INT = ("ann_id", "image_id", "width", "height", "keypoints/num")
FIXED_BYTES = ("image", "image_path", "coco_url")
FIXED_FLOAT = ("bbox_x", "bbox_y")
VAR_FLOAT = ("keypoints/x", "keypoints/y", "original_bbox")
VAR_INT = ("keypoints/vis",)


def parse_tfrecord_fn(payload: bytes) -> dict:
    """DatasetBuilder.parse_tfrecord_fn (dataset_builder.py:241-268) minus the image decode (done on the GPU in batches):
    FixedLenFeature scalars, VarLenFeature -> dense arrays; `image` stays the encoded JPEG bytes."""
    raw = parse_example(payload)
    ex = {}
    for name in FIXED_INT + FIXED_BYTES + FIXED_FLOAT:
        v = raw.get(name)
        if v is None or len(v) != 1:
            raise ValueError(f"Feature: {name} (data type: {'string' if name in FIXED_BYTES else 'number'}) is required but could not be found.")
        ex[name] = v[0] if name in FIXED_BYTES else (int(v[0]) if name in FIXED_INT else np.float32(v[0]))
    for name in VAR_FLOAT:
        v = raw.get(name)
        ex[name] = np.zeros(0, np.float32) if v is None else np.asarray(v, np.float32)
    for name in VAR_INT:
        v = raw.get(name)
        ex[name] = np.zeros(0, np.int64) if v is None else np.asarray(v, np.int64)
    return ex


# ------------------------------------------------------------------ JPEG
def jpeg_info(data: bytes):
    h, w, c = C.c_int(), C.c_int(), C.c_int()
    check(lib.hgb_jpeg_info(data, len(data), C.byref(h), C.byref(w), C.byref(c)))
    return h.value, w.value, c.value


def decode_jpeg_batch(streams):
    """List of encoded JPEG byte strings -> list of (h,w,3) uint8 CUDA tensors (RGB), decoded by nvJPEG on the current stream."""
    torch = _lib.require_cuda()
    n = len(streams)
    if n == 0:
        return []
    sizes = [jpeg_info(s)[:2] for s in streams]
    outs = [torch.empty((h, w, 3), dtype=torch.uint8, device="cuda") for h, w in sizes]
    keep = [C.create_string_buffer(s, len(s)) for s in streams]
    datas = (C.c_void_p * n)(*[C.cast(b, C.c_void_p) for b in keep])
    lens = (C.c_int64 * n)(*[len(s) for s in streams])
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
    hw = (C.c_int32 * (2 * n))(*[v for s in sizes for v in s])
    check(lib.hgb_jpeg_decode(datas, lens, n, ptrs, hw, _lib.stream_ptr()))
    torch.cuda.current_stream().synchronize()            # nvJPEG reads the host streams asynchronously: keep them until done
    return outs
