"""TFRecord framing / tf.train.Example wire format / JPEG header parsing (tfrecord.py, csrc/io_host.cu) and the
TFRecord-backed DatasetBuilder's host logic.  CPU only: no decode, no kernels."""
import struct
import types

import cv2
import numpy as np
import pytest


@pytest.fixture(scope="module")
def hgb():
    import hgb200
    return hgb200


def test_crc32c_known_answers(hgb):
    from hgb200 import tfrecord
    crc = hgb._lib.lib.hgb_crc32c
    assert crc(b"", 0) == 0
    assert crc(b"123456789", 9) == 0xE3069283                      # CRC-32C check value
    assert crc(bytes(32), 32) == 0x8A9136AA                        # RFC 3720 B.4: 32 bytes of zeros
    assert crc(bytes([0xFF] * 32), 32) == 0x62A8AB43               # RFC 3720 B.4: 32 bytes of ones
    assert crc(bytes(range(32)), 32) == 0x46DD794E                 # RFC 3720 B.4: incrementing
    data = np.random.default_rng(0).integers(0, 256, 10007, dtype=np.uint8).tobytes()
    ref = 0xFFFFFFFF
    for b in data[:999]:                                           # bitwise definition on a prefix with an odd length
        ref ^= b
        for _ in range(8):
            ref = (ref >> 1) ^ 0x82F63B78 if ref & 1 else ref >> 1
    assert crc(data, 999) == ref ^ 0xFFFFFFFF
    assert crc(data[1:], 998) != crc(data, 998)                    # unaligned start goes through the byte prologue
    # the dispatched routine (CRC32 instruction when the CPU has it) against the portable table walk, at every alignment
    import ctypes
    buf = ctypes.create_string_buffer(data, len(data))
    portable = hgb._lib.lib.hgb_crc32c_portable
    for off in range(9):
        for n in (0, 1, 7, 8, 9, 31, 64, 1001, len(data) - 8):
            ptr = ctypes.cast(ctypes.byref(buf, off), ctypes.c_void_p)
            assert crc(ptr, n) == portable(ptr, n), (off, n)
    assert portable(b"123456789", 9) == 0xE3069283
    masked = tfrecord.masked_crc32c(b"123456789")
    assert masked == (((0xE3069283 >> 15) | (0xE3069283 << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_example_wire_format_known_answer(hgb):
    from hgb200 import tfrecord
    # tf.train.Example(features=Features(feature={'a': Feature(int64_list=Int64List(value=[1]))})).SerializeToString()
    want = b"\n\x0c\n\n\n\x01a\x12\x05\x1a\x03\n\x01\x01"
    assert tfrecord.build_example({"a": np.array([1])}) == want
    assert tfrecord.parse_example(want)["a"].tolist() == [1]
    # unpacked repeated scalars (what old writers emit) parse to the same arrays
    int64_list = b"\x08\x01" + b"\x08\xff\xff\xff\xff\x0f"
    feature = b"\x1a" + bytes([len(int64_list)]) + int64_list
    entry = b"\n\x01a" + b"\x12" + bytes([len(feature)]) + feature
    features = b"\n" + bytes([len(entry)]) + entry
    unpacked = b"\n" + bytes([len(features)]) + features
    assert tfrecord.parse_example(unpacked)["a"].tolist() == [1, 0xFFFFFFFF]


def _example(rng, k=17):
    img = (rng.random((40, 40, 3)) * 255).astype(np.uint8)
    ok, enc = cv2.imencode(".jpg", img)
    assert ok
    return {"ann_id": 1234567890123, "image_id": 42, "image": enc.tobytes(), "image_path": "dataset/images/x.jpg",
            "coco_url": "http://images.cocodataset.org/x.jpg", "width": 40, "height": 40,
            "keypoints/x": (rng.random(k) * 40).astype(np.float32), "keypoints/y": (rng.random(k) * 40).astype(np.float32),
            "keypoints/vis": rng.integers(0, 3, k), "keypoints/num": 9, "bbox_x": np.float32(-3.25), "bbox_y": np.float32(17.5),
            "original_bbox": np.array([1.5, 2.5, 30.0, 35.0], np.float32)}


def test_record_round_trip_and_corruption(hgb, tmp_path):
    from hgb200 import tfrecord
    rng = np.random.default_rng(1)
    exs = [_example(rng) for _ in range(5)]
    path = str(tmp_path / "file_train_00-5.tfrec")
    tfrecord.write_records(path, [tfrecord.build_example(e) for e in exs])
    back = [tfrecord.parse_tfrecord_fn(p) for p in tfrecord.read_records(path)]
    assert len(back) == 5
    for a, b in zip(exs, back):
        assert b["ann_id"] == a["ann_id"] and b["image_id"] == 42 and b["image"] == a["image"]
        assert b["coco_url"] == a["coco_url"].encode() and b["bbox_x"] == a["bbox_x"] and b["bbox_x"].dtype == np.float32
        np.testing.assert_array_equal(b["keypoints/x"], a["keypoints/x"])
        np.testing.assert_array_equal(b["keypoints/vis"], a["keypoints/vis"])
        assert b["keypoints/vis"].dtype == np.int64
        np.testing.assert_array_equal(b["original_bbox"], a["original_bbox"])
    # framing: length | masked crc | payload | masked crc
    raw = open(path, "rb").read()
    (n,) = struct.unpack("<Q", raw[:8])
    assert raw[12:12 + n] == tfrecord.build_example(exs[0])
    bad = bytearray(raw)
    bad[40] ^= 0x10
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="corrupted"):
        list(tfrecord.read_records(path))
    assert len(list(tfrecord.read_records(path, verify=False))) == 5
    open(path, "wb").write(raw[:-7])
    with pytest.raises(ValueError, match="truncated"):
        list(tfrecord.read_records(path))
    # a missing required feature is an error like tf.io.parse_single_example's
    e = dict(exs[0])
    del e["bbox_x"]
    with pytest.raises(ValueError, match="bbox_x"):
        tfrecord.parse_tfrecord_fn(tfrecord.build_example(e))
    # negative int64 and empty lists survive
    odd = tfrecord.parse_example(tfrecord.build_example({"n": np.array([-1, -2 ** 63, 7]), "e": np.zeros(0, np.float32)}))
    assert odd["n"].tolist() == [-1, -2 ** 63, 7] and odd["e"] is None or len(odd["e"]) == 0


def test_jpeg_header_parsing(hgb):
    from hgb200 import tfrecord
    rng = np.random.default_rng(2)
    img = (rng.random((37, 53, 3)) * 255).astype(np.uint8)
    for flags in ([], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1], [cv2.IMWRITE_JPEG_QUALITY, 30]):
        ok, enc = cv2.imencode(".jpg", img, flags)
        assert tfrecord.jpeg_info(enc.tobytes()) == (37, 53, 3)
    ok, enc = cv2.imencode(".jpg", img[:, :, 0])
    assert tfrecord.jpeg_info(enc.tobytes()) == (37, 53, 1)
    with pytest.raises(ValueError):
        tfrecord.jpeg_info(b"\x89PNG\r\n\x1a\n" + bytes(32))
    with pytest.raises(ValueError):
        tfrecord.jpeg_info(enc.tobytes()[:12])


def test_dataset_builder_host_contract(hgb, tmp_path, capsys):
    from hgb200 import tfrecord
    rng = np.random.default_rng(3)
    train, valid = tmp_path / "train", tmp_path / "valid"
    train.mkdir()
    valid.mkdir()
    for i, n in enumerate((3, 2, 4)):
        tfrecord.write_records(str(train / f"file_train_{i:02d}-{n}.tfrec"), [tfrecord.build_example(_example(rng)) for _ in range(n)])
    tfrecord.write_records(str(valid / "file_valid_00-2.tfrec"), [tfrecord.build_example(_example(rng)) for _ in range(2)])
    cfg = types.SimpleNamespace(**{k: getattr(hgb.default_config, k) for k in dir(hgb.default_config) if k.isupper()})
    cfg.TRAIN_TFRECORDS_DIR, cfg.VALID_TFRECORDS_DIR, cfg.BATCH_SIZE, cfg.SHUFFLE_BUFFER = str(train), str(valid), 4, 5
    b = hgb.dataset_builder.DatasetBuilder(cfg, seed=0)
    out = capsys.readouterr().out
    assert "Train dataset with 3 tfrecords and 9 examples." in out and "Valid dataset with 1 tfrecords and 2 examples." in out
    assert (b.num_train_examples, b.num_valid_examples, b.batch_size) == (9, 2, 4)
    half = hgb.dataset_builder.DatasetBuilder(cfg, ratio=0.5)
    assert len(half.train_filenames) == 2 and half.num_train_examples == 5
    with pytest.raises(AssertionError):
        hgb.dataset_builder.DatasetBuilder(cfg, ratio=0)
    # shuffle: a permutation of the pass; batching: 4 + 4 + 1 per pass
    recs = list(b._records(b.train_filenames))
    assert len(recs) == 9
    shuffled = list(b._shuffled(iter(recs)))
    assert sorted(shuffled) == sorted(recs) and shuffled != recs
    assert [len(x) for x in b._batches(iter(recs))] == [4, 4, 1]
    xs, ys, vs = b.flip_labels(np.arange(17.0), np.arange(17.0) + 100, np.arange(17), cfg.COCO_INDEX_FLIP_PAIRS)
    assert xs[:5].tolist() == [0, 2, 1, 4, 3] and ys[15:].tolist() == [116, 115] and vs[5:7].tolist() == [6, 5]


def test_dataset_builder_shards_records_across_ranks(hgb, tmp_path):
    """Data-parallel input: rank r of w reads records r, r+w, ... (no collective); the union over ranks is the pass."""
    from hgb200 import tfrecord
    rng = np.random.default_rng(4)
    d = tmp_path / "train"
    d.mkdir()
    (tmp_path / "valid").mkdir()
    tfrecord.write_records(str(d / "file_train_00-5.tfrec"), [tfrecord.build_example(_example(rng)) for _ in range(5)])
    tfrecord.write_records(str(d / "file_train_01-4.tfrec"), [tfrecord.build_example(_example(rng)) for _ in range(4)])
    cfg = types.SimpleNamespace(**{k: getattr(hgb.default_config, k) for k in dir(hgb.default_config) if k.isupper()})
    cfg.TRAIN_TFRECORDS_DIR, cfg.VALID_TFRECORDS_DIR, cfg.BATCH_SIZE = str(d), str(tmp_path / "valid"), 2
    whole = list(hgb.dataset_builder.DatasetBuilder(cfg)._records(sorted(str(p) for p in d.iterdir())))
    parts = []
    for rank in range(3):
        b = hgb.dataset_builder.DatasetBuilder(cfg, shard=(rank, 3))
        assert b.num_train_examples == 9                                  # global count
        parts.append(list(b._records(b.train_filenames)))
    assert [len(p) for p in parts] == [3, 3, 3]
    assert [parts[k % 3][k // 3] for k in range(9)] == whole
    with pytest.raises(ValueError):
        hgb.dataset_builder.DatasetBuilder(cfg, shard=(3, 3))


def test_prefetcher_order_errors_and_shutdown(hgb):
    import threading
    import time
    from hgb200.dataset_builder import Prefetcher
    assert list(Prefetcher(iter(range(50)), depth=3)) == list(range(50))          # order kept, finite generators end

    def boom():
        yield 1
        yield 2
        raise RuntimeError("corrupted record")
    p = Prefetcher(boom(), depth=2)
    assert next(p) == 1 and next(p) == 2
    with pytest.raises(RuntimeError, match="corrupted record"):
        next(p)
    with pytest.raises(StopIteration):
        next(p)

    produced = []

    def endless():
        k = 0
        while True:
            produced.append(k)
            yield {"images": k, "meta": [k, (k,)]}
            k += 1
    p = Prefetcher(endless(), depth=2)
    assert next(p)["images"] == 0
    time.sleep(0.3)
    assert len(produced) <= 4                                                    # runs at most `depth` (+ one in hand) ahead
    p.close()
    p._thread.join(timeout=2)
    assert not p._thread.is_alive()
    assert threading.active_count() >= 1


def test_native_example_parser_matches_the_python_walk(hgb):
    from hgb200 import tfrecord
    rng = np.random.default_rng(9)

    def same(payload):
        a, b = tfrecord.parse_example(payload), tfrecord._parse_example_py(payload)
        assert list(a) == list(b)
        for k in a:
            if isinstance(b[k], list) or b[k] is None:
                assert a[k] == b[k]
            else:
                assert a[k].dtype == b[k].dtype
                np.testing.assert_array_equal(a[k], b[k])
        return a

    for _ in range(20):
        ex = _example(rng, k=int(rng.integers(0, 40)))
        ex["neg"] = rng.integers(-2 ** 62, 2 ** 62, int(rng.integers(0, 9)))
        ex["blob"] = bytes(rng.integers(0, 256, int(rng.integers(0, 300)), dtype=np.uint8))
        got = same(tfrecord.build_example(ex))
        np.testing.assert_array_equal(got["neg"] if len(ex["neg"]) else np.zeros(0, np.int64), ex["neg"])
    # unpacked encodings, a feature with no list, and a multi-valued bytes feature (Python path)
    int64_list = b"\x08\x01" + b"\x08\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01"
    float_list = b"\x0d" + np.float32(1.5).tobytes() + b"\x0d" + np.float32(-2.0).tobytes()
    multi = b"\n\x02ab\n\x03cde"

    def entry(name, feature):
        e = b"\n" + bytes([len(name)]) + name + b"\x12" + bytes([len(feature)]) + feature
        return b"\n" + bytes([len(e)]) + e
    body = (entry(b"i", b"\x1a" + bytes([len(int64_list)]) + int64_list) + entry(b"f", b"\x12" + bytes([len(float_list)]) + float_list) +
            entry(b"none", b"") + entry(b"m", b"\n" + bytes([len(multi)]) + multi))
    payload = b"\n" + bytes([len(body)]) + body
    got = same(payload)
    assert got["i"].tolist() == [1, -1] and got["f"].tolist() == [1.5, -2.0] and got["none"] is None and got["m"] == [b"ab", b"cde"]
    # capacities: more values than the native scratch holds -> Python path, same result
    big = tfrecord.build_example({"x": np.arange(10000, dtype=np.float32), "y": np.arange(5000)})
    assert same(big)["x"].shape == (10000,)
    # malformed input is an error, not a crash
    for bad in (b"\n\xff", payload[:-3], b"\n\x05\n\x03\n\x05a"):
        with pytest.raises(ValueError):
            tfrecord.parse_example(bad)


def test_create_example_matches_the_reference_writer(hgb, golden_dir, tmp_path):
    """gen_tfrecords.py:12-86 executed by the reference itself (tests/golden/make_tfrecord_golden.py) vs hgb200.gen_tfrecords:
    every feature -- names, types, values, the crop bytes -- with the crop taken from the numpy oracle here (the device crop is
    checked against the same oracle in test_gpu_input.py) and an identity encoder."""
    import json
    import os
    import pandas as pd
    from hgb200 import gen_tfrecords, tfrecord
    from oracle import input_oracle as iorc
    g = np.load(os.path.join(golden_dir, "tfrecord_golden.npz"))
    image, rows = g["image"], json.loads(str(g["rows"]))
    crop_fn = lambda img, bbox: iorc.crop_and_pad(img, bbox)          # noqa: E731
    encode_fn = lambda crop: b"RAW" + crop.astype(np.uint8).tobytes()  # noqa: E731
    for i, row in enumerate(rows):
        for scale in (1.25, 1):
            payload = gen_tfrecords.create_example(image, f"dataset/images/val2017/{i}.jpg", row, 5000 + i, scale, crop_fn, encode_fn)
            got = tfrecord.parse_example(payload)
            names = [k[len(f"ex{i}_{scale}_"):] for k in g.files if k.startswith(f"ex{i}_{scale}_")]
            assert sorted(got) == sorted(names) and len(names) == 14
            for name in names:
                want = g[f"ex{i}_{scale}_{name}"]
                if want.dtype == np.uint8:
                    assert got[name] == [want.tobytes()], name
                elif want.dtype == np.int64:
                    assert got[name].dtype == np.int64 and got[name].tolist() == want.tolist(), name
                else:
                    assert got[name].dtype == np.float32
                    np.testing.assert_array_equal(got[name], want.astype(np.float32), err_msg=name)   # what FloatList stores
            ex = tfrecord.parse_tfrecord_fn(payload)                  # and the reader's schema accepts it
            assert ex["width"] == ex["height"] and len(ex["keypoints/x"]) == 17
    # the uint8 round trip of the device crop is exact: rint(float32(u) * float32(1/255) * 255) == u
    u = np.arange(256, dtype=np.float32)
    assert np.array_equal(np.rint((u * np.float32(1.0 / 255)).astype(np.float32) * np.float32(255.0)), u)
    # shard naming / counts (gen_tfrecords.py:88-115)
    df = pd.DataFrame([{**r, "image_path": "a.jpg"} for r in rows], index=[11, 12, 13, 14])
    cfg = types.SimpleNamespace(NUM_EXAMPLER_PER_TFRECORD=3, TRAIN_TFRECORDS_DIR=str(tmp_path / "tfrecords" / "train"),
                                VALID_TFRECORDS_DIR=str(tmp_path / "tfrecords" / "valid"), TRAIN_IMAGES_DIR="imgs", VALID_IMAGES_DIR="imgs", BBOX_SCALE=1.25)
    real_create = gen_tfrecords.create_example
    gen_tfrecords.create_example = lambda im, p, r, idx, s: real_create(im, p, r, idx, s, crop_fn, encode_fn)
    try:
        gen_tfrecords.gen_TFRecords(df, cfg, True, read_image=lambda path: image)
    finally:
        gen_tfrecords.create_example = real_create
    files = sorted(os.listdir(cfg.TRAIN_TFRECORDS_DIR))
    assert files == ["file_train_00-3.tfrec", "file_train_01-1.tfrec"]
    assert hgb.dataset_builder.DatasetBuilder.get_ds_length([os.path.join(cfg.TRAIN_TFRECORDS_DIR, f) for f in files]) == 4
    first = [tfrecord.parse_tfrecord_fn(p) for p in tfrecord.read_records(os.path.join(cfg.TRAIN_TFRECORDS_DIR, files[0]))]
    assert [e["image_id"] for e in first] == [11, 12, 13] and first[0]["ann_id"] == 900000


def test_coco_dataframe_matches_reference_semantics_and_feeds_the_writer(hgb, tmp_path, capsys):
    """coco_df.py:6-82 without pycocotools: filtering rules, index, columns; then annotations -> dataframe -> TFRecords ->
    DatasetBuilder bookkeeping as one chain."""
    import json
    from hgb200 import coco_df, gen_tfrecords, tfrecord
    from oracle import input_oracle as iorc
    rng = np.random.default_rng(12)

    def dataset(n_images, first_ann):
        images, anns = [], []
        for i in range(n_images):
            images.append({"id": 100 + i, "file_name": f"{100 + i:012d}.jpg", "width": 200, "height": 150, "coco_url": f"http://x/{i}.jpg"})
            for p in range(int(rng.integers(0, 4))):
                nk = int(rng.integers(0, 12))
                kps = []
                for k in range(17):
                    kps += [float(rng.integers(20, 120)), float(rng.integers(20, 120)), 2 if k < nk else 0]
                anns.append({"id": first_ann + len(anns), "image_id": 100 + i, "category_id": 1, "iscrowd": int(rng.random() < 0.2),
                             "bbox": [10.0, 12.0, 110.0, 120.0], "num_keypoints": nk, "keypoints": kps, "area": 100.0})
        return {"images": images, "annotations": anns, "categories": [{"id": 1, "name": "person"}]}

    train, valid = dataset(12, 1), dataset(6, 1000)
    (tmp_path / "train.json").write_text(json.dumps(train))
    (tmp_path / "valid.json").write_text(json.dumps(valid))
    cfg = types.SimpleNamespace(**{k: getattr(hgb.default_config, k) for k in dir(hgb.default_config) if k.isupper()})
    cfg.TRAIN_ANNOT_FILE, cfg.VALID_ANNOT_FILE = str(tmp_path / "train.json"), str(tmp_path / "valid.json")
    train_df, valid_df = coco_df.gen_trainval_df(cfg, drop_min_num_kps=True)
    out = capsys.readouterr().out
    assert "num_keypoints >= 5 are chosen" in out and f"Length of train df: {len(train_df)}" in out
    want = [a for a in train["annotations"] if a["iscrowd"] == 0 and a["num_keypoints"] >= 5]
    assert sorted(train_df["ann_id"]) == sorted(a["id"] for a in want) and len(want) > 0
    assert train_df.index.name == "image_id" and set(train_df.index) == {a["image_id"] for a in want}
    assert list(train_df.columns) == ["coco_url", "image_path", "width", "height", "ann_id", "is_crowd", "bbox", "num_keypoints", "keypoints"]
    loose, _ = coco_df.gen_trainval_df(cfg)
    assert len(loose) == sum(a["iscrowd"] == 0 and a["num_keypoints"] >= 1 for a in train["annotations"])
    # dataframe -> TFRecords -> reader
    cfg.TRAIN_TFRECORDS_DIR, cfg.VALID_TFRECORDS_DIR = str(tmp_path / "rec" / "train"), str(tmp_path / "rec" / "valid")
    cfg.NUM_EXAMPLER_PER_TFRECORD = 4
    frame = rng.integers(0, 256, (150, 200, 3), dtype=np.uint8)
    real_create = gen_tfrecords.create_example
    gen_tfrecords.create_example = lambda im, p, r, idx, s: real_create(im, p, r, idx, s, lambda a, b: iorc.crop_and_pad(a, b),
                                                                        lambda c: b"RAW" + c.tobytes())
    try:
        gen_tfrecords.gen_TFRecords(train_df, cfg, True, read_image=lambda path: frame)
        gen_tfrecords.gen_TFRecords(valid_df, cfg, False, read_image=lambda path: frame)
    finally:
        gen_tfrecords.create_example = real_create
    b = hgb.dataset_builder.DatasetBuilder(cfg)
    assert b.num_train_examples == len(train_df) and b.num_valid_examples == len(valid_df)
    recs = [tfrecord.parse_tfrecord_fn(r) for r in b._records(b.train_filenames)]
    assert [r["ann_id"] for r in recs] == list(train_df["ann_id"]) and [r["image_id"] for r in recs] == list(train_df.index)


def test_shard_naming_reproduces_the_reference_dataset_counts(hgb, tmp_path, capsys):
    """Train.ipynb cell 7: 'Train dataset with 66 tfrecords and 134214 examples. / Valid dataset with 3 tfrecords and 5647
    examples.' -- 134214 and 5647 people (gen_tfrecords.ipynb cell 4) in shards of 2048, counted from the file names."""
    train, valid = tmp_path / "train", tmp_path / "valid"
    train.mkdir()
    valid.mkdir()

    def shards(folder, total, per=2048):
        n = total // per + (1 if total % per else 0)
        for k in range(n):
            count = per if k < n - 1 or total % per == 0 else total % per
            (folder / ("file_" + folder.name + "_%.2i-%i.tfrec" % (k, count))).touch()       # gen_tfrecords.py:107

    shards(train, 134214)
    shards(valid, 5647)
    cfg = types.SimpleNamespace(**{k: getattr(hgb.default_config, k) for k in dir(hgb.default_config) if k.isupper()})
    cfg.TRAIN_TFRECORDS_DIR, cfg.VALID_TFRECORDS_DIR = str(train), str(valid)
    b = hgb.dataset_builder.DatasetBuilder(cfg)
    assert capsys.readouterr().out == ("Train dataset with 66 tfrecords and 134214 examples.\n"
                                       "Valid dataset with 3 tfrecords and 5647 examples.\n")
    # and the steps per epoch the reference trained with (Train.ipynb cell 20): 8388 and 352 at batch 16
    assert (b.num_train_examples // 16, b.num_valid_examples // 16) == (8388, 352)
