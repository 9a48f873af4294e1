"""TFRecords -> nvJPEG -> resize -> augmentation -> targets -> model, and Demo.detect, on a GPU (dataset_builder.py:38-137,
demo.py:25-71, eval.py:99-146 with the reference's dataset object in the loop)."""
import json
import types

import cv2
import numpy as np
import pytest

from oracle import heatmap_oracle as horc
from oracle import input_oracle as iorc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hgb():
    import hgb200
    return hgb200


def _smooth_image(rng, h, w):
    """Photo-like content: band-limited luminance and slowly varying colour.  JPEG decoders differ in IDCT rounding and in
    how they upsample 4:2:0 chroma (libjpeg's triangle filter vs replication), so chroma is kept low-frequency as in
    photographs; the 4:4:4 case in the decode test has no upsampling at all."""
    luma = cv2.resize(rng.random((h // 8 + 2, w // 8 + 2)).astype(np.float32), (w, h), interpolation=cv2.INTER_CUBIC)
    tint = cv2.resize(rng.random((h // 64 + 2, w // 64 + 2, 3)).astype(np.float32), (w, h), interpolation=cv2.INTER_CUBIC)
    return np.clip((0.7 * luma[:, :, None] + 0.3 * tint) * 255, 0, 255).astype(np.uint8)


def _example(rng, ann_id, side):
    img = _smooth_image(rng, side, side)
    ok, enc = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, 95])
    assert ok
    vis = rng.integers(0, 3, 17)
    xs = np.where(vis > 0, rng.random(17) * side, 0).astype(np.float32)
    ys = np.where(vis > 0, rng.random(17) * side, 0).astype(np.float32)
    return {"ann_id": ann_id, "image_id": 7000 + ann_id, "image": enc.tobytes(), "image_path": f"img{ann_id}.jpg", "coco_url": "http://x",
            "width": side, "height": side, "keypoints/x": xs, "keypoints/y": ys, "keypoints/vis": vis, "keypoints/num": int((vis > 0).sum()),
            "bbox_x": np.float32(10.5 + ann_id), "bbox_y": np.float32(20.25), "original_bbox": np.array([15.0, 25.0, side * 0.7, side * 0.8], np.float32)}


@pytest.fixture(scope="module")
def records(hgb, tmp_path_factory):
    from hgb200 import tfrecord
    root = tmp_path_factory.mktemp("tfrecords")
    rng = np.random.default_rng(0)
    train, valid = root / "train", root / "valid"
    train.mkdir()
    valid.mkdir()
    exs_t = [_example(rng, i, int(rng.integers(90, 300))) for i in range(10)]
    exs_v = [_example(rng, 100 + i, int(rng.integers(90, 300))) for i in range(6)]
    tfrecord.write_records(str(train / "file_train_00-6.tfrec"), [tfrecord.build_example(e) for e in exs_t[:6]])
    tfrecord.write_records(str(train / "file_train_01-4.tfrec"), [tfrecord.build_example(e) for e in exs_t[6:]])
    tfrecord.write_records(str(valid / "file_valid_00-6.tfrec"), [tfrecord.build_example(e) for e in exs_v])
    cfg = types.SimpleNamespace(**{k: getattr(hgb.default_config, k) for k in dir(hgb.default_config) if k.isupper()})
    cfg.TRAIN_TFRECORDS_DIR, cfg.VALID_TFRECORDS_DIR, cfg.BATCH_SIZE, cfg.SHUFFLE_BUFFER = str(train), str(valid), 4, 8
    return cfg, exs_t, exs_v


def test_nvjpeg_decode_close_to_libjpeg(hgb):
    from hgb200 import tfrecord
    rng = np.random.default_rng(1)
    streams, want = [], []
    s444 = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    cases = (((120, 200), [cv2.IMWRITE_JPEG_QUALITY, 95] + s444, 0.75, 6), ((64, 64), [cv2.IMWRITE_JPEG_QUALITY, 100] + s444, 0.75, 6),
             ((257, 131), [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_PROGRESSIVE, 1], 1.0, 12),      # 4:2:0, progressive
             ((300, 300), [cv2.IMWRITE_JPEG_QUALITY, 95], 1.0, 12))                                       # 4:2:0, baseline
    tol = []
    for (h, w), flags, mean_tol, max_tol in cases:
        img = _smooth_image(rng, h, w)
        ok, enc = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_RGB2BGR), flags)
        streams.append(enc.tobytes())
        want.append(cv2.cvtColor(cv2.imdecode(enc, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB))
        tol.append((mean_tol, max_tol))
    grey = _smooth_image(rng, 80, 96)[:, :, 0]
    ok, enc = cv2.imencode(".jpg", grey)
    streams.append(enc.tobytes())
    want.append(np.repeat(cv2.imdecode(enc, cv2.IMREAD_GRAYSCALE)[:, :, None], 3, 2))
    tol.append((0.75, 6))
    got = tfrecord.decode_jpeg_batch(streams)
    report = []
    for g, w, (mean_tol, max_tol) in zip(got, want, tol):
        g = g.cpu().numpy()
        assert g.shape == w.shape and g.dtype == np.uint8
        diff = np.abs(g.astype(np.int32) - w.astype(np.int32))
        report.append((round(float(diff.mean()), 3), int(diff.max())))
    for (m, x), (mean_tol, max_tol) in zip(report, tol):
        assert m < mean_tol and x <= max_tol, report                         # IDCT rounding (4:4:4, grey) / chroma upsampling (4:2:0)
    print("nvJPEG vs libjpeg (mean, max abs difference in 8-bit levels):", report)
    assert tfrecord.decode_jpeg_batch([]) == []
    with pytest.raises(ValueError):
        tfrecord.decode_jpeg_batch([b"not a jpeg at all"])


def test_valid_stream_matches_host_restatement(hgb, records):
    cfg, _, exs_v = records
    b = hgb.dataset_builder.DatasetBuilder(cfg, seed=0)
    _, ds_valid = b.build_datasets()
    shapes = []
    for step in range(3):                                            # 4 + 2, then the pass repeats
        images, heat = next(ds_valid)
        shapes.append(images.shape[0])
        batch = (exs_v[:4], exs_v[4:], exs_v[:4])[step]
        assert heat.shape == (len(batch), 64, 64, 17) and images.shape == (len(batch), 256, 256, 3)
        kx = np.stack([horc.scale_keypoints(e["keypoints/x"], e["width"], 64) for e in batch])
        ky = np.stack([horc.scale_keypoints(e["keypoints/y"], e["height"], 64) for e in batch])
        kv = np.stack([e["keypoints/vis"] for e in batch])
        np.testing.assert_array_equal(heat.cpu().numpy(), horc.render_targets(kx, ky, kv, 64, 64))
        for n, e in enumerate(batch):
            dec = cv2.cvtColor(cv2.imdecode(np.frombuffer(e["image"], np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
            want = iorc.crop_resize(dec, None, 256, 256)
            assert np.abs(images[n].cpu().numpy() - want).mean() < 2.0 / 255   # decoder differences only (4:2:0 chroma)
    assert shapes == [4, 2, 4]


def test_train_stream_feeds_the_trainer_contract(hgb, records):
    import torch
    cfg, _, _ = records
    b = hgb.dataset_builder.DatasetBuilder(cfg, seed=3)
    ds_train, _ = b.build_datasets()
    sizes = []
    for _ in range(4):
        images, heat = next(ds_train)
        sizes.append(images.shape[0])
        assert images.is_cuda and images.dtype == torch.float32 and heat.shape[1:] == (64, 64, 17)
        lo, hi = images.amin(dim=(1, 2, 3)), images.amax(dim=(1, 2, 3))
        assert torch.equal(lo, torch.zeros_like(lo)) and torch.equal(hi, torch.ones_like(hi))   # augment_2's final normalisation
        assert float(heat.max()) <= 1.0 and float(heat.min()) == 0.0
    assert sizes == [4, 4, 2, 4]                                     # batch() before repeat(): short batch at the end of a pass
    model = hgb.create_hourglass_model(17, 1, 256, (256, 256, 3), "sigmoid")
    model.compile(optimizer=hgb.Adam(1e-3), loss=hgb.loss.weighted_mse)
    images, heat = next(ds_train)
    l0 = model.train_on_batch(images, heat)
    l1 = model.train_on_batch(images, heat)
    assert np.isfinite(l0).all() and np.isfinite(l1).all()


def test_prediction_stream_through_predict_ds(hgb, records, tmp_path):
    cfg, _, exs_v = records
    b = hgb.dataset_builder.DatasetBuilder(cfg)
    model = hgb.create_hourglass_model(17, 1, 256, (256, 256, 3), "sigmoid")
    out = tmp_path / "result.json"
    preds = hgb.eval.predict_ds(model, b.get_ds_prediction(), b.num_valid_examples, cfg.BATCH_SIZE, hgb.heatmaps_to_keypoints_v2, str(out))
    assert len(preds) == 6 and json.loads(out.read_text()) == preds
    for p, e in zip(preds, exs_v):
        assert p["ann_id"] == e["ann_id"] and p["image_id"] == e["image_id"]
        np.testing.assert_allclose(p["original_bbox"], e["original_bbox"])
        np.testing.assert_allclose(p["xs/gt"], e["keypoints/x"].astype(np.float64) + float(e["bbox_x"]), rtol=1e-6)
        assert p["vs"] == e["keypoints/vis"].tolist() and len(p["confs"]) == 17
    counts = hgb.eval.pck_counts(preds)
    assert counts[1].sum() == sum(int((e["keypoints/vis"] > 0).sum()) for e in exs_v)


def test_demo_detect_matches_the_per_person_reference_flow(hgb):
    rng = np.random.default_rng(5)
    frame = _smooth_image(rng, 360, 480)
    xyxy = np.array([[100.3, 50.2, 220.9, 300.7], [-10.0, 200.0, 90.0, 350.0], [400.0, 10.0, 479.0, 120.0]])
    model = hgb.create_hourglass_model(17, 1, 256, (256, 256, 3), "sigmoid")
    cfg = hgb.default_config
    demo = hgb.demo.Demo(lambda image: xyxy, model, cfg, max_num_ppl=2)
    kps = demo.detect(frame)
    assert len(kps) == 2 and len(demo.square_bboxes) == 2 and len(demo.original_bboxes) == 2       # max_num_ppl applied
    # the same people through the reference's sequence of calls (demo.py:47-63): crop_and_pad, resize, predict, decode, normalise
    crops = []
    for n, box in enumerate(demo.original_bboxes):
        sq = hgb.data_utils.transform_bbox_square(box, cfg.BBOX_SCALE)
        assert sq == demo.square_bboxes[n]
        crops.append(iorc.crop_resize(frame, sq, 256, 256))
        np.testing.assert_array_equal(demo.cropped_images[n].cpu().numpy(), crops[-1])
    heat = model.predict(np.stack(crops))[-1]
    for n in range(2):
        want = horc.decode_batch(heat[n][None], 1e-6, 2)[1][0]
        want[:, 0] /= cfg.LABEL_WIDTH
        want[:, 1] /= cfg.LABEL_HEIGHT
        np.testing.assert_array_equal(kps[n], want)
    pts = demo.keypoints_in_image()
    assert pts[0].shape == (17, 3)
    assert hgb.demo.Demo(lambda image: np.zeros((0, 4)), model, cfg).detect(frame) == []
    with pytest.raises(NotImplementedError):
        demo.show()
