"""The reference's dataset_builder.py on the GPU: keypoint scaling (:107-111), target-heatmap rendering (:220-238),
and the per-example work of the input pipeline -- resize (:99), flip + affine augmentation of image and keypoints
(:143-185, :270-300) and the colour augmentation (:190-204) -- as batched kernels (`make_train_label_batch`).
The random draws stay on the host (a handful of scalars per example, numpy Generator instead of imgaug / tf.random);
everything that touches pixels runs in libhgb200.  `SyntheticDatasetBuilder` provides the (images, heatmaps)
contract Trainer consumes without TFRecords.
"""
from __future__ import annotations

import numpy as np

from . import _lib, ops


def scale_keypoints(kps, extent, label_extent):
    """dataset_builder.py:107-111: `kps /= float32(extent); kps *= label_extent` (two float32 ops)."""
    kps = np.asarray(kps, dtype=np.float32)
    return ((kps / np.float32(extent)).astype(np.float32) * np.float32(label_extent)).astype(np.float32)


def np_gen_heatmaps(kps_x, kps_y, kps_v, label_shape=(64, 64, 17)):
    """One sample, host arrays in / host array out -- same contract as DatasetBuilder.np_gen_heatmaps."""
    h, w, k = label_shape
    if not (len(kps_x) == len(kps_y) == k):
        raise AssertionError("expected one (x, y) per keypoint")
    out = ops.render_targets(np.asarray(kps_x, np.float32)[None], np.asarray(kps_y, np.float32)[None],
                             np.asarray(kps_v)[None], h, w)
    return out[0].cpu().numpy()


class SyntheticDatasetBuilder:
    """Stands in for DatasetBuilder(config, ratio): `.build_datasets()` -> (ds_train, ds_valid), infinite
    iterables of (images (B,256,256,3) f32 CUDA, heatmaps (B,64,64,17) f32 CUDA); `.num_train_examples`,
    `.num_valid_examples`.  Targets come from hgb_render_targets inside the iterator (config 2 of BASELINE.json)."""

    def __init__(self, config, num_train_examples=64, num_valid_examples=32, seed=0):
        self.image_shape = tuple(config.IMAGE_SHAPE)
        self.label_shape = tuple(config.LABEL_SHAPE)
        self.num_keypoints = int(config.NUM_KEYPOINTS)
        self.batch_size = int(config.BATCH_SIZE)
        self.num_train_examples, self.num_valid_examples = int(num_train_examples), int(num_valid_examples)
        self.seed = seed

    def _stream(self, seed):
        torch = _lib.require_cuda()
        gen = torch.Generator(device="cuda").manual_seed(seed)
        h, w, k = self.label_shape
        while True:
            img = torch.rand((self.batch_size,) + self.image_shape, device="cuda", generator=gen)
            kx = torch.rand((self.batch_size, k), device="cuda", generator=gen) * (w + 8) - 4
            ky = torch.rand((self.batch_size, k), device="cuda", generator=gen) * (h + 8) - 4
            kv = torch.randint(0, 3, (self.batch_size, k), device="cuda", generator=gen, dtype=torch.int32)
            yield img, ops.render_targets(kx, ky, kv, h, w)

    def build_datasets(self):
        return self._stream(self.seed), self._stream(self.seed + 1)

    def np_gen_heatmaps(self, kps_x, kps_y, kps_v):
        return np_gen_heatmaps(kps_x, kps_y, kps_v, self.label_shape)


# ------------------------------------------------------------------ augmentation (dataset_builder.py:143-204)
def flip_partner(num_keypoints, index_flip_pairs):
    """Permutation applied by flip_labels (dataset_builder.py:270-300): slot k receives joint partner[k]."""
    partner = list(range(num_keypoints))
    for a, b in index_flip_pairs:
        partner[a], partner[b] = partner[b], partner[a]
    return np.asarray(partner, np.int32)


_SHIFT_CACHE = {}


def _shift_matrices(height, width, shift_add):
    key = (height, width, shift_add)
    m = _SHIFT_CACHE.get(key)
    if m is None:
        sy, sx = height / 2.0 - shift_add, width / 2.0 - shift_add
        m = _SHIFT_CACHE[key] = (np.array([[1.0, 0.0, -sx], [0.0, 1.0, -sy], [0.0, 0.0, 1.0]]),
                                 np.array([[1.0, 0.0, sx], [0.0, 1.0, sy], [0.0, 0.0, 1.0]]))
    return m


def affine_matrix(height, width, scale, rotate_deg, shift_add):
    """Forward 3x3 float64 matrix of imgaug's Affine(scale, rotate): about (size/2 - shift_add); shift_add is 0.5 for
    pixel arrays and 0 for keypoint coordinates (imgaug `to_matrix` / `to_matrix_cba`).  The two matrix products are
    numpy's, as in imgaug/skimage, so the last bits are theirs."""
    to_topleft, to_center = _shift_matrices(height, width, shift_add)
    rot = np.deg2rad(rotate_deg)
    c, s_ = scale * np.cos(rot), scale * np.sin(rot)
    lin = np.array([[c, -s_, 0.0], [s_, c, 0.0], [0.0, 0.0, 1.0]])
    return to_center @ (lin @ to_topleft)


def _opencv_inverse(forward):
    """The inverse map cv2.warpAffine derives from a forward 2x3 matrix (its order of double operations; plain Python
    floats are IEEE doubles, so this is the same arithmetic as numpy float64 scalars without their overhead)."""
    m0, m1, m2 = float(forward[0][0]), float(forward[0][1]), float(forward[0][2])
    m3, m4, m5 = float(forward[1][0]), float(forward[1][1]), float(forward[1][2])
    det = m0 * m4 - m1 * m3
    det = 1.0 / det if det != 0 else 0.0
    a11, a22 = m4 * det, m0 * det
    m0, m4 = a11, a22
    m1 *= -det
    m3 *= -det
    b1 = -m0 * m2 - m1 * m5
    b2 = -m3 * m2 - m4 * m5
    return np.array([[m0, m1, b1], [m3, m4, b2]])


def augment_1_batch(images, kps_x, kps_y, kps_v, flip, scale, rotate_deg, label_shape=(64, 64, 17),
                    index_flip_pairs=((1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16))):
    """DatasetBuilder.np_augment_1 for a batch with the random draws given: flip (N,) bool, scale (N,), rotate_deg (N,).
    images (N,H,W,3) float32 -> warped images; keypoints in label-map coordinates -> augmented (N,K) float32 x, y."""
    images_shape = tuple(images.shape)
    n, h, w = images_shape[0], images_shape[1], images_shape[2]
    lh, lw, k = label_shape
    inv = np.stack([_opencv_inverse(affine_matrix(h, w, float(s), float(r), 0.5)) for s, r in zip(scale, rotate_deg)]) \
        if n else np.zeros((0, 2, 3))
    fwd = np.stack([affine_matrix(lh, lw, float(s), float(r), 0.0)[:2] for s, r in zip(scale, rotate_deg)]) if n else np.zeros((0, 2, 3))
    aug_images = ops.augment_affine(images, inv, flip)
    ax, ay = ops.augment_keypoints(kps_x, kps_y, kps_v, flip, fwd, flip_partner(k, index_flip_pairs), lw)
    return aug_images, ax, ay


def augment_2_batch(images, brightness_delta, contrast_factor, saturation_factor, hue_delta):
    """DatasetBuilder.augment_2 (:190-204) in place on a CUDA batch, the four tf.image.random_* draws given per example."""
    params = np.stack([np.asarray(brightness_delta, np.float32), np.asarray(contrast_factor, np.float32),
                       np.asarray(saturation_factor, np.float32), np.asarray(hue_delta, np.float32)], axis=1)
    return ops.color_augment(images, params)


def draw_augmentation(rng, n):
    """One set of random draws per example with the reference's ranges: flip p=.5 (:160), scale U(.75,1.25) and rotation
    U(-30,30) degrees (:167), brightness +-0.2, contrast U(.5,2), saturation U(.75,1.25), hue +-0.1 (:194-197)."""
    return {"flip": rng.integers(0, 2, n).astype(bool), "scale": rng.uniform(0.75, 1.25, n), "rotate_deg": rng.uniform(-30.0, 30.0, n),
            "brightness_delta": rng.uniform(-0.2, 0.2, n), "contrast_factor": rng.uniform(0.5, 2.0, n),
            "saturation_factor": rng.uniform(0.75, 1.25, n), "hue_delta": rng.uniform(-0.1, 0.1, n)}


def make_train_label_batch(images, kps_x, kps_y, kps_v, draws, label_shape=(64, 64, 17)):
    """DatasetBuilder.make_train_label (:68-78) for a batch: augmentation 1, augmentation 2, target rendering.  As in the
    reference the renderer receives the ORIGINAL visibility flags (the swapped copy made by flip_labels is dropped at :185)."""
    aug, ax, ay = augment_1_batch(images, kps_x, kps_y, kps_v, draws["flip"], draws["scale"], draws["rotate_deg"], label_shape)
    augment_2_batch(aug, draws["brightness_delta"], draws["contrast_factor"], draws["saturation_factor"], draws["hue_delta"])
    return aug, ops.render_targets(ax, ay, kps_v, label_shape[0], label_shape[1])


def make_valid_label_batch(images, kps_x, kps_y, kps_v, label_shape=(64, 64, 17)):
    """DatasetBuilder.make_valid_label (:81-85)."""
    return images, ops.render_targets(kps_x, kps_y, kps_v, label_shape[0], label_shape[1])


class _Dataset:
    """Restartable stand-in for a tf.data.Dataset: `iter(ds)` builds a new stream; `next(ds)` keeps working on a default one."""

    def __init__(self, factory):
        self._factory = factory
        self._default = None

    def __iter__(self):
        return iter(self._factory())

    def __next__(self):
        if self._default is None:
            self._default = iter(self)
        return next(self._default)

    def close(self):
        if self._default is not None and hasattr(self._default, "close"):
            self._default.close()
        self._default = None


# ------------------------------------------------------------------ prefetch (dataset_builder.py:46,54,65: .prefetch(AUTOTUNE))
class Prefetcher:
    """Runs a batch generator in a background thread, `depth` batches ahead, on its own CUDA stream: record parsing, the
    host half of JPEG decode and the input kernels of batch i+1 overlap the training step of batch i.  Batches are handed
    over with a CUDA event (the consumer's stream waits on it; tensors are marked as used by the consumer's stream), so
    the hand-over never blocks the host.  Iteration order is the generator's; exceptions surface at the consumer."""

    _END = object()

    def __init__(self, generator, depth=2):
        import queue
        import threading
        self._gen = generator
        self._q = queue.Queue(maxsize=max(1, int(depth)))
        self._stop = False
        self._done = False
        self._device = None
        try:
            import torch
            if torch.cuda.is_available():
                self._device = torch.cuda.current_device()      # the worker thread must use the consumer's device
        except Exception:  # pragma: no cover
            pass
        self._thread = threading.Thread(target=self._run, name="hgb200-prefetch", daemon=True)
        self._thread.start()
        import atexit
        import weakref
        ref = weakref.ref(self)
        atexit.register(lambda: ref() is not None and ref().close())

    @staticmethod
    def _tensors(item):
        if hasattr(item, "record_stream"):
            yield item
        elif isinstance(item, dict):
            for v in item.values():
                yield from Prefetcher._tensors(v)
        elif isinstance(item, (tuple, list)):
            for v in item:
                yield from Prefetcher._tensors(v)

    def _put(self, payload):
        import queue
        while not self._stop:
            try:
                self._q.put(payload, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _run(self):
        import contextlib
        cuda = self._device is not None
        if cuda:
            import torch
            torch.cuda.set_device(self._device)
        # highest priority (out-of-range values are clamped): the input kernels are short and must not queue behind the
        # persistent CTAs of the training step, whose lanes run at raised priorities (csrc/model.cu ensure_lanes)
        stream = torch.cuda.Stream(priority=-100) if cuda else None
        try:
            with (torch.cuda.stream(stream) if cuda else contextlib.nullcontext()):
                for item in self._gen:
                    event = None
                    if cuda:
                        event = torch.cuda.Event()
                        event.record(stream)
                    if not self._put((item, event, None)):
                        return
            self._put((Prefetcher._END, None, None))
        except BaseException as ex:  # noqa: BLE001 - handed to the consumer
            self._put((None, None, ex))

    def __iter__(self):
        return self

    def __next__(self):
        if self._done:
            raise StopIteration
        item, event, error = self._q.get()
        if error is not None:
            self._done = True
            raise error
        if item is Prefetcher._END:
            self._done = True
            raise StopIteration
        if event is not None:
            import torch
            current = torch.cuda.current_stream()
            current.wait_event(event)
            for t in Prefetcher._tensors(item):
                if t.is_cuda:
                    t.record_stream(current)
        return item

    def close(self):
        """Stop the worker and wait for it: a daemon thread killed inside the decoder at interpreter exit would take the
        process down with it."""
        self._stop = True
        self._done = True
        try:
            while True:
                self._q.get_nowait()          # unblock a worker waiting for queue space
        except Exception:
            pass
        if self._thread.is_alive() and self._thread is not __import__("threading").current_thread():
            self._thread.join(timeout=10.0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------ the TFRecord-backed builder (dataset_builder.py:10-66)
class DatasetBuilder:
    """Drop-in for the reference's DatasetBuilder(config, ratio): same constructor arguments, attributes, console lines and
    `build_datasets()` / `get_ds_prediction()` contracts, without TensorFlow.  Records are read and parsed on the host
    (tfrecord.py); JPEG decode (nvJPEG), resize, augmentation and target rendering run per BATCH on the GPU, so a batch
    costs a handful of launches instead of one Python callback per example (the reference's tf.numpy_function path).
    Yields CUDA float32 tensors: (images (B,256,256,3), heatmaps (B,64,64,17))."""

    def __init__(self, config, ratio=1, seed=None, shard=None, prefetch=2):
        """`prefetch` > 0 prepares that many batches ahead on a side stream in a background thread -- the reference ends its
        pipelines with `.prefetch(AUTOTUNE)` (dataset_builder.py:46,54,65), so it is ON by default (2 batches); 0 runs the
        input path synchronously on the caller's stream.  Measured on B200 (tools_pipeline_bench.py,
        profiles/r02_pipeline_bench.txt; 8-stack training at batch 128 from JPEG TFRecords): 1682 img/s with the prefetcher,
        1445 synchronous, 1710 from a resident batch.  (Round 1 measured 49.8 img/s: nvJPEG's per-image host synchronisation
        sat on a default-priority stream behind the step's persistent CTAs; the worker's and the decoder's streams now run at
        the highest priority)
        `shard=(rank, world_size)` keeps every world_size-th record starting at `rank` (tf.data's `shard`), so each data-
        parallel process reads a disjoint slice with no exchange; default: the active hgb200.parallel context, else no sharding.
        `num_*_examples` stay GLOBAL counts (steps per epoch = n // BATCH_SIZE with BATCH_SIZE the per-process batch means a
        global batch of BATCH_SIZE * world_size, as with a mirrored strategy)."""
        import glob
        assert 0 < ratio <= 1
        if shard is None:
            from .parallel import current_allreduce
            ar = current_allreduce()
            shard = (ar.rank, ar.world_size) if ar is not None else (0, 1)
        self.shard = (int(shard[0]), int(shard[1]))
        if not 0 <= self.shard[0] < self.shard[1]:
            raise ValueError(f"shard must be (rank, world_size) with 0 <= rank < world_size, got {shard}")
        self.image_shape = tuple(config.IMAGE_SHAPE)
        self.label_shape = tuple(config.LABEL_SHAPE)
        self.num_keypoints = int(config.NUM_KEYPOINTS)
        self.gaussian_kernel = getattr(config, "GAUSSIAN_KERNEL", 7)
        self.sigma = getattr(config, "HM_SIGMA", 1)
        self.index_flip_pairs = [list(p) for p in config.COCO_INDEX_FLIP_PAIRS]
        self.batch_size = int(config.BATCH_SIZE)
        self.shuffle_buffer = int(getattr(config, "SHUFFLE_BUFFER", 1000))
        self.train_filenames = sorted(glob.glob(f"{config.TRAIN_TFRECORDS_DIR}/*.tfrec"))
        self.valid_filenames = sorted(glob.glob(f"{config.VALID_TFRECORDS_DIR}/*.tfrec"))
        if ratio < 1:
            self.train_filenames = self.train_filenames[:int(np.ceil(ratio * len(self.train_filenames)))]
            self.valid_filenames = self.valid_filenames[:int(np.ceil(ratio * len(self.valid_filenames)))]
        self.num_train_examples = self.get_ds_length(self.train_filenames)
        self.num_valid_examples = self.get_ds_length(self.valid_filenames)
        self._rng = np.random.default_rng(seed)
        self.prefetch = int(prefetch)
        print(f"Train dataset with {len(self.train_filenames)} tfrecords and {self.num_train_examples} examples.")
        print(f"Valid dataset with {len(self.valid_filenames)} tfrecords and {self.num_valid_examples} examples.")

    @staticmethod
    def get_ds_length(filenames):
        """dataset_builder.py:303-310: the example count is the number between the last '-' and the extension."""
        length = 0
        for name in filenames:
            length += int(name.split("-")[-1].split(".")[0])
        return length

    @staticmethod
    def parse_tfrecord_fn(example):
        from . import tfrecord
        return tfrecord.parse_tfrecord_fn(example)

    @staticmethod
    def flip_labels(xs, ys, vs, flip_index_pairs):
        """dataset_builder.py:270-300 on plain arrays: swapped copies of x, y and v."""
        partner = flip_partner(len(xs), flip_index_pairs)
        return np.asarray(xs)[partner], np.asarray(ys)[partner], np.asarray(vs)[partner]

    # -- record streams
    def _records(self, filenames, equal_shards=False):
        """Records of this rank's shard (record k belongs to rank k % world).  equal_shards: every rank gets the same
        number of records (the remainder of a pass is dropped), so the training / validation batches of all ranks have the
        same sizes -- the global batch, the BatchNorm statistics and the loss weighting of a data-parallel step assume it."""
        from . import tfrecord
        rank, world = self.shard
        limit = None
        if equal_shards and world > 1:
            limit = self.get_ds_length(filenames) // world * world
        k = 0
        for name in filenames:
            for record in tfrecord.read_records(name):
                if limit is not None and k >= limit:
                    return
                if k % world == rank:
                    yield record
                k += 1

    def _shuffled(self, records):
        """tf.data shuffle(buffer): fill the buffer, then emit a uniformly chosen slot and refill it."""
        buf = []
        for r in records:
            if len(buf) < self.shuffle_buffer:
                buf.append(r)
                continue
            i = int(self._rng.integers(len(buf)))
            out, buf[i] = buf[i], r
            yield out
        while buf:
            yield buf.pop(int(self._rng.integers(len(buf))))

    def _batches(self, records):
        batch = []
        for r in records:
            batch.append(self.parse_tfrecord_fn(r))
            if len(batch) == self.batch_size:
                yield batch
                batch = []
        if batch:
            yield batch                                   # batch() before repeat(): the last batch of a pass may be short

    # -- per-batch device work
    def prepare_examples(self, examples):
        """prepare_example (:88-111) for a batch: decode + resize on device, keypoints to label-map units (two float32 ops)."""
        from . import tfrecord
        from .utilities import data_utils
        images = data_utils.resize_images(tfrecord.decode_jpeg_batch([e["image"] for e in examples]), self.image_shape[0], self.image_shape[1])
        kx = np.stack([scale_keypoints(e["keypoints/x"], e["width"], self.label_shape[1]) for e in examples])
        ky = np.stack([scale_keypoints(e["keypoints/y"], e["height"], self.label_shape[0]) for e in examples])
        kv = np.stack([e["keypoints/vis"] for e in examples]).astype(np.int32)
        return images, kx, ky, kv

    def make_train_label(self, images, kps_x, kps_y, kps_v):
        draws = draw_augmentation(self._rng, int(images.shape[0]))
        aug, ax, ay = augment_1_batch(images, kps_x, kps_y, kps_v, draws["flip"], draws["scale"], draws["rotate_deg"], self.label_shape,
                                      self.index_flip_pairs)
        augment_2_batch(aug, draws["brightness_delta"], draws["contrast_factor"], draws["saturation_factor"], draws["hue_delta"])
        return aug, ops.render_targets(ax, ay, kps_v, self.label_shape[0], self.label_shape[1])

    def make_valid_label(self, images, kps_x, kps_y, kps_v):
        return images, ops.render_targets(kps_x, kps_y, kps_v, self.label_shape[0], self.label_shape[1])

    def np_gen_heatmaps(self, kps_x, kps_y, kps_v):
        return np_gen_heatmaps(kps_x, kps_y, kps_v, self.label_shape)

    def _train_stream(self):
        while True:                                       # .repeat()
            for batch in self._batches(self._shuffled(self._records(self.train_filenames, equal_shards=True))):
                yield self.make_train_label(*self.prepare_examples(batch))

    def _valid_stream(self):
        while True:
            for batch in self._batches(self._records(self.valid_filenames, equal_shards=True)):
                yield self.make_valid_label(*self.prepare_examples(batch))

    def _ahead(self, generator):
        return Prefetcher(generator, self.prefetch) if self.prefetch > 0 else generator

    def build_datasets(self):
        """(ds_train, ds_valid): infinite streams like the reference's (.repeat()); each `iter()` starts a fresh pass from
        the first record -- what Keras does with the validation dataset at the start of every epoch."""
        return (_Dataset(lambda: self._ahead(self._train_stream())), _Dataset(lambda: self._ahead(self._valid_stream())))

    def get_ds_prediction(self):
        return self._ahead(self._prediction_stream())

    def _prediction_stream(self):
        """:58-66 + prepare_prediction_example (:113-137): one pass over the validation records, (images, meta) per batch
        with the meta keys predict_ds reads (eval.py:117-139)."""
        from . import tfrecord
        from .utilities import data_utils
        for batch in self._batches(self._records(self.valid_filenames)):
            images = data_utils.resize_images(tfrecord.decode_jpeg_batch([e["image"] for e in batch]), self.image_shape[0], self.image_shape[1])
            meta = {"ann_id": np.array([e["ann_id"] for e in batch]), "image_id": np.array([e["image_id"] for e in batch]),
                    "coco_url": [e["coco_url"] for e in batch],
                    "keypoints/x": np.stack([e["keypoints/x"] for e in batch]), "keypoints/y": np.stack([e["keypoints/y"] for e in batch]),
                    "keypoints/vis": np.stack([e["keypoints/vis"] for e in batch]),
                    "bbox_x": np.array([e["bbox_x"] for e in batch], np.float32), "bbox_y": np.array([e["bbox_y"] for e in batch], np.float32),
                    "bbox_h": np.array([e["height"] for e in batch]), "bbox_w": np.array([e["width"] for e in batch]),
                    "original_bbox": np.stack([e["original_bbox"] for e in batch])}
            yield images, meta
