"""Drop-in for the reference's model/hourglass.py: `create_hourglass_model(...)` returns an object
with the Keras-Model methods the reference's trainer / eval / demo call
(compile, fit, predict, evaluate, save_weights, load_weights, count_params, summary, optimizer).

The network itself lives in libhgb200.so (csrc/model.cu); this file only owns torch tensors that
hold device memory and marshals numpy arrays in and out.  Reference: model/hourglass.py:5-32.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from collections import OrderedDict

import numpy as np

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr

_ACTIVATIONS = {"linear": 0, None: 0, "sigmoid": 1}
_GEMM_KERNELS = ("/kernel", "/pointwise_kernel")     # stored OHWI in the flat parameter buffer


class Adam:
    """tf.keras.optimizers.Adam stand-in (trainer.py:31): legacy-OptimizerV2 hyper-parameters and the
    attributes the reference touches (.learning_rate / .lr with .numpy(), .get_config())."""

    class _Var:
        def __init__(self, v):
            self.value = float(v)

        def numpy(self):
            return np.float32(self.value)

        def assign(self, v):
            self.value = float(v)

        def __float__(self):
            return self.value

        def __repr__(self):
            return f"<hgb200 Variable learning_rate={self.value}>"

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, name="Adam"):
        if amsgrad:
            raise ValueError("amsgrad is not supported")
        self._lr = Adam._Var(learning_rate)
        self.beta_1, self.beta_2, self.epsilon, self.name = float(beta_1), float(beta_2), float(epsilon), name
        self.iterations = 0

    @property
    def learning_rate(self):
        return self._lr

    @learning_rate.setter
    def learning_rate(self, v):
        self._lr.assign(v)

    lr = learning_rate

    def get_config(self):
        return {"name": self.name, "learning_rate": self._lr.value, "decay": 0.0, "beta_1": self.beta_1,
                "beta_2": self.beta_2, "epsilon": self.epsilon, "amsgrad": False}


class History:
    def __init__(self):
        self.history = {}
        self.epoch = []


class _PendingLosses:
    """Per-output losses of a step that has been enqueued but not waited for: the device->host copy into a pinned
    buffer is already on the stream, result() blocks on its event only.  Lets a training loop enqueue step i+1 before it
    reads step i's numbers, so the device never waits for the host between steps (what Keras' fit does)."""

    def __init__(self, host, event, allreduce_on_host):
        self._host, self._event, self._ar, self._value = host, event, allreduce_on_host, None

    def result(self):
        if self._value is None:
            self._event.synchronize()
            per = self._host.numpy().copy()
            if self._ar is not None:
                per = self._ar.sum_host(per)
            self._value = [float(per.sum())] + [float(v) for v in per]
        return self._value


class _Plan:
    """One execution plan (fixed batch, training or inference) = one C handle + its arena."""

    def __init__(self, model, batch, training):
        torch = _lib.require_cuda()
        cfg = _lib.ModelConfig(model.num_classes, model.num_stacks, model.num_channels, model.input_shape[0],
                               model.input_shape[1], model._act, int(batch), int(bool(training)), int(model.mobile))
        self.handle = C.c_void_p()
        check(lib.hgb_model_create(C.byref(cfg), torch.cuda.current_device(), C.byref(self.handle)))
        self.batch, self.training = int(batch), bool(training)
        nbytes = int(lib.hgb_model_buffer_bytes(self.handle, _lib.BUF_ARENA))
        self.arena = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        model._ensure_device_state(self.handle, training)
        check(lib.hgb_model_bind(self.handle, _lib.BUF_PARAMS, ptr(model._params), model._params.numel() * 4))
        if training:
            check(lib.hgb_model_bind(self.handle, _lib.BUF_GRADS, ptr(model._grads), model._grads.numel() * 4))
            check(lib.hgb_model_bind(self.handle, _lib.BUF_ADAM_M, ptr(model._adam_m), model._adam_m.numel() * 4))
            check(lib.hgb_model_bind(self.handle, _lib.BUF_ADAM_V, ptr(model._adam_v), model._adam_v.numel() * 4))
        check(lib.hgb_model_bind(self.handle, _lib.BUF_ARENA, ptr(self.arena), nbytes))
        check(lib.hgb_model_sync_weights(self.handle, stream_ptr()))
        self.weights_version = model._weights_version
        self.comm_key = None          # (communicator, sync_bn) attached to the handle

    def close(self):
        if self.handle:
            lib.hgb_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HourglassModel:
    """What `create_hourglass_model` returns (stands in for the keras.Model of hourglass.py:25)."""

    def __init__(self, num_classes, num_stacks, num_channels, input_shape, predict_activation, seed=None, mobile=False):
        self.mobile = bool(mobile)      # bottleneck_block_mobile (model/hourglass.py:9-11): SeparableConv2D bottlenecks
        if predict_activation not in _ACTIVATIONS:
            raise ValueError(f"predict_activation must be 'sigmoid' or 'linear', got {predict_activation!r}")
        self.num_classes, self.num_stacks, self.num_channels = int(num_classes), int(num_stacks), int(num_channels)
        self.input_shape = tuple(int(v) for v in input_shape)
        if len(self.input_shape) != 3 or self.input_shape[2] != 3:
            raise ValueError("input_shape must be (H, W, 3)")
        self.predict_activation = predict_activation
        self._act = _ACTIVATIONS[predict_activation]
        self.output_names = [f"hg{i}_conv_1x1_predict" for i in range(self.num_stacks)]
        self.optimizer = None
        self.loss = None
        self._loss_kind = None
        self.stop_training = False
        self._plans = {}
        self._weights_version = 0
        self._params = self._grads = self._adam_m = self._adam_v = None
        self._comm_stream = None
        # a host-only handle: parameter table and counts (no GPU needed)
        cfg = _lib.ModelConfig(self.num_classes, self.num_stacks, self.num_channels, self.input_shape[0],
                               self.input_shape[1], self._act, 1, 0, int(self.mobile))
        h = C.c_void_p()
        check(lib.hgb_model_create(C.byref(cfg), 0, C.byref(h)))
        try:
            self._table = self._read_table(h)
            self._n_total = int(lib.hgb_model_param_count(h, 0))
            self._n_train = int(lib.hgb_model_param_count(h, 1))
            self._param_floats = int(lib.hgb_model_buffer_bytes(h, _lib.BUF_PARAMS)) // 4
            self._train_floats = int(lib.hgb_model_buffer_bytes(h, _lib.BUF_GRADS)) // 4
            self._convs = self._read_convs(h)
        finally:
            lib.hgb_model_destroy(h)
        self._host = self._initial_weights(seed)

    # ------------------------------------------------------------------ tables
    @staticmethod
    def _read_table(h):
        name, rank, dims, off, tr = C.c_char_p(), C.c_int(), (C.c_int64 * 4)(), C.c_int64(), C.c_int()
        table = OrderedDict()
        for i in range(lib.hgb_model_num_tensors(h)):
            check(lib.hgb_model_tensor_info(h, i, C.byref(name), C.byref(rank), C.byref(dims), C.byref(off), C.byref(tr)))
            table[name.value.decode()] = (tuple(dims[j] for j in range(rank.value)), int(off.value), bool(tr.value))
        return table

    @staticmethod
    def _read_convs(h):
        name, k, cin, cout, hh, ww, fl = C.c_char_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_double()
        out = []
        for i in range(lib.hgb_model_num_convs(h)):
            check(lib.hgb_model_conv_info(h, i, C.byref(name), C.byref(k), C.byref(cin), C.byref(cout), C.byref(hh),
                                          C.byref(ww), C.byref(fl)))
            out.append(dict(name=name.value.decode(), k=k.value, cin=cin.value, cout=cout.value, h=hh.value, w=ww.value,
                            flops_per_image=fl.value))
        return out

    def _initial_weights(self, seed):
        """Keras defaults: glorot_uniform kernels, zero biases, BN gamma=1 beta=0 mean=0 var=1."""
        rng = np.random.default_rng(seed)
        out = OrderedDict()
        for name, (shape, _off, _tr) in self._table.items():
            if name.endswith("kernel"):     # kernel | depthwise_kernel (k,k,cin,1) | pointwise_kernel (1,1,cin,cout)
                k1, k2, cin, cout = shape
                limit = np.sqrt(6.0 / (k1 * k2 * (cin + cout)))
                out[name] = rng.uniform(-limit, limit, size=shape).astype(np.float32)
            elif name.endswith("/gamma") or name.endswith("/moving_variance"):
                out[name] = np.ones(shape, np.float32)
            else:
                out[name] = np.zeros(shape, np.float32)
        return out

    # ------------------------------------------------------------------ weights
    def _pack(self, weights):
        flat = np.zeros(self._param_floats, np.float32)
        for name, (shape, off, _tr) in self._table.items():
            a = np.asarray(weights[name], np.float32)
            if a.shape != tuple(shape):
                raise ValueError(f"{name}: expected shape {tuple(shape)}, got {a.shape}")
            if name.endswith(_GEMM_KERNELS):
                a = a.transpose(3, 0, 1, 2)  # HWIO -> OHWI: the K-major GEMM operand (depthwise kernels stay [k][k][c])
            flat[off:off + a.size] = a.reshape(-1)
        return flat

    def _unpack(self, flat):
        out = OrderedDict()
        for name, (shape, off, _tr) in self._table.items():
            n = int(np.prod(shape))
            a = flat[off:off + n]
            if name.endswith(_GEMM_KERNELS):
                k1, k2, cin, cout = shape
                a = a.reshape(cout, k1, k2, cin).transpose(1, 2, 3, 0)
            out[name] = np.array(a.reshape(shape), np.float32)
        return out

    def _ensure_device_state(self, handle, training):
        torch = _lib.require_cuda()
        if self._params is None:
            self._params = torch.as_tensor(self._pack(self._host), device="cuda")
            self._host = None
        if training and self._grads is None:
            self._grads = torch.zeros(self._train_floats, dtype=torch.float32, device="cuda")
            self._adam_m = torch.zeros_like(self._grads)
            self._adam_v = torch.zeros_like(self._grads)

    def get_weights_dict(self):
        """OrderedDict Keras-name -> numpy (kernels HWIO)."""
        if self._params is None:
            return OrderedDict((k, v.copy()) for k, v in self._host.items())
        return self._unpack(self._params.cpu().numpy())

    def set_weights_dict(self, weights):
        missing = [k for k in self._table if k not in weights]
        if missing:
            raise KeyError(f"missing weights: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        if self._params is None:
            self._host = OrderedDict((k, np.array(weights[k], np.float32)) for k in self._table)
            self._pack(self._host)  # shape check
        else:
            torch = _lib.require_cuda()
            self._params.copy_(torch.as_tensor(self._pack(weights)))
            self._weights_version += 1

    def get_weights(self):
        return list(self.get_weights_dict().values())

    def set_weights(self, arrays):
        self.set_weights_dict(OrderedDict(zip(self._table.keys(), arrays)))

    def count_params(self):
        return self._n_total

    @property
    def trainable_count(self):
        return self._n_train

    def summary(self, print_fn=print):
        print_fn(f'Model: "hourglass" ({self.num_stacks} stacks, {self.num_channels} channels)')
        for name, (shape, _off, tr) in self._table.items():
            print_fn(f"  {name:60s} {str(tuple(shape)):22s} {'' if tr else '(non-trainable)'}")
        print_fn(f"Total params: {self._n_total:,}")
        print_fn(f"Trainable params: {self._n_train:,}")
        print_fn(f"Non-trainable params: {self._n_total - self._n_train:,}")

    def conv_table(self):
        return list(self._convs)

    def _adam_slots(self):
        """(iterations, m, v) with m / v as {weight name: array} in Keras layout, or None before the first training step."""
        if self.optimizer is None or self._adam_m is None:
            return None
        pad = np.zeros(self._param_floats - self._train_floats, np.float32)
        m = self._unpack(np.concatenate([self._adam_m.cpu().numpy(), pad]))
        v = self._unpack(np.concatenate([self._adam_v.cpu().numpy(), pad]))
        train = [k for k, (_s, _o, tr) in self._table.items() if tr]
        return int(self.optimizer.iterations), {k: m[k] for k in train}, {k: v[k] for k in train}

    def save_weights(self, path, save_format="tf"):
        """keras.Model.save_weights (trainer.py:63-64,141).  `save_format="tf"` (default) writes a real TensorFlow checkpoint
        -- `<path>.index` + `<path>.data-00000-of-00001` with Keras' object-graph keys, BN moving statistics, and the Adam
        step / m / v slots (tf_checkpoint.py) -- so files interchange with the reference; "hgb" writes the same file pair
        with an npz/json payload."""
        from ..parallel import current_allreduce
        ar = current_allreduce()
        if ar is not None:           # data parallel: replicas are identical; rank 0 writes, everyone waits for the file
            try:
                if ar.rank == 0:
                    self._save_weights_local(path, save_format)
            finally:
                ar.barrier()
            return
        self._save_weights_local(path, save_format)

    def _save_weights_local(self, path, save_format):
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        if save_format == "tf":
            from .. import tf_checkpoint
            tf_checkpoint.save_keras_weights(self, path, adam=self._adam_slots())
            return
        if save_format != "hgb":
            raise ValueError(f"unknown save_format {save_format!r}")
        blobs = OrderedDict(self.get_weights_dict())
        meta = {"format": "hgb200-ckpt-1", "num_stacks": self.num_stacks, "num_channels": self.num_channels,
                "num_classes": self.num_classes, "iterations": 0, "optimizer": None}
        if self.optimizer is not None and self._adam_m is not None:
            blobs["__adam_m__"] = self._adam_m.cpu().numpy()
            blobs["__adam_v__"] = self._adam_v.cpu().numpy()
            meta["iterations"] = int(self.optimizer.iterations)
            meta["optimizer"] = self.optimizer.get_config()
        # both files are written under temporary names and renamed (data first), so a reader never sees a torn pair
        with open(path + ".data-00000-of-00001.tmp", "wb") as f:
            np.savez(f, **blobs)
        with open(path + ".index.tmp", "w") as f:
            json.dump(meta, f)
        os.replace(path + ".data-00000-of-00001.tmp", path + ".data-00000-of-00001")
        os.replace(path + ".index.tmp", path + ".index")

    def load_weights(self, path):
        """keras.Model.load_weights (trainer.py:85,188,198): TensorFlow checkpoints written by the reference or by
        save_weights (also a SavedModel's `variables/variables`), and the "hgb" payload; the format is detected."""
        with open(path + ".index", "rb") as f:
            is_json = f.read(1) == b"{"
        if not is_json:
            from .. import tf_checkpoint
            weights, adam = tf_checkpoint.load_keras_weights(self, path)
            self.set_weights_dict(weights)
            self._pending_opt = None
            if adam is not None:
                iterations, m, v = adam
                zeros = {k: np.zeros(sh, np.float32) for k, (sh, _o, tr) in self._table.items() if not tr}
                self._pending_opt = (self._pack({**m, **zeros})[:self._train_floats], self._pack({**v, **zeros})[:self._train_floats],
                                     iterations)
            self._restore_optimizer_state()
            return self
        with open(path + ".index") as f:
            meta = json.load(f)
        if (meta["num_stacks"], meta["num_channels"], meta["num_classes"]) != (self.num_stacks, self.num_channels, self.num_classes):
            raise ValueError("checkpoint was written by a different architecture")
        with np.load(path + ".data-00000-of-00001") as z:
            self.set_weights_dict({k: z[k] for k in self._table})
            self._pending_opt = None
            if "__adam_m__" in z.files:
                self._pending_opt = (z["__adam_m__"], z["__adam_v__"], int(meta.get("iterations", 0)))
        self._restore_optimizer_state()
        return self

    def _restore_optimizer_state(self):
        pend = getattr(self, "_pending_opt", None)
        if pend is None or self.optimizer is None or self._adam_m is None:
            return
        torch = _lib.require_cuda()
        self._adam_m.copy_(torch.as_tensor(pend[0]))
        self._adam_v.copy_(torch.as_tensor(pend[1]))
        self.optimizer.iterations = pend[2]
        self._pending_opt = None

    def save(self, path):
        self.save_weights(os.path.join(path, "variables", "variables"))

    # ------------------------------------------------------------------ plans
    def _plan(self, batch, training):
        key = (int(batch), bool(training))
        p = self._plans.get(key)
        if p is None:
            # at most two plans per mode (the full batch + the short tail batch a pass over a dataset ends with): the
            # least recently used one of the same mode goes first, so a tail batch no longer evicts the full-batch plan
            same = [k for k in self._plans if k[1] == key[1]]
            while len(same) >= 2:
                self._plans.pop(same.pop(0)).close()
            p = _Plan(self, batch, training)
            self._plans[key] = p
        else:
            self._plans[key] = self._plans.pop(key)      # most recently used last
        if p.weights_version != self._weights_version:
            check(lib.hgb_model_sync_weights(p.handle, stream_ptr()))
            p.weights_version = self._weights_version
        return p

    def _to_device_images(self, x):
        torch = _lib.require_cuda()
        if isinstance(x, torch.Tensor):
            t = x.to(device="cuda", dtype=torch.float32, non_blocking=True)
        else:
            t = torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to("cuda", non_blocking=True)
        if t.dim() != 4 or tuple(t.shape[1:]) != self.input_shape:
            raise ValueError(f"images must be (B,{self.input_shape[0]},{self.input_shape[1]},3), got {tuple(t.shape)}")
        return t.contiguous()

    @property
    def heatmap_shape(self):
        return (self.input_shape[0] // 4, self.input_shape[1] // 4, self.num_classes)

    # ------------------------------------------------------------------ inference
    def forward_device(self, images, training=False, plan=None):
        """images: CUDA f32 (B,H,W,3) -> list of S CUDA f32 (B,h,w,K) tensors."""
        torch = _lib.require_cuda()
        B = images.shape[0]
        plan = plan or self._plan(B, training)
        outs = [torch.empty((B,) + self.heatmap_shape, dtype=torch.float32, device="cuda") for _ in range(self.num_stacks)]
        arr = (C.c_void_p * self.num_stacks)(*[o.data_ptr() for o in outs])
        check(lib.hgb_model_forward(plan.handle, ptr(images), int(bool(training)), arr, stream_ptr()))
        return outs

    def predict(self, x, batch_size=32, verbose=0, **_):
        """keras.Model.predict: inference-mode BN; returns a LIST of S numpy arrays (eval.py:106-108)."""
        torch = _lib.require_cuda()
        if not isinstance(x, (np.ndarray, torch.Tensor)) and hasattr(x, "numpy"):
            x = x.numpy()
        n = x.shape[0]
        bs = int(min(batch_size, n)) if n > 0 else 1
        outs = [np.empty((n,) + self.heatmap_shape, np.float32) for _ in range(self.num_stacks)]
        for i in range(0, n, bs):
            chunk = x[i:i + bs]
            m = chunk.shape[0]
            dev = self._to_device_images(chunk)
            if m < bs:  # pad the tail batch: inference BN is per-sample, padding rows are discarded
                dev = torch.cat([dev, dev.new_zeros((bs - m,) + self.input_shape)], 0)
            res = self.forward_device(dev, training=False)
            for s in range(self.num_stacks):
                outs[s][i:i + m] = res[s][:m].cpu().numpy()
        return outs


    # ------------------------------------------------------------------ training
    def compile(self, optimizer=None, loss=None, **_):
        """trainer.py:35: one loss fn applied to every output, summed."""
        from .. import loss as loss_mod
        if optimizer is None or isinstance(optimizer, str):
            optimizer = Adam()
        if not all(hasattr(optimizer, a) for a in ("learning_rate", "beta_1", "beta_2", "epsilon")):
            raise TypeError("optimizer must be an hgb200 Adam (or expose learning_rate/beta_1/beta_2/epsilon)")
        self.optimizer = optimizer
        if not hasattr(optimizer, "iterations"):
            optimizer.iterations = 0
        self.loss = loss
        self._loss_kind = loss_mod.kind_of(loss)
        self._restore_optimizer_state()

    def _current_lr(self):
        lr = self.optimizer.learning_rate
        return float(lr.numpy()) if hasattr(lr, "numpy") else float(lr)

    def train_step_device(self, images, targets, global_batch=None, allreduce=None):
        """One optimizer step on device tensors.  Returns a CUDA float64 tensor [S] of per-output losses.
        allreduce(bucket_tensor) -> handle/None is called per finished gradient segment (DP)."""
        torch = _lib.require_cuda()
        if self._loss_kind is None:
            raise RuntimeError("compile(optimizer, loss) with one of the hgb200.loss functions first")
        B = images.shape[0]
        plan = self._plan(B, True)
        if self.optimizer is not None:
            self._restore_optimizer_state()
        gb = B if global_batch is None else int(global_batch)
        h, w, K = self.heatmap_shape
        inv = 1.0 / (gb * K) if self._loss_kind == 2 else 1.0 / (gb * h * w * K)
        losses = torch.zeros(self.num_stacks, dtype=torch.float64, device="cuda")
        st = stream_ptr()
        if allreduce is not None and getattr(allreduce, "comm", None) is not None:
            key = (allreduce.comm.value, allreduce.sync_bn)       # sync-BN all-reduces statistics in the forward pass already
            if plan.comm_key != key:
                check(lib.hgb_model_set_comm(plan.handle, allreduce.comm, int(allreduce.sync_bn)))
                plan.comm_key = key
        elif plan.comm_key is not None:
            check(lib.hgb_model_set_comm(plan.handle, None, 0))
            plan.comm_key = None
        check(lib.hgb_model_forward(plan.handle, ptr(images), 1, None, st))
        check(lib.hgb_model_loss(plan.handle, self._loss_kind, ptr(targets), inv, ptr(losses), st))
        nseg = lib.hgb_model_num_segments(plan.handle)
        world = 1
        if allreduce is None:
            if plan.comm_key is not None:
                check(lib.hgb_model_set_comm(plan.handle, None, 0))
                plan.comm_key = None
            check(lib.hgb_model_backward(plan.handle, 0, nseg, st))
        elif getattr(allreduce, "comm", None) is not None:
            # the library's own communicator (NCCL): a few contiguous buckets, each all-reduced on a side stream as soon as
            # its segments' backward has been issued (the main chain of the next group starts at once)
            world = allreduce.world_size
            key = (allreduce.comm.value, allreduce.sync_bn)
            if plan.comm_key != key:
                check(lib.hgb_model_set_comm(plan.handle, allreduce.comm, int(allreduce.sync_bn)))
                plan.comm_key = key
            comm = allreduce.comm_stream
            for lo, hi in allreduce.bucket_groups(nseg):
                if allreduce.sync_bn:      # statistics all-reduces share the communicator: everything in plan order on one stream
                    check(lib.hgb_model_backward(plan.handle, lo, hi, st))
                    check(lib.hgb_grad_allreduce_bucket(plan.handle, lo, hi, st))
                    continue
                check(lib.hgb_model_backward_nojoin(plan.handle, lo, hi, st))
                check(lib.hgb_model_lanes_join(plan.handle, C.c_void_p(comm.cuda_stream), 0))
                check(lib.hgb_grad_allreduce_bucket(plan.handle, lo, hi, C.c_void_p(comm.cuda_stream)))
            if not allreduce.sync_bn:
                check(lib.hgb_model_lanes_join(plan.handle, st, 1))
                torch.cuda.current_stream().wait_stream(comm)
        else:
            world = allreduce.world_size
            off, cnt = C.c_int64(), C.c_int64()
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream()
            comm = self._comm_stream
            for seg in range(nseg - 1, -1, -1):
                # the segment's backward does not join its lanes into the main stream (the next segment's chain starts
                # at once); the all-reduce is issued from a side stream that is ordered after everything issued so far
                check(lib.hgb_model_backward_nojoin(plan.handle, seg, seg + 1, st))
                check(lib.hgb_model_lanes_join(plan.handle, C.c_void_p(comm.cuda_stream), 0))
                check(lib.hgb_model_segment_grads(plan.handle, seg, C.byref(off), C.byref(cnt)))
                with torch.cuda.stream(comm):
                    allreduce(self._grads[off.value:off.value + cnt.value])
            check(lib.hgb_model_lanes_join(plan.handle, st, 1))
            allreduce.wait()
        opt = self.optimizer
        opt.iterations += 1
        # The loss kernel already divides by `gb`: with gb = the GLOBAL batch every rank's gradient carries 1/B_global and
        # the SUM all-reduce is the global-mean gradient itself (grad_scale 1).  Only a caller that normalised by its
        # LOCAL batch needs the remaining 1/world here.  Either way Adam sees exactly the single-device gradient of the
        # concatenated batch, so m / v / epsilon behave as on one GPU (tests/test_gpu_multi.py).
        grad_scale = gb / float(B * world)
        check(lib.hgb_model_adam_step(plan.handle, self._current_lr(), opt.beta_1, opt.beta_2, opt.epsilon,
                                      int(opt.iterations), grad_scale, st))
        self._weights_version += 1
        plan.weights_version = self._weights_version   # adam_step refreshed this plan's bf16 operands
        return losses

    def _to_device_targets(self, y):
        torch = _lib.require_cuda()
        if isinstance(y, torch.Tensor):
            t = y.to(device="cuda", dtype=torch.float32, non_blocking=True)
        else:
            if hasattr(y, "numpy") and not isinstance(y, np.ndarray):
                y = y.numpy()
            t = torch.as_tensor(np.ascontiguousarray(np.asarray(y, dtype=np.float32))).to("cuda", non_blocking=True)
        return t.contiguous()

    def _read_losses_deferred(self, losses, ar):
        """Enqueue the (all-reduced) read-back of a step's loss tensor; returns a _PendingLosses."""
        torch = _lib.require_cuda()
        if getattr(self, "_loss_slots", None) is None:
            self._loss_slots = [(torch.empty(self.num_stacks, dtype=torch.float64).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._loss_slot = 0
        host, ev = self._loss_slots[self._loss_slot]
        self._loss_slot = (self._loss_slot + 1) % len(self._loss_slots)
        on_host = None
        if ar:       # per-shard losses carry 1/global_batch: the global loss is their sum (on the device when the group is NCCL)
            if ar.dist.get_backend(ar.group) == "nccl":
                ar.dist.all_reduce(losses, op=ar.dist.ReduceOp.SUM, group=ar.group)
            else:
                on_host = ar
        host.copy_(losses, non_blocking=True)
        ev.record(torch.cuda.current_stream())
        return _PendingLosses(host, ev, on_host)

    def train_on_batch_deferred(self, x, y):
        """train_on_batch without the wait: the step is enqueued, the returned handle's result() gives what
        train_on_batch would have returned.  At most three handles may be outstanding (four pinned read-back slots)."""
        from ..parallel import current_allreduce
        ar = current_allreduce()
        x = self._to_device_images(x.numpy() if hasattr(x, "numpy") and not isinstance(x, np.ndarray) and not hasattr(x, "is_cuda") else x)
        y = self._to_device_targets(y)
        gb = x.shape[0] * (ar.world_size if ar else 1)
        return self._read_losses_deferred(self.train_step_device(x, y, global_batch=gb, allreduce=ar), ar)

    def train_on_batch(self, x, y):
        return self.train_on_batch_deferred(x, y).result()

    def train_on_keypoints(self, images, kps_x, kps_y, kps_v):
        """One optimizer step from (images, keypoints): the Gaussian targets are rendered on the device
        (hgb_render_targets) instead of travelling over PCIe -- BASELINE.json config 2."""
        from .. import ops
        from ..parallel import current_allreduce
        ar = current_allreduce()
        x = self._to_device_images(images)
        h, w, _k = self.heatmap_shape
        y = ops.render_targets(kps_x, kps_y, kps_v, h, w)
        gb = x.shape[0] * (ar.world_size if ar else 1)
        per = self.train_step_device(x, y, global_batch=gb, allreduce=ar).cpu().numpy()
        if ar:
            per = ar.sum_host(per)
        return [float(per.sum())] + [float(v) for v in per]

    def train_on_keypoints_stream(self, batches):
        """Pipelined train_on_keypoints over an iterable of host (images, kps_x, kps_y, kps_v) batches: a generator that
        yields, in order, exactly what train_on_keypoints returns for each batch.

        The reference ends its input pipelines with `.prefetch(AUTOTUNE)` (dataset_builder.py:46,54) and Keras' fit keeps
        the device one step ahead of the host; this is the same overlap, stated explicitly: batch i+1 travels host->device
        on a copy stream (two device slots) while step i computes, and step i's losses come back through a pinned buffer
        that is read only after step i+1 has been enqueued, so neither the PCIe copy nor the host's launch work for the
        next step sits between two steps.  Every batch is still copied host->device and every loss device->host; results
        are identical to calling train_on_keypoints per batch (same kernels, same order on the compute stream)."""
        torch = _lib.require_cuda()
        from .. import ops
        from ..parallel import current_allreduce
        ar = current_allreduce()
        h, w, _k = self.heatmap_shape
        compute = torch.cuda.current_stream()
        if getattr(self, "_h2d_stream", None) is None:
            self._h2d_stream = torch.cuda.Stream(priority=-1)
        copy = self._h2d_stream
        slots = [dict(bufs=None, uploaded=torch.cuda.Event(), consumed=None) for _ in range(2)]

        def as_host(a, dtype):
            if isinstance(a, torch.Tensor):
                return a if a.dtype == dtype else a.to(dtype)
            return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype={torch.float32: np.float32, torch.int32: np.int32}[dtype])))

        def upload(batch, s):
            src = [as_host(batch[0], torch.float32), as_host(batch[1], torch.float32), as_host(batch[2], torch.float32),
                   as_host(batch[3], torch.int32)]
            if src[0].dim() != 4 or tuple(src[0].shape[1:]) != self.input_shape:
                raise ValueError(f"images must be (B,{self.input_shape[0]},{self.input_shape[1]},3), got {tuple(src[0].shape)}")
            sl = slots[s]
            if sl["bufs"] is None or any(b.shape != t.shape for b, t in zip(sl["bufs"], src)):
                sl["bufs"] = [torch.empty(t.shape, dtype=t.dtype, device="cuda") for t in src]
                sl["consumed"] = torch.cuda.Event()      # the allocator may hand back blocks the compute stream still reads
                sl["consumed"].record(compute)
            with torch.cuda.stream(copy):
                if sl["consumed"] is not None:
                    copy.wait_event(sl["consumed"])      # the step that last read this slot has been enqueued and must finish
                for b, t in zip(sl["bufs"], src):
                    b.copy_(t, non_blocking=True)
                sl["uploaded"].record(copy)
            sl["src"] = src      # keeps the host tensors alive until the copy has run

        it = iter(batches)
        try:
            upload(next(it), 0)
        except StopIteration:
            return
        i, pending, more = 0, None, True
        while more:
            s = i & 1
            try:
                upload(next(it), s ^ 1)
            except StopIteration:
                more = False
            sl = slots[s]
            compute.wait_event(sl["uploaded"])
            x, kx, ky, kv = sl["bufs"]
            y = ops.render_targets(kx, ky, kv, h, w)
            gb = x.shape[0] * (ar.world_size if ar else 1)
            losses = self.train_step_device(x, y, global_batch=gb, allreduce=ar)
            sl["consumed"] = torch.cuda.Event()
            sl["consumed"].record(compute)
            nxt = self._read_losses_deferred(losses, ar)
            if pending is not None:
                yield pending.result()
            pending = nxt
            i += 1
        yield pending.result()

    def test_on_batch(self, x, y):
        from .. import ops
        from ..parallel import current_allreduce
        if self._loss_kind is None:
            raise RuntimeError("compile(optimizer, loss) with one of the hgb200.loss functions before evaluate / validation")
        x = self._to_device_images(x.numpy() if hasattr(x, "numpy") and not isinstance(x, np.ndarray) and not hasattr(x, "is_cuda") else x)
        y = self._to_device_targets(y)
        outs = self.forward_device(x, training=False)
        per = [float(ops.loss_fwd_bwd(self._loss_kind, y, o, want_grad=False)[0].item()) for o in outs]
        ar = current_allreduce()
        if ar is not None:      # every rank must see the same val_loss (ModelCheckpoint decides on it): mean over equal shards
            per = (ar.sum_host(np.asarray(per)) / ar.world_size).tolist()
        return [sum(per)] + per

    def _metric_names(self, prefix=""):
        if self.num_stacks == 1:
            return [prefix + "loss"]
        return [prefix + "loss"] + [f"{prefix}{n}_loss" for n in self.output_names]

    def evaluate(self, ds, steps=None, verbose=1, **_):
        it = iter(ds)
        tot = None
        n = 0
        while steps is None or n < steps:
            try:
                x, y = next(it)
            except StopIteration:
                break
            v = np.array(self.test_on_batch(x, y))
            tot = v if tot is None else tot + v
            n += 1
        res = (tot / max(n, 1)).tolist() if tot is not None else [0.0] * (1 + self.num_stacks)
        if self.num_stacks == 1:
            res = res[:1]
        if verbose:
            print(" - ".join(f"{k}: {v:.4f}" for k, v in zip(self._metric_names(), res)))
        return res if len(res) > 1 else res[0]

    def fit(self, ds, epochs=1, callbacks=None, steps_per_epoch=None, validation_data=None, validation_steps=None,
            initial_epoch=0, verbose=1, **_):
        """keras.Model.fit over an (infinite) iterable of (images, heatmaps) batches (trainer.py:49-56)."""
        if steps_per_epoch is None:
            raise ValueError("steps_per_epoch is required (the reference datasets repeat forever)")
        if int(steps_per_epoch) <= 0:
            raise ValueError(f"steps_per_epoch must be positive, got {steps_per_epoch} (fewer examples than one batch?)")
        callbacks = list(callbacks or [])
        hist = History()
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            if hasattr(cb, "on_train_begin"):
                cb.on_train_begin()
        it = iter(ds)
        names = self._metric_names()
        self.stop_training = False
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                if hasattr(cb, "on_epoch_begin"):
                    cb.on_epoch_begin(epoch)
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs}")
            tot = np.zeros(1 + self.num_stacks)
            pending = None       # step i's losses are read after step i+1 has been enqueued: no host round trip between steps
            for _step in range(steps_per_epoch):
                x, y = next(it)
                nxt = self.train_on_batch_deferred(x, y)
                if pending is not None:
                    tot += np.array(pending.result())
                pending = nxt
            tot += np.array(pending.result())
            logs = dict(zip(names, (tot / steps_per_epoch).tolist()))
            if validation_data is not None and validation_steps:
                # Keras creates a fresh validation iterator every epoch: the same first `validation_steps` batches are
                # scored each time.  A one-shot generator cannot restart and is consumed where it stands.
                val_it = iter(validation_data)
                vt = np.zeros(1 + self.num_stacks)
                for _step in range(validation_steps):
                    x, y = next(val_it)
                    vt += np.array(self.test_on_batch(x, y))
                logs.update(zip(self._metric_names("val_"), (vt / validation_steps).tolist()))
            if verbose:
                print(f"{steps_per_epoch}/{steps_per_epoch} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            if hasattr(cb, "on_train_end"):
                cb.on_train_end()
        self.history = hist
        return hist


def create_hourglass_model(num_classes, num_stacks, num_channels, input_shape, predict_activation, mobile=False):
    """Same signature and console output as the reference (model/hourglass.py:5-32)."""
    model = HourglassModel(num_classes, num_stacks, num_channels, input_shape, predict_activation, mobile=mobile)
    print(f'''Created Hourglass model:
    1. {num_stacks} stacks.
    2. {model.count_params()} parameters. Call model.get_summary() for more detail.
    ''')
    return model
