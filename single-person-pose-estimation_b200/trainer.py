"""Drop-in for the reference's trainer.py: Trainer(model, ds_builder, epochs, learning_rate, loss_str, config) with
train() / resume_training() (alias resume_train()) / get_best_weights_model() / get_lattest_weights_model().

What is kept is the *protocol* callers and later sessions depend on -- attribute names, console text, and the files on
disk (`best_val_loss_weights.ckpt`, `temp.ckpt`, `E{epochs}_{dd-mm-YYYY}_cont.ckpt`, `log_E{epochs}_lr{lr}.csv`); the step
itself (forward, loss, backward, Adam) is the CUDA plan behind model.fit.  tests/test_cpu_trainer.py pins the protocol.
"""
from __future__ import annotations

import glob
import os
import time
from datetime import date, timedelta

import pandas as pd

from . import parallel
from .callbacks import PrintLR, make_checkpoint_callback
from .loss import IOU, mean_squared_error, weighed_keypoint_mse, weighted_mse
from .model.hourglass import Adam

_RULE = "-" * 57
_CKPT_PARTS = (".data-00000-of-00001", ".index")

# loss_str -> (what the reference prints, loss callable); trainer.py:224-245
_LOSSES = {}
for _names, _label, _fn in ((("weighted_mse", "weight_mean_squared_error"), "Weighed Mean Squared Error", weighted_mse),
                            (("mse", "mean_squared_error"), "Mean Squared Error", mean_squared_error),
                            (("iou",), "Intersection over union", IOU),
                            (("weighted_keypoint_mse",), "Weighted keypoint mean squared error", weighed_keypoint_mse)):
    for _n in _names:
        _LOSSES[_n] = (_label, _fn)

_FIRST_BANNER = """First training with:
    1. Current date {today}.
    2. Number of epochs {epochs}.
    3. Batch size {batch}.
    4. Optimizer configs: {optimizer}
    """
_RESUME_BANNER = """Resume training with:
    1. Train session number {session}.
    2. Current date {today}.
    3. Resume training for {more} epochs, from epoch {start} to epoch {stop}.
    4. Batch size {batch}.
    5. Optimizer configs: {optimizer}
    """
_FIRST_DONE = """Finished training!!
    - Total training time {elapsed}
    - Temporary checkpoints are saved at {where}
    - Log is save at {logs}
    """
_RESUME_DONE = """Finished training!!
    Total training time {elapsed}
    Temporary checkpoints are saved at {where}.
    Log is saved at {logs}
    """


def _epoch_of(index_file):
    """`.../E{epochs}_{date}_cont.ckpt.index` -> epochs."""
    return int(os.path.basename(index_file).split("_")[0][1:])


def _print_columns(frame):
    for column, values in frame.items():
        if column != "Unnamed: 0":                       # the index column pandas writes into the csv
            print(f"{column}: {values.values[0]}")


class Trainer:
    def __init__(self, model, ds_builder, epochs, learning_rate, loss_str, config):
        self.model = model
        self.ds_train, self.ds_valid = ds_builder.build_datasets()
        self.batch_size = config.BATCH_SIZE
        self.steps_per_epoch = ds_builder.num_train_examples // self.batch_size      # whole batches only (trainer.py:23-24)
        self.valid_steps = ds_builder.num_valid_examples // self.batch_size
        self.epochs = epochs
        self.learning_rate = learning_rate
        self.checkpoints_path, self.logs_path = config.CHECKPOINTS_PATH, config.LOGS_PATH
        self.optimizer = Adam(learning_rate=learning_rate)
        self.loss = self.get_loss_from_string(loss_str)

    # ------------------------------------------------------------------ shared by both kinds of session
    def _ckpt(self, name):
        return f"{self.checkpoints_path}/{name}"

    def _fit(self, best_checkpoint, initial_epoch=None):
        """model.fit with the session's two callbacks -> (History, seconds)."""
        extra = {} if initial_epoch is None else {"initial_epoch": initial_epoch}
        began = time.time()
        history = self.model.fit(self.ds_train, epochs=self.epochs, steps_per_epoch=self.steps_per_epoch,
                                 validation_data=self.ds_valid, validation_steps=self.valid_steps,
                                 callbacks=[make_checkpoint_callback(self._ckpt(best_checkpoint)), PrintLR()], **extra)
        return history, time.time() - began

    def _persist(self, history, today):
        """CSV log of the session and the `_cont` checkpoint the next session resumes from -> its path."""
        if parallel.is_primary():          # data parallel: one writer; save_weights gates and synchronises itself
            os.makedirs(self.logs_path, exist_ok=True)
            pd.DataFrame(history.history).to_csv(f"{self.logs_path}/log_E{self.epochs}_lr{self.learning_rate}.csv")
        path = self._ckpt(f"E{self.epochs}_{today}_cont.ckpt")
        self.model.save_weights(path)
        return path

    # ------------------------------------------------------------------ trainer.py:34-71
    def train(self):
        self.model.compile(optimizer=self.optimizer, loss=self.loss)
        today = date.today().strftime("%d-%m-%Y")
        print(_FIRST_BANNER.format(today=today, epochs=self.epochs, batch=self.batch_size, optimizer=self.model.optimizer.get_config()))
        history, seconds = self._fit("best_val_loss_weights.ckpt")
        self._persist(history, today)
        print(_RULE)
        print(_FIRST_DONE.format(elapsed=timedelta(seconds=seconds), where=self.checkpoints_path, logs=self.logs_path))
        return history

    # ------------------------------------------------------------------ trainer.py:73-178
    def resume_training(self):
        """Continue from the newest `_cont` checkpoint for `epochs` more epochs.  The session's best weights go to
        `temp.ckpt` and replace `best_val_loss_weights.ckpt` only if they beat the best val_loss of all earlier logs."""
        assert os.path.exists(self.checkpoints_path) and os.path.exists(self.logs_path)
        ckpt_name, done_epochs, index_name = self.get_epochs_from_name(self.checkpoints_path)
        self.epochs += done_epochs
        print(f"Loading weights from epoch {done_epochs}")
        self.model.load_weights(self._ckpt(ckpt_name))
        print(f"Loaded: {index_name}")
        self.model.compile(optimizer=self.optimizer, loss=self.loss)
        self.model.optimizer.learning_rate = self.learning_rate          # the checkpoint restores Adam's state; the rate is forced (:90)

        logs = sorted(glob.glob(self.logs_path + "/*"))
        past = pd.concat(map(pd.read_csv, logs), ignore_index=True)
        best_so_far = past[past["val_loss"] == past["val_loss"].min()]
        print(_RULE)
        print(f"- Result from last train session number {len(logs)} at epoch {done_epochs}:")
        _print_columns(past.iloc[-1:])
        print(_RULE)
        print(f"- Best current val_loss at epoch {best_so_far.index.values[0] + 1}:")
        _print_columns(best_so_far)
        print(_RULE)
        today = date.today().strftime("%d-%m-%Y")
        print(_RESUME_BANNER.format(session=len(logs) + 1, today=today, more=self.epochs - done_epochs, start=done_epochs,
                                    stop=self.epochs, batch=self.batch_size, optimizer=self.model.optimizer.get_config()))
        history, seconds = self._fit("temp.ckpt", initial_epoch=done_epochs)
        path = self._persist(history, today)

        print()
        print(_RULE)
        print("Comparing current best val_loss with previous best val_loss checkpoints")
        self._promote_session_best(best_so_far["val_loss"].values[0], min(history.history["val_loss"]))
        print(_RULE)
        print(_RESUME_DONE.format(elapsed=timedelta(seconds=seconds), where=path, logs=self.logs_path))
        return history

    resume_train = resume_training      # the README / BASELINE.json spelling

    def _promote_session_best(self, previous_best, session_best):
        ar = parallel.current_allreduce()
        if ar is not None:                 # every rank took the same decision (val_loss is all-reduced); rank 0 moves the files
            try:
                if ar.rank == 0:
                    self._promote_local(previous_best, session_best)
            finally:
                ar.barrier()
            return
        self._promote_local(previous_best, session_best)

    def _promote_local(self, previous_best, session_best):
        kept = [self._ckpt("best_val_loss_weights.ckpt") + part for part in _CKPT_PARTS]
        fresh = [self._ckpt("temp.ckpt") + part for part in _CKPT_PARTS]
        if not session_best < previous_best:
            for f in fresh:
                if os.path.exists(f):
                    os.remove(f)
            print("No improvement")
            return
        print("Current best val_loss is lower/better than previous best val_loss")
        print(f"Old best: {previous_best}")
        print(f"New best: {session_best}")
        if not all(os.path.exists(f) for f in kept + fresh):
            print("Paths do not exist!!")
            return
        for old, new in zip(kept, fresh):
            os.replace(new, old)
        print("Replaced old val_loss with new val_loss checkpoints")

    # ------------------------------------------------------------------ trainer.py:181-201
    def get_best_weights_model(self):
        print(f"Loading best weights from {self.checkpoints_path}")
        self.model.load_weights(self._ckpt("best_val_loss_weights.ckpt"))
        self.model.compile(optimizer=self.optimizer, loss=self.loss)
        return self.model

    def get_lattest_weights_model(self):
        ckpt_name, done_epochs, index_name = self.get_epochs_from_name(self.checkpoints_path)
        print(f"Loading lattest trained weights from epoch {done_epochs}")
        self.model.load_weights(self._ckpt(ckpt_name))
        print(f"Loaded: {index_name}")
        self.model.compile(optimizer=self.optimizer, loss=self.loss)
        return self.model

    @staticmethod
    def get_epochs_from_name(path):
        """The `_cont` checkpoint with the highest epoch number -> (name to load, epochs, index file name); trainer.py:203-222."""
        candidates = glob.glob(path + "/*_cont.ckpt.index")
        assert candidates
        newest = os.path.basename(max(candidates, key=_epoch_of))
        return newest[:-len(".index")], _epoch_of(newest), newest

    @staticmethod
    def get_loss_from_string(loss_str):
        """Case-insensitive string table of trainer.py:224-245; an unknown name prints 'None' and returns None."""
        label, fn = _LOSSES.get(loss_str.lower(), ("None", None))
        print(label)
        return fn
