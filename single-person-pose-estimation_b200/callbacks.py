"""Drop-in for the reference's callbacks.py (make_checkpoint_callback, PrintLR) without Keras."""
import numpy as np


class ModelCheckpoint:
    """tf.keras.callbacks.ModelCheckpoint(save_weights_only, monitor='val_loss', mode='min', save_best_only)."""

    def __init__(self, filepath, monitor="val_loss", mode="min", save_best_only=True, verbose=True):
        self.filepath, self.monitor, self.save_best_only, self.verbose = filepath, monitor, save_best_only, verbose
        self.sign = 1.0 if mode == "min" else -1.0
        self.best = np.inf
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if not self.save_best_only or self.sign * cur < self.best:
            if self.verbose:
                print(f"\nEpoch {epoch + 1}: {self.monitor} improved from {self.sign * self.best:.5f} to {cur:.5f}, "
                      f"saving model to {self.filepath}")
            self.best = self.sign * cur
            self.model.save_weights(self.filepath)
        elif self.verbose:
            print(f"\nEpoch {epoch + 1}: {self.monitor} did not improve from {self.sign * self.best:.5f}")


def make_checkpoint_callback(checkpoints_path):
    """callbacks.py:2-8."""
    return ModelCheckpoint(filepath=checkpoints_path, monitor="val_loss", mode="min", save_best_only=True, verbose=True)


class PrintLR:
    """callbacks.py:11-14."""

    def set_model(self, model):
        self.model = model

    def on_epoch_begin(self, epoch, logs=None):
        print("\nLearning rate for epoch {} is {}".format(epoch + 1, self.model.optimizer.lr.numpy()))
