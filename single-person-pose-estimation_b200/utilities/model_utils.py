"""Drop-in for the reference's utilities/model_utils.py (checkpoint helpers used by its evaluation notebooks)."""
from __future__ import annotations

import glob


def compile_model_from_checkpoint(model, ckpt_path, optimizer, loss):
    """utilities/model_utils.py:5-22: load the weights at `ckpt_path` (anything before '.index') and compile."""
    model.load_weights(ckpt_path)
    model.compile(optimizer=optimizer, loss=loss)
    return model


def get_epochs_from_ckpt_path(path):
    """utilities/model_utils.py:24-44: every `E{epochs}_{date}_cont.ckpt` under `path` in name order with its epoch number,
    then the best-validation checkpoint with epoch -1."""
    names = sorted(glob.glob(path + "/*_cont.ckpt.index"))
    ckpt_names = [name[:-len(".index")] for name in names]
    epochs = [int(name.split("/")[-1].split("_")[0][1:]) for name in names]
    ckpt_names.append(path + "/best_val_loss_weights.ckpt")
    epochs.append(-1)
    return ckpt_names, epochs
