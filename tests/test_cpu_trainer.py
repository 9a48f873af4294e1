"""Trainer protocol (trainer.py:19-245) on the CPU with a stand-in model: console output, checkpoint / log file names,
the calls made on the model, the best-checkpoint promotion after a resumed session, the loss string table."""
import os
import re
import types
from datetime import date

import pytest


class _FakeModel:
    """Records what Trainer asks of a Keras model; `fit` replays a scripted val_loss curve and drives the callbacks."""

    def __init__(self, val_losses):
        self.calls, self.val_losses, self.optimizer = [], list(val_losses), None

    def compile(self, optimizer=None, loss=None):
        self.optimizer = optimizer
        self.calls.append(("compile", loss.__name__ if loss else None))

    def fit(self, ds, epochs=1, callbacks=(), steps_per_epoch=None, validation_data=None, validation_steps=None, initial_epoch=0):
        self.calls.append(("fit", epochs, steps_per_epoch, validation_steps, initial_epoch))
        hist = {"loss": [], "val_loss": []}
        for cb in callbacks:
            cb.set_model(self)
        for e in range(initial_epoch, epochs):
            for cb in callbacks:
                if hasattr(cb, "on_epoch_begin"):
                    cb.on_epoch_begin(e)
            logs = {"loss": 1.0 / (e + 1), "val_loss": self.val_losses.pop(0)}
            for k, v in logs.items():
                hist[k].append(v)
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(e, logs)
        return types.SimpleNamespace(history=hist)

    def save_weights(self, path):
        self.calls.append(("save_weights", os.path.basename(path)))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        for suffix in (".index", ".data-00000-of-00001"):
            open(path + suffix, "w").write(f"weights after {len(self.calls)} calls")

    def load_weights(self, path):
        self.calls.append(("load_weights", os.path.basename(path)))
        assert os.path.exists(path + ".index")


def _builder():
    return types.SimpleNamespace(build_datasets=lambda: ("train-stream", "valid-stream"), num_train_examples=103, num_valid_examples=37)


def _config(tmp_path):
    return types.SimpleNamespace(BATCH_SIZE=16, CHECKPOINTS_PATH=str(tmp_path / "checkpoints"), LOGS_PATH=str(tmp_path / "logs"))


def test_first_training_then_resume_protocol(tmp_path, capsys):
    import hgb200
    cfg, today = _config(tmp_path), date.today().strftime("%d-%m-%Y")
    model = _FakeModel([0.9, 0.7, 0.8])
    tr = hgb200.Trainer(model, _builder(), epochs=3, learning_rate=0.01, loss_str="Weighted_MSE", config=cfg)
    assert capsys.readouterr().out == "Weighed Mean Squared Error\n"
    assert (tr.steps_per_epoch, tr.valid_steps, tr.batch_size) == (6, 2, 16)            # floor division, trainer.py:23-24
    tr.train()
    out = capsys.readouterr().out
    assert out.startswith(f"First training with:\n    1. Current date {today}.\n    2. Number of epochs 3.\n    3. Batch size 16.\n"
                          f"    4. Optimizer configs: {model.optimizer.get_config()}\n    \n")
    assert "\nLearning rate for epoch 1 is 0.009999999776482582\n" in out and "Learning rate for epoch 3 is 0.009999999776482582" in out   # float32, as Keras prints it
    assert "Epoch 2: val_loss improved from 0.90000 to 0.70000, saving model to " + cfg.CHECKPOINTS_PATH + "/best_val_loss_weights.ckpt" in out
    assert "Epoch 3: val_loss did not improve from 0.70000" in out
    assert re.search(r"---------------------------------------------------------\nFinished training!!\n    - Total training time 0:00:0\d\.\d+\n"
                     rf"    - Temporary checkpoints are saved at {re.escape(cfg.CHECKPOINTS_PATH)}\n    - Log is save at {re.escape(cfg.LOGS_PATH)}\n    \n$", out)
    assert model.calls == [("compile", "weighted_mse"), ("fit", 3, 6, 2, 0), ("save_weights", "best_val_loss_weights.ckpt"),
                           ("save_weights", "best_val_loss_weights.ckpt"), ("save_weights", f"E3_{today}_cont.ckpt")]
    assert sorted(os.listdir(cfg.LOGS_PATH)) == ["log_E3_lr0.01.csv"]
    assert open(os.path.join(cfg.LOGS_PATH, "log_E3_lr0.01.csv")).read().splitlines()[0] == ",loss,val_loss"
    best_before = open(cfg.CHECKPOINTS_PATH + "/best_val_loss_weights.ckpt.index").read()

    # ---- resumed session that improves: temp.ckpt replaces best_val_loss_weights.ckpt
    model2 = _FakeModel([0.75, 0.6])
    tr2 = hgb200.Trainer(model2, _builder(), epochs=2, learning_rate=0.001, loss_str="iou", config=cfg)
    capsys.readouterr()
    tr2.resume_training()
    out = capsys.readouterr().out
    assert out.startswith(f"Loading weights from epoch 3\nLoaded: E3_{today}_cont.ckpt.index\n")
    assert "- Result from last train session number 1 at epoch 3:\nloss: 0.3333333333333333\nval_loss: 0.8\n" in out
    assert "- Best current val_loss at epoch 2:\nloss: 0.5\nval_loss: 0.7\n" in out
    assert (f"Resume training with:\n    1. Train session number 2.\n    2. Current date {today}.\n"
            "    3. Resume training for 2 epochs, from epoch 3 to epoch 5.\n    4. Batch size 16.\n    5. Optimizer configs: ") in out
    assert "Learning rate for epoch 4 is 0.0010000000474974513" in out                                   # forced after load (trainer.py:90)
    assert ("\n---------------------------------------------------------\nComparing current best val_loss with previous best val_loss checkpoints\n"
            "Current best val_loss is lower/better than previous best val_loss\nOld best: 0.7\nNew best: 0.6\n"
            "Replaced old val_loss with new val_loss checkpoints\n---------------------------------------------------------\nFinished training!!\n") in out
    assert re.search(rf"    Temporary checkpoints are saved at {re.escape(cfg.CHECKPOINTS_PATH)}/E5_{today}_cont.ckpt\.\n    Log is saved at ", out)
    assert model2.calls[:3] == [("load_weights", f"E3_{today}_cont.ckpt"), ("compile", "IOU"), ("fit", 5, 6, 2, 3)]
    assert model2.calls[-1] == ("save_weights", f"E5_{today}_cont.ckpt")
    files = sorted(os.listdir(cfg.CHECKPOINTS_PATH))
    assert not any(f.startswith("temp.ckpt") for f in files)
    assert open(cfg.CHECKPOINTS_PATH + "/best_val_loss_weights.ckpt.index").read() != best_before
    assert sorted(os.listdir(cfg.LOGS_PATH)) == ["log_E3_lr0.01.csv", "log_E5_lr0.001.csv"]

    # ---- resumed session without improvement: temp.ckpt is dropped, best stays
    best_now = open(cfg.CHECKPOINTS_PATH + "/best_val_loss_weights.ckpt.index").read()
    model3 = _FakeModel([0.65])
    tr3 = hgb200.Trainer(model3, _builder(), epochs=1, learning_rate=0.001, loss_str="mse", config=cfg)
    capsys.readouterr()
    tr3.resume_train()
    out = capsys.readouterr().out
    assert "Loading weights from epoch 5\n" in out and "No improvement\n" in out and "Train session number 3" in out
    assert open(cfg.CHECKPOINTS_PATH + "/best_val_loss_weights.ckpt.index").read() == best_now
    assert not any(f.startswith("temp.ckpt") for f in os.listdir(cfg.CHECKPOINTS_PATH))

    # ---- loaders
    model4 = _FakeModel([])
    tr4 = hgb200.Trainer(model4, _builder(), epochs=1, learning_rate=0.01, loss_str="weighted_keypoint_mse", config=cfg)
    capsys.readouterr()
    assert tr4.get_best_weights_model() is model4 and tr4.get_lattest_weights_model() is model4
    out = capsys.readouterr().out
    assert out == (f"Loading best weights from {cfg.CHECKPOINTS_PATH}\nLoading lattest trained weights from epoch 6\n"
                   f"Loaded: E6_{today}_cont.ckpt.index\n")
    assert model4.calls == [("load_weights", "best_val_loss_weights.ckpt"), ("compile", "weighed_keypoint_mse"),
                            ("load_weights", f"E6_{today}_cont.ckpt"), ("compile", "weighed_keypoint_mse")]


def test_resume_requires_previous_session_and_names_sort_by_epoch(tmp_path, capsys):
    import hgb200
    cfg = _config(tmp_path)
    tr = hgb200.Trainer(_FakeModel([]), _builder(), 1, 0.01, "nonsense", cfg)
    assert capsys.readouterr().out == "None\n" and tr.loss is None                       # trainer.py:243-245
    with pytest.raises(AssertionError):
        tr.resume_training()
    os.makedirs(cfg.CHECKPOINTS_PATH)
    with pytest.raises(AssertionError):
        hgb200.Trainer.get_epochs_from_name(cfg.CHECKPOINTS_PATH)
    for name in ("E9_01-01-2022_cont.ckpt.index", "E10_02-01-2022_cont.ckpt.index", "E2_03-01-2022_cont.ckpt.index", "best_val_loss_weights.ckpt.index"):
        open(os.path.join(cfg.CHECKPOINTS_PATH, name), "w").close()
    assert hgb200.Trainer.get_epochs_from_name(cfg.CHECKPOINTS_PATH) == ("E10_02-01-2022_cont.ckpt", 10, "E10_02-01-2022_cont.ckpt.index")
    for s, printed in (("MSE", "Mean Squared Error"), ("weight_mean_squared_error", "Weighed Mean Squared Error"), ("IoU", "Intersection over union"),
                       ("mean_squared_error", "Mean Squared Error"), ("weighted_keypoint_mse", "Weighted keypoint mean squared error")):
        assert hgb200.Trainer.get_loss_from_string(s) is not None
        assert capsys.readouterr().out == printed + "\n"
