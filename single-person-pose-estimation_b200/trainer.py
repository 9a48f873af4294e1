"""Drop-in for the reference's trainer.py: Trainer(model, ds_builder, epochs, learning_rate, loss_str, config)
with train() / resume_training() (alias resume_train()) / get_best_weights_model() /
get_lattest_weights_model(), the same checkpoint and CSV-log file-name protocol and console output.
The step itself (forward, loss, backward, Adam) is the CUDA plan behind model.fit.
"""
from __future__ import annotations

import glob
import math
import os
import time
from datetime import date, timedelta

import pandas as pd

from .callbacks import PrintLR, make_checkpoint_callback
from .loss import IOU, mean_squared_error, weighed_keypoint_mse, weighted_mse
from .model.hourglass import Adam


class Trainer:
    def __init__(self, model, ds_builder, epochs, learning_rate, loss_str, config):
        self.model = model
        self.ds_train, self.ds_valid = ds_builder.build_datasets()
        self.steps_per_epoch = math.ceil(ds_builder.num_train_examples // config.BATCH_SIZE)   # a floor, as in trainer.py:23
        self.valid_steps = math.ceil(ds_builder.num_valid_examples // config.BATCH_SIZE)
        self.epochs = epochs
        self.checkpoints_path = config.CHECKPOINTS_PATH
        self.logs_path = config.LOGS_PATH
        self.learning_rate = learning_rate
        self.batch_size = config.BATCH_SIZE
        self.optimizer = Adam(learning_rate=self.learning_rate)
        self.loss = self.get_loss_from_string(loss_str)

    # ------------------------------------------------------------------ trainer.py:34-71
    def train(self):
        self.model.compile(optimizer=self.optimizer, loss=self.loss)
        today = date.today().strftime("%d-%m-%Y")
        callbacks = [make_checkpoint_callback(self.checkpoints_path + "/best_val_loss_weights.ckpt"), PrintLR()]
        print(f'''First training with:
    1. Current date {today}.
    2. Number of epochs {self.epochs}.
    3. Batch size {self.batch_size}.
    4. Optimizer configs: {self.model.optimizer.get_config()}
    ''')
        start = time.time()
        H = self.model.fit(self.ds_train, epochs=self.epochs, callbacks=callbacks, steps_per_epoch=self.steps_per_epoch,
                           validation_data=self.ds_valid, validation_steps=self.valid_steps)
        end = time.time()
        os.makedirs(self.logs_path, exist_ok=True)
        pd.DataFrame(H.history).to_csv(self.logs_path + f"/log_E{self.epochs}_lr{self.learning_rate}.csv")
        path = self.checkpoints_path + f"/E{self.epochs}_{today}_cont.ckpt"
        self.model.save_weights(path)
        print("---------------------------------------------------------")
        print(f'''Finished training!!
    - Total training time {str(timedelta(seconds=end - start))}
    - Temporary checkpoints are saved at {self.checkpoints_path}
    - Log is save at {self.logs_path}
    ''')
        return H

    # ------------------------------------------------------------------ trainer.py:73-178
    def resume_training(self):
        assert os.path.exists(self.checkpoints_path) and os.path.exists(self.logs_path)
        ckpt_name, previous_epochs, full_name = self.get_epochs_from_name(self.checkpoints_path)
        self.epochs += previous_epochs
        print(f"Loading weights from epoch {previous_epochs}")
        self.model.load_weights(self.checkpoints_path + "/" + ckpt_name)
        print(f"Loaded: {full_name}")
        self.model.compile(optimizer=self.optimizer, loss=self.loss)
        self.model.optimizer.learning_rate = self.learning_rate      # the checkpoint restores Adam state; the LR is forced (trainer.py:90)

        today = date.today().strftime("%d-%m-%Y")
        callbacks = [make_checkpoint_callback(self.checkpoints_path + "/temp.ckpt"), PrintLR()]

        log_filenames = sorted(glob.glob(self.logs_path + "/*"))
        df = pd.concat(map(pd.read_csv, log_filenames), ignore_index=True)
        print("---------------------------------------------------------")
        print(f"- Result from last train session number {len(log_filenames)} at epoch {previous_epochs}:")
        for col, val in df.iloc[-1:].items():
            if col != "Unnamed: 0":
                print(f"{col}: {val.values[0]}")
        print("---------------------------------------------------------")
        min_val_loss = df[df["val_loss"] == df["val_loss"].min()]
        print(f"- Best current val_loss at epoch {min_val_loss.index.values[0] + 1}:")
        for col, val in min_val_loss.items():
            if col != "Unnamed: 0":
                print(f"{col}: {val.values[0]}")
        print("---------------------------------------------------------")
        print(f'''Resume training with:
    1. Train session number {len(log_filenames) + 1}.
    2. Current date {today}.
    3. Resume training for {self.epochs - previous_epochs} epochs, from epoch {previous_epochs} to epoch {self.epochs}.
    4. Batch size {self.batch_size}.
    5. Optimizer configs: {self.model.optimizer.get_config()}
    ''')
        start = time.time()
        H = self.model.fit(self.ds_train, epochs=self.epochs, callbacks=callbacks, steps_per_epoch=self.steps_per_epoch,
                           validation_data=self.ds_valid, validation_steps=self.valid_steps, initial_epoch=previous_epochs)
        end = time.time()
        os.makedirs(self.logs_path, exist_ok=True)
        pd.DataFrame(H.history).to_csv(self.logs_path + f"/log_E{self.epochs}_lr{self.learning_rate}.csv")
        path = self.checkpoints_path + f"/E{self.epochs}_{today}_cont.ckpt"
        self.model.save_weights(path)

        print()
        print("---------------------------------------------------------")
        print("Comparing current best val_loss with previous best val_loss checkpoints")
        prev_min = min_val_loss["val_loss"].values[0]
        curr_min = min(H.history["val_loss"])
        best = [self.checkpoints_path + "/best_val_loss_weights.ckpt" + s for s in (".data-00000-of-00001", ".index")]
        temp = [self.checkpoints_path + "/temp.ckpt" + s for s in (".data-00000-of-00001", ".index")]
        if curr_min < prev_min:
            print("Current best val_loss is lower/better than previous best val_loss")
            print(f"Old best: {prev_min}")
            print(f"New best: {curr_min}")
            if all(os.path.exists(p) for p in best + temp):
                for b, t in zip(best, temp):
                    os.remove(b)
                    os.rename(t, b)
                print("Replaced old val_loss with new val_loss checkpoints")
            else:
                print("Paths do not exist!!")
        else:
            for t in temp:
                if os.path.exists(t):
                    os.remove(t)
            print("No improvement")
        print("---------------------------------------------------------")
        print(f'''Finished training!!
    Total training time {str(timedelta(seconds=end - start))}
    Temporary checkpoints are saved at {path}.
    Log is saved at {self.logs_path}
    ''')
        return H

    resume_train = resume_training      # the README / BASELINE.json spelling

    # ------------------------------------------------------------------ trainer.py:181-201
    def get_best_weights_model(self):
        print(f"Loading best weights from {self.checkpoints_path}")
        self.model.load_weights(self.checkpoints_path + "/best_val_loss_weights.ckpt")
        self.model.compile(optimizer=self.optimizer, loss=self.loss)
        return self.model

    def get_lattest_weights_model(self):
        ckpt_name, previous_epochs, full_name = self.get_epochs_from_name(self.checkpoints_path)
        print(f"Loading lattest trained weights from epoch {previous_epochs}")
        self.model.load_weights(self.checkpoints_path + "/" + ckpt_name)
        print(f"Loaded: {full_name}")
        self.model.compile(optimizer=self.optimizer, loss=self.loss)
        return self.model

    @staticmethod
    def get_epochs_from_name(path):
        """Newest `E{epochs}_{date}_cont.ckpt.index` by epoch number -> (ckpt name, epochs, index file name)."""
        names = glob.glob(path + "/*_cont.ckpt.index")
        assert names
        epoch_of = lambda s: int(os.path.basename(s).split("_")[0][1:])  # noqa: E731
        last = os.path.basename(max(names, key=epoch_of))
        return last[:-len(".index")], epoch_of(last), last

    @staticmethod
    def get_loss_from_string(loss_str):
        """trainer.py:224-245 string table (case-insensitive; unknown -> prints 'None', returns None)."""
        table = {
            "weighted_mse": ("Weighed Mean Squared Error", weighted_mse),
            "weight_mean_squared_error": ("Weighed Mean Squared Error", weighted_mse),
            "mse": ("Mean Squared Error", mean_squared_error),
            "mean_squared_error": ("Mean Squared Error", mean_squared_error),
            "iou": ("Intersection over union", IOU),
            "weighted_keypoint_mse": ("Weighted keypoint mean squared error", weighed_keypoint_mse),
        }
        msg, fn = table.get(loss_str.lower(), ("None", None))
        print(msg)
        return fn
