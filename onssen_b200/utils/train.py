"""trainer -- same constructor contract and `.run()` loop as /root/reference/onssen/utils/train.py:7-124
(epoch loop, torch.mean(loss), grad-clip 5, early stop after 8 non-improving validations, best-only
`final.mdl` with the reference's dict layout :111-122).

Repairs of reference defects: `resume_from_checkpoint` is a flag + a separate
`_resume` method (the reference overwrites its own method with a bool, :19-22 vs :36,49) and loads the saved
state_dict into `args.model` (the reference treats it as a module, :51-52)."""
import os
import time

import torch

from .basic import AverageMeter
from .optim import clip_grad_norm_


class trainer:
    def __init__(self, args):
        self.resume = "resume_from_checkpoint" in args and str(args["resume_from_checkpoint"]) == "True"
        self.device = args["device"]
        self.cv_device = args["cv_device"] if "cv_device" in args else self.device
        self.model_name = args["model_name"]
        self.train_loader = args["train_loader"]
        self.valid_loader = args["valid_loader"]
        self.loss_fn = args["loss_fn"]
        self.model = args["model"]
        self.optimizer = args["optimizer"]
        self.epoch = 0
        self.min_loss = float("inf")
        self.early_stop_count = 0
        self.num_epoch = args["num_epoch"]
        self.checkpoint_path = args["checkpoint_path"]
        self.verbose = args["verbose"] if "verbose" in args else True
        if self.resume:
            self._resume(self.checkpoint_path)
        if self.verbose:
            print("Loaded the model...")
        os.makedirs(self.checkpoint_path, exist_ok=True)

    def _resume(self, checkpoint_path):
        saved = torch.load(os.path.join(checkpoint_path, "final.mdl"), weights_only=False)
        self.model.load_state_dict(saved["model"])
        self.epoch = saved["epoch"] + 1
        self.min_loss = saved["cv_loss"]
        self.early_stop_count = saved["early_stop_count"]

    def run(self):
        for epoch in range(self.epoch, self.num_epoch):
            self.train(epoch)
            self.validate(epoch)
            if self.early_stop_count == 8:
                print("Model stops improving, stop the training")
                break
        if self.verbose:
            print("Model training is finished.")

    def train(self, epoch):
        losses, times = AverageMeter(), AverageMeter()
        self.model = self.model.train()
        len_d = len(self.train_loader)
        init_time = end = time.time()
        for i, (input, label) in enumerate(self.train_loader):
            output = self.model(input)
            loss = self.loss_fn(output, label)
            loss_avg = torch.mean(loss)
            losses.update(loss_avg.item())
            self.optimizer.zero_grad()
            loss_avg.backward()
            clip_grad_norm_(self.model.parameters(), 5)
            self.optimizer.step()
            times.update(time.time() - end)
            end = time.time()
            if self.verbose:
                print('epoch %d, %d/%d, training loss: %f, time estimated: %.2f/%.2f seconds' %
                      (epoch, i + 1, len_d, losses.avg, end - init_time, times.avg * len_d), end='\r')
        if self.verbose:
            print("\n")
        return losses.avg

    def validate(self, epoch):
        self.model = self.model.eval()
        losses, times = AverageMeter(), AverageMeter()
        len_d = len(self.valid_loader)
        init_time = end = time.time()
        with torch.no_grad():
            for i, (input, label) in enumerate(self.valid_loader):
                output = self.model(input)
                loss_avg = torch.mean(self.loss_fn(output, label))
                losses.update(loss_avg.item())
                times.update(time.time() - end)
                end = time.time()
                if self.verbose:
                    print('epoch %d, %d/%d, validation loss: %f, time estimated: %.2f/%.2f seconds' %
                          (epoch, i + 1, len_d, losses.avg, end - init_time, times.avg * len_d), end='\r')
        if self.verbose:
            print("\n")
        if losses.avg < self.min_loss:
            self.early_stop_count = 0
            self.min_loss = losses.avg
            torch.save({'model': self.model.state_dict(), 'epoch': epoch, 'optimizer': self.optimizer,
                        'cv_loss': self.min_loss, 'early_stop_count': self.early_stop_count},
                       os.path.join(self.checkpoint_path, "final.mdl"))
            if self.verbose:
                print("Saved new model")
        else:
            self.early_stop_count += 1
        return losses.avg
