"""trainer -- same constructor contract and `.run()` loop as /root/reference/onssen/utils/train.py:7-124
(epoch loop, torch.mean(loss), grad-clip 5, early stop after 8 non-improving validations, best-only
`final.mdl` with the reference's dict layout :111-122).

Differences that only affect speed: the running loss is accumulated ON THE DEVICE and read back
every `log_interval` steps (default 20) and at the end of the epoch instead of `loss_avg.item()` every step
(train.py:80), so the host never waits for the device inside an epoch and kernel launches of step i+1 overlap the
tail of step i; under torch.distributed the model's gradients are averaged by `utils.ddp.GradSync` buckets during
the backward, only rank 0 prints and writes checkpoints, and the validation loss is averaged over ranks.

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
        self.log_interval = args["log_interval"] if "log_interval" in args else 20
        import torch.distributed as dist
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if self.world > 1:
            from .ddp import GradSync, broadcast_parameters, enable_global_loss_mean
            broadcast_parameters(self.model)
            enable_global_loss_mean(True)
            if getattr(self.model, "grad_sync", None) is None:
                self.model.grad_sync = GradSync()
            self.verbose = self.verbose and self.rank == 0
        if self.resume:
            self._resume(self.checkpoint_path)
        if self.verbose:
            print("Loaded the model...")
        os.makedirs(self.checkpoint_path, exist_ok=True)

    def _resume(self, checkpoint_path):
        saved = torch.load(os.path.join(checkpoint_path, "final.mdl"), weights_only=False)
        self.model.load_state_dict(saved["model"])
        # Adam moments and step count continue where they stopped: the reference pickles the optimizer object (:120) and
        # never reads it back; the checkpoint keeps exactly its keys, the state is taken from that object
        if saved.get("optimizer") is not None and hasattr(saved["optimizer"], "state_dict"):
            self.optimizer.load_state_dict(saved["optimizer"].state_dict())
        self.epoch = saved["epoch"] + 1
        self.min_loss = saved["cv_loss"]
        self.early_stop_count = saved["early_stop_count"]

    def run(self):
        for epoch in range(self.epoch, self.num_epoch):
            self.train(epoch)
            self.validate(epoch)
            if self.early_stop_count == 8:
                if self.verbose:
                    print("Model stops improving, stop the training")
                break
        if self.verbose:
            print("Model training is finished.")

    def train(self, epoch):
        times = AverageMeter()
        self.model = self.model.train()
        len_d = len(self.train_loader)
        init_time = end = time.time()
        loss_sum, n_steps, avg = None, 0, float("nan")
        for i, (input, label) in enumerate(self.train_loader):
            output = self.model(input)
            loss = self.loss_fn(output, label)
            loss_avg = torch.mean(loss)
            loss_sum = loss_avg.detach().clone() if loss_sum is None else loss_sum + loss_avg.detach()
            n_steps += 1
            self.optimizer.zero_grad()
            loss_avg.backward()
            clip_grad_norm_(self.model.parameters(), 5)
            self.optimizer.step()
            times.update(time.time() - end)
            end = time.time()
            if self.verbose and ((i + 1) % self.log_interval == 0 or i + 1 == len_d):
                avg = float(loss_sum.item()) / n_steps       # the only device->host read of the loop
                self._check_bptt_saturation()
                print('epoch %d, %d/%d, training loss: %f, time estimated: %.2f/%.2f seconds' %
                      (epoch, i + 1, len_d, avg, end - init_time, times.avg * len_d), end='\r')
        if self.verbose:
            print("\n")
        return float(loss_sum.item()) / n_steps if n_steps else avg

    def _check_bptt_saturation(self):
        """surface clamped recurrent gradients of the persistent BPTT kernel (read with the logged loss)"""
        from .. import _lib
        n = _lib.bptt_saturation_count(reset=True)
        if n:
            print(f"\nWARNING: {n} recurrent-gradient values exceeded the fp16 exchange range in the last "
                  f"{self.log_interval} steps (exploding gradient); those steps' gradients were clamped")
        return n

    def validate(self, epoch):
        self.model = self.model.eval()
        times = AverageMeter()
        len_d = len(self.valid_loader)
        init_time = end = time.time()
        loss_sum, n_steps = None, 0
        with torch.no_grad():
            for i, (input, label) in enumerate(self.valid_loader):
                output = self.model(input)
                loss_avg = torch.mean(self.loss_fn(output, label))
                loss_sum = loss_avg.clone() if loss_sum is None else loss_sum + loss_avg
                n_steps += 1
                times.update(time.time() - end)
                end = time.time()
                if self.verbose and ((i + 1) % self.log_interval == 0 or i + 1 == len_d):
                    print('epoch %d, %d/%d, validation loss: %f, time estimated: %.2f/%.2f seconds' %
                          (epoch, i + 1, len_d, float(loss_sum.item()) / n_steps, end - init_time, times.avg * len_d),
                          end='\r')
        if self.verbose:
            print("\n")
        cv_loss = float(loss_sum.item()) / n_steps if n_steps else float("inf")
        if self.world > 1:
            from .dist import sum_over_ranks
            cv_loss = sum_over_ranks(cv_loss, loss_sum.device if loss_sum is not None else None) / self.world
        if cv_loss < self.min_loss:
            self.early_stop_count = 0
            self.min_loss = cv_loss
            if self.rank == 0:
                torch.save({'model': self.model.state_dict(), 'epoch': epoch, 'optimizer': self.optimizer,
                            'cv_loss': self.min_loss, 'early_stop_count': self.early_stop_count},
                           os.path.join(self.checkpoint_path, "final.mdl"))
            if self.verbose:
                print("Saved new model")
        else:
            self.early_stop_count += 1
        return cv_loss
