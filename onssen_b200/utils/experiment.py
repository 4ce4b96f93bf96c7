"""One experiment = one JSON config + four choices (model class, loader factory, loss, optional tester).

The reference's egs/*/run.py scripts each spell out the same sequence (parse `-c config.json`, wrap it in an AttrDict,
hang model / loaders / optimizer / loss onto it, `trainer(args).run()`, evaluate).  Here the sequence lives once; the
scripts under egs/ only name their four choices.  The `args` object handed to `trainer` / `tester` has exactly the
attributes the reference scripts set."""
import argparse
import json
import os

import torch

from .basic import AttrDict, build_optimizer
from .train import trainer


def load_config(default_dir=None, argv=None):
    """`-c/--config path` like the reference scripts; falls back to config.json next to the calling script."""
    ap = argparse.ArgumentParser(description="onssen_b200 experiment")
    ap.add_argument("-c", "--config", dest="path", default=None, help="path to the JSON config")
    ns = ap.parse_args(argv)
    path = ns.path if ns.path else os.path.join(default_dir or os.getcwd(), "config.json")
    with open(path) as f:
        return AttrDict(json.load(f))


def run_experiment(args, model_cls, make_loader, loss_fn, partitions, tester_cls=None, drop_model_keys=()):
    """partitions: (train, validation[, test]) names understood by `make_loader(args, partition, device)`."""
    device = torch.device(args.device)
    options = {k: v for k, v in args["model_options"].items() if k not in drop_model_keys}
    args.model = model_cls(**options).to(device)
    args.train_loader = make_loader(args, partitions[0], device)
    args.valid_loader = make_loader(args, partitions[1], device)
    if len(partitions) > 2:
        args.test_loader = make_loader(args, partitions[2], device)
    args.optimizer = build_optimizer(args.model.parameters(), args.optimizer_options)
    args.loss_fn = loss_fn
    trainer(args).run()
    if tester_cls is not None and len(partitions) > 2:
        score = tester_cls(args).eval()
        print("SI-SDR: %.2f dB" % score)
        return score
    return None
