"""Chimera (MSA mask objective) on wsj0-2mix -- counterpart of the reference's
egs/wsj0-2mix/chimera/msa/run.py:11-29 + ../evaluate.py; config.json next to this file unless `-c` is given."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(HERE, "..", "..", "..", ".."))

from onssen_b200 import data, loss, nn, utils
from onssen_b200.utils.experiment import load_config, run_experiment


def loader(args, partition, device):
    return data.wsj0_2mix_dataloader(args.model_name, args.feature_options, partition, device)


if __name__ == "__main__":
    run_experiment(load_config(HERE), nn.chimera, loader, loss.loss_chimera_msa, ("tr", "cv", "tt"), utils.tester_chimera)
