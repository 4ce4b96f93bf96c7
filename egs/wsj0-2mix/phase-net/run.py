"""Phase network (chimera + phase BLSTM) on wsj0-2mix with loss_phase.  Both are the REPAIRED versions: the reference
classes raise as written (repairs in onssen_b200/nn/phase_network.py and loss/loss_phase.py)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(HERE, "..", "..", ".."))

from onssen_b200 import data, loss, nn
from onssen_b200.utils.experiment import load_config, run_experiment


def loader(args, partition, device):
    return data.wsj0_2mix_dataloader(args.model_name, args.feature_options, partition, device)


if __name__ == "__main__":
    run_experiment(load_config(HERE), nn.phase_net, loader, loss.loss_phase, ("tr", "cv"))
