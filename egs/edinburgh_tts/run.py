"""Chimera++ on the Edinburgh noisy/clean corpus ("speaker 2" = the noise).  Working counterpart of the reference's
egs/edinburgh_tts/run.py:1-31, which cannot run as written (no argparse import, model options not splatted,
`self.device` at module level, two extra loader arguments, a tester without get_est_sig)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(HERE, "..", ".."))

from onssen_b200 import data, loss, nn
from onssen_b200.utils.experiment import load_config, run_experiment


def loader(args, partition, device):
    return data.edinburgh_tts_dataloader(args.model_name, args.feature_options, partition, device)


if __name__ == "__main__":
    run_experiment(load_config(HERE), nn.chimera, loader, loss.loss_chimera_psa, ("train", "validation"))
