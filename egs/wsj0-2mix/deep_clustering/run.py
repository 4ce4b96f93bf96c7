"""Deep clustering on wsj0-2mix: `python run.py -c config.json` (what egs/wsj0-2mix/deep_clustering/run.py:13-34 and
evaluate.py of the reference do, on onssen_b200: training, then SI-SDR of the K-means masks on `tt`)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(HERE, "..", "..", ".."))

from onssen_b200 import data, loss, nn, utils
from onssen_b200.utils.experiment import load_config, run_experiment


def loader(args, partition, device):
    return data.wsj0_2mix_dataloader(args.model_name, args.feature_options, partition, device)


if __name__ == "__main__":
    run_experiment(load_config(HERE), nn.deep_clustering, loader, loss.loss_dc, ("tr", "cv", "tt"), utils.tester_dc)
