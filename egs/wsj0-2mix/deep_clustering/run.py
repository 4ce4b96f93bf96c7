"""python run.py -c config.json -- the reference entry point (egs/wsj0-2mix/deep_clustering/run.py:13-34) with
the import line swapped to onssen_b200 (and the uninstallable `attrdict` replaced by the in-repo shim)."""
import argparse
import json
import os
import sys

sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))

import torch

from onssen_b200 import data, loss, nn, utils
from onssen_b200.utils import AttrDict


def main():
    parser = argparse.ArgumentParser(description='Parse the config path')
    parser.add_argument("-c", "--config", dest="path", help='The path to the config file. e.g. python run.py --config config.json')
    config = parser.parse_args()
    with open(config.path) as f:
        args = AttrDict(json.load(f))
    device = torch.device(args.device)
    args.model = nn.deep_clustering(**(args['model_options']))
    args.model.to(device)
    args.train_loader = data.wsj0_2mix_dataloader(args.model_name, args.feature_options, 'tr', device)
    args.valid_loader = data.wsj0_2mix_dataloader(args.model_name, args.feature_options, 'cv', device)
    args.test_loader = data.wsj0_2mix_dataloader(args.model_name, args.feature_options, 'tt', device)
    args.optimizer = utils.build_optimizer(args.model.parameters(), args.optimizer_options)
    args.loss_fn = loss.loss_dc
    utils.trainer(args).run()
    print("SI-SDR: %.2f" % utils.tester_dc(args).eval())


if __name__ == "__main__":
    main()
