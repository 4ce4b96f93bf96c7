"""The reference entry point egs/wsj0-2mix/chimera/msa/run.py:11-29 with the import line swapped to onssen_b200."""
import json
import os
import sys

sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", ".."))

import torch

from onssen_b200 import data, loss, nn, utils
from onssen_b200.utils import AttrDict


def main():
    config_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'config.json')
    with open(config_path) as f:
        args = AttrDict(json.load(f))
    device = torch.device(args.device)
    args.model = nn.chimera(**(args['model_options']))
    args.model.to(device)
    args.train_loader = data.wsj0_2mix_dataloader(args.model_name, args.feature_options, 'tr', device)
    args.valid_loader = data.wsj0_2mix_dataloader(args.model_name, args.feature_options, 'cv', device)
    args.test_loader = data.wsj0_2mix_dataloader(args.model_name, args.feature_options, 'tt', device)
    args.optimizer = utils.build_optimizer(args.model.parameters(), args.optimizer_options)
    args.loss_fn = loss.loss_chimera_msa
    utils.trainer(args).run()
    print("SI-SDR: %.2f" % utils.tester_chimera(args).eval())


if __name__ == "__main__":
    main()
