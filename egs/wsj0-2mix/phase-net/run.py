"""python run.py -c config.json -- phase network (chimera + phase BLSTM) on wsj0-2mix with loss_phase.

phase_net and loss_phase are the REPAIRED versions (the reference classes raise as written, SURVEY.md section 0.3;
see onssen_b200/nn/phase_network.py and onssen_b200/loss/loss_phase.py for the repairs)."""
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
    args.model = nn.phase_net(**(args['model_options']))
    args.model.to(device)
    args.train_loader = data.wsj0_2mix_dataloader(args.model_name, args.feature_options, 'tr', device)
    args.valid_loader = data.wsj0_2mix_dataloader(args.model_name, args.feature_options, 'cv', device)
    args.optimizer = utils.build_optimizer(args.model.parameters(), args.optimizer_options)
    args.loss_fn = loss.loss_phase
    utils.trainer(args).run()


if __name__ == "__main__":
    main()
