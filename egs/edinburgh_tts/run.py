"""python run.py -c config.json -- chimera++ on the Edinburgh noisy/clean corpus ("speaker 2" = the noise).

Working counterpart of the reference's egs/edinburgh_tts/run.py:1-31, which cannot run as written (no argparse
import, `onssen.nn.chimera(args.model_options)` without the ** splat, `self.device` at module level, two extra loader
arguments, a `tester` without `get_est_sig`): same config keys, same objects assigned onto `args`, same loss."""
import argparse
import json
import os
import sys

sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

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
    args.model = nn.chimera(**(args['model_options']))
    args.model.to(device)
    args.train_loader = data.edinburgh_tts_dataloader(args.model_name, args.feature_options, 'train', device)
    args.valid_loader = data.edinburgh_tts_dataloader(args.model_name, args.feature_options, 'validation', device)
    args.optimizer = utils.build_optimizer(args.model.parameters(), args.optimizer_options)
    args.loss_fn = loss.loss_chimera_psa
    utils.trainer(args).run()


if __name__ == "__main__":
    main()
