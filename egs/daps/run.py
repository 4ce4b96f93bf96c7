"""python run.py -c config.json -- enhancement BLSTM with restoration layers on DAPS noisy/clean pairs.

Working counterpart of the reference's egs/daps/run.py:1-31 (same defects as egs/edinburgh_tts/run.py; also
`enhance` does not take the `output_dim` its config passes, enhancement.py:12-18 vs egs/daps/config.json:18 -- the key
is dropped here).  An epoch is `train_num_batch` batches of consecutive frame_length-frame segments."""
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
    options = {k: v for k, v in args['model_options'].items() if k != "output_dim"}
    args.model = nn.enhance(**options)
    args.model.to(device)
    args.train_loader = data.daps_enhance_dataloader(args.train_num_batch, args.feature_options, 'train', device)
    args.valid_loader = data.daps_enhance_dataloader(args.validate_num_batch, args.feature_options, 'validation', device)
    args.optimizer = utils.build_optimizer(args.model.parameters(), args.optimizer_options)
    args.loss_fn = loss.loss_mask_msa
    utils.trainer(args).run()


if __name__ == "__main__":
    main()
