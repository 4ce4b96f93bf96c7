"""Enhancement BLSTM with restoration layers on DAPS noisy/clean pairs.  Working counterpart of the reference's
egs/daps/run.py:1-31 (same defects as its Edinburgh script; `enhance` also does not take the `output_dim` that
egs/daps/config.json:18 passes -- dropped here).  An epoch is `<partition>_num_batch` batches of consecutive segments."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.append(os.path.join(HERE, "..", ".."))

from onssen_b200 import data, loss, nn
from onssen_b200.utils.experiment import load_config, run_experiment


def loader(args, partition, device):
    n = args.train_num_batch if partition == "train" else args.validate_num_batch
    return data.daps_enhance_dataloader(n, args.feature_options, partition, device)


if __name__ == "__main__":
    run_experiment(load_config(HERE), nn.enhance, loader, loss.loss_mask_msa, ("train", "validation"),
                   drop_model_keys=("output_dim",))
