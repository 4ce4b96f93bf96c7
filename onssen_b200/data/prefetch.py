"""Asynchronous device featurizer: batch i+1 is copied (H2D from pinned memory) and featurized on a side stream
while the model runs batch i on the main stream.

The reference featurizes inline on the host inside `for data in train_loader` (onssen/utils/train.py:75, SURVEY.md
section 3.2 "the real wall-clock bottleneck"); here the whole featurizer is two kernel launches, and the
persistent recurrence leaves ~34 of 148 SMs idle.  Measured on B200 (cfg2): useful when the
consumer synchronises every step (it hides the H2D copy), but NOT in a free-running loop: STFT blocks co-reside on the
SMs of the latency-bound persistent recurrent CTAs and take issue slots from them (8.0k -> 6.9k utt/s; 6.0k with the
model on a high-priority stream, which only changes which blocks are placed first) -- so bench.py featurizes inline."""
import torch

from . import feature_utils


class DevicePrefetcher:
    """Wraps an iterable of (wav_mix, wav_s1, wav_s2, crop_start[, lengths]) tuples (host pinned or device
    tensors) and yields (input_list, label_list) on `device`, one batch ahead."""

    def __init__(self, batches, model_name, window_size, hop_size, frame_length, db_threshold, device,
                 label_dtype=torch.float32):
        self.batches = batches
        self.args = (model_name, window_size, hop_size, frame_length, db_threshold)
        self.device = torch.device(device)
        self.label_dtype = label_dtype
        self.side = torch.cuda.Stream(device=self.device)

    def _launch(self, item):
        mix, s1, s2, start = item[:4]
        lengths = item[4] if len(item) > 4 else None
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)          # inputs produced on the main stream (if any) are ready
        with torch.cuda.stream(self.side):
            dev = [t.to(self.device, non_blocking=True) for t in (mix, s1, s2, start)]
            inp, lab = feature_utils.featurize_batch(dev[0], dev[1], dev[2], self.args[0], self.args[1], self.args[2],
                                                     self.args[3], self.args[4], crop_start=dev[3],
                                                     label_dtype=self.label_dtype, lengths=lengths)
            ev = torch.cuda.Event()
            ev.record(self.side)
        for t in inp + lab:
            t.record_stream(main)            # consumed on the main stream: keep the allocator honest
        return inp, lab, ev

    def __iter__(self):
        it = iter(self.batches)
        try:
            nxt = self._launch(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur = nxt
            try:
                nxt = self._launch(next(it))     # enqueued before the consumer's kernels -> overlaps with them
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(cur[2])
            yield cur[0], cur[1]
