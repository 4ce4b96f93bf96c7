"""WAV decoding for the loaders: header parsing + raw PCM on the host, everything numeric on the device.

The reference decodes one file at a time with `librosa.load(fn, sr=None)` (mono float32, int16 / 32768) followed by
`librosa.core.resample` when the file's rate differs from the configured one
(/root/reference/onssen/data/feature_utils.py:15-19), serially, on the training thread (wsj0_2mix.py:114-116 calls it
3x per item).  Here:

  * a thread pool reads the files of a batch and drops their RAW int16 samples into one pinned staging buffer
    (no per-sample arithmetic on the host; file reads and memcpy release the GIL);
  * a background thread keeps `depth` batches staged ahead of the consumer, so decoding never runs on the thread that
    launches the training step;
  * on the device: one H2D copy of the int16 block, `onssen_pcm16_to_f32` (scale 1/32768 + channel mean) and, only
    if the rate differs, `onssen_resample_poly` (the polyphase filter scipy.signal.resample_poly would apply -- the
    same Kaiser-windowed FIR, designed on the host once per rate pair; librosa's resampy kernel is a different
    low-pass, so resampled audio stays "parity unpinned", as documented for the host path).
"""
import queue
import struct
import threading
from concurrent.futures import ThreadPoolExecutor
from math import gcd

import numpy as np
import torch

from .. import _lib

_FMT_PCM, _FMT_FLOAT, _FMT_EXT = 1, 3, 0xFFFE


def read_pcm(fn):
    """-> (rate, channels, dtype_code, samples) with samples a numpy array [n_frames, channels] of the FILE's sample
    type (int16 / int32 / float32): RIFF chunk walk, no conversion.  dtype_code: 'i2', 'i4' or 'f4'."""
    with open(fn, "rb") as f:
        raw = f.read()
    if raw[:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise ValueError(f"{fn}: not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        cid, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = pos + 8
        if cid == b"fmt ":
            tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", raw[body:body + 16])
            if tag == _FMT_EXT and size >= 26:
                tag = struct.unpack("<H", raw[body + 24:body + 26])[0]
            fmt = (tag, ch, rate, bits)
        elif cid == b"data":
            data = (body, min(size, len(raw) - body))
            break
        pos = body + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError(f"{fn}: missing fmt/data chunk")
    tag, ch, rate, bits = fmt
    code = {(_FMT_PCM, 16): "i2", (_FMT_PCM, 32): "i4", (_FMT_FLOAT, 32): "f4"}.get((tag, bits))
    if code is None:
        raise ValueError(f"{fn}: unsupported WAV encoding (format tag {tag}, {bits} bits)")
    width = bits // 8 * ch
    n = data[1] // width
    arr = np.frombuffer(raw, dtype="<" + code, count=n * ch, offset=data[0]).reshape(n, ch)
    return rate, ch, code, arr


def to_float_mono(arr, code):
    """host conversion of read_pcm's samples (the soundfile/librosa convention): mono float32"""
    if code == "i2":
        x = arr.astype(np.float32) / 32768.0
    elif code == "i4":
        x = arr.astype(np.float32) / 2147483648.0
    else:
        x = arr.astype(np.float32)
    return x.mean(axis=1) if x.shape[1] > 1 else x[:, 0]


def resample_filter(rate_in, rate_out):
    """The FIR scipy.signal.resample_poly(x, up, down) applies (window ('kaiser', 5.0), half length 10 * max(up, down),
    gain up) with its padding bookkeeping: -> (up, down, h float32 incl. the leading zeros, n_pre_remove)."""
    from scipy.signal import firwin
    g = gcd(int(rate_in), int(rate_out))
    up, down = int(rate_out) // g, int(rate_in) // g
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)) * up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    h = np.concatenate([np.zeros(n_pre_pad), h]).astype(np.float32)
    return up, down, h, n_pre_remove


_FILTERS = {}


def _pinned(t):
    """page-locked staging when a CUDA driver is present (the stager itself also runs on a CPU-only host)"""
    return t.pin_memory() if torch.cuda.is_available() else t


def device_waveforms(staged, sampling_rate, device):
    """staged (PcmStager item) -> ([mix, s1, s2] float32 (B, pitch) on `device`, lengths int32 host tensor (B,)):
    H2D of the raw int16 block, PCM -> float mono on the device, polyphase resampling on the device if needed."""
    rate = staged["rate"]
    lengths = staged["lengths"]
    len_dev = lengths.to(device, non_blocking=True)
    if "pcm" in staged:
        pcm = staged["pcm"].to(device, non_blocking=True)             # (nsig, B, pitch * channels) int16
        nsig, B, _ = pcm.shape
        wav = _lib.pcm16_to_f32(pcm.view(nsig * B, -1), staged["channels"], len_dev.repeat(nsig))
    else:                                                             # files that are not 16-bit PCM: host-converted
        wav = staged["wav"].to(device, non_blocking=True)
        nsig, B, _ = wav.shape
        wav = wav.view(nsig * B, -1)
    if rate != sampling_rate:
        key = (rate, sampling_rate, str(device))
        if key not in _FILTERS:
            up, down, h, pre = resample_filter(rate, sampling_rate)
            _FILTERS[key] = (up, down, torch.from_numpy(h).to(device), pre)
        up, down, h_dev, pre = _FILTERS[key]
        wav, n_out = _lib.resample_poly(wav, len_dev.repeat(nsig), int(lengths.max()), up, down, h_dev, pre)
        lengths = (lengths.to(torch.int64) * up + down - 1).div(down, rounding_mode="floor").to(torch.int32)
    wav = wav.view(nsig, B, -1)
    return [wav[k] for k in range(nsig)], lengths


class PcmStager:
    """Iterable over staged batches: `name_batches` is a list of lists of file-name tuples (one tuple per utterance,
    e.g. (mix, s1, s2)); each item is a dict(pcm=pinned int16 (nsig, B, pitch*channels), lengths=int32 (B,) frames
    per utterance, rate, channels, names).  Decoding runs in `workers` pool threads, `depth` batches ahead.
    Files that are not 16-bit PCM are converted on the host (rare) and re-quantised is NOT done: such a batch is
    delivered as float32 under the key `wav` instead of `pcm`."""

    def __init__(self, name_batches, workers=8, depth=2, transform=None):
        self.name_batches = name_batches
        self.workers, self.depth = workers, depth
        self.transform = transform          # optional callable(list of per-signal arrays) -> list (Edinburgh: noise)

    def _stage(self, pool, names):
        nsig = len(names[0])
        flat = [fn for tup in names for fn in tup]
        dec = list(pool.map(read_pcm, flat))
        rate, ch = dec[0][0], dec[0][1]
        if any(d[0] != rate or d[1] != ch for d in dec):
            raise ValueError("files of one batch differ in sampling rate / channel count")
        B = len(names)
        lengths = np.array([min(dec[b * nsig + k][3].shape[0] for k in range(nsig)) for b in range(B)], dtype=np.int32)
        pitch = int(lengths.max())
        item = dict(rate=rate, channels=ch, names=names, lengths=torch.from_numpy(lengths))
        if all(d[2] == "i2" for d in dec) and self.transform is None:
            # page-locked block straight from torch's caching host allocator (no pageable copy, no full memset: only
            # the tail behind each utterance is zeroed)
            buf = (torch.empty(nsig, B, pitch * ch, dtype=torch.int16, pin_memory=True) if torch.cuda.is_available()
                   else torch.empty(nsig, B, pitch * ch, dtype=torch.int16))
            view = buf.numpy()

            def put(i):
                b, k = divmod(i, nsig)
                n = lengths[b]
                view[k, b, :n * ch] = dec[i][3][:n].reshape(-1)
                view[k, b, n * ch:] = 0

            list(pool.map(put, range(len(dec))))
            item["pcm"] = buf
        else:
            wav = np.zeros((nsig, B, pitch), dtype=np.float32)
            for b in range(B):
                sigs = [to_float_mono(dec[b * nsig + k][3][:lengths[b]], dec[b * nsig + k][2]) for k in range(nsig)]
                if self.transform is not None:
                    sigs = self.transform(sigs)
                for k, x in enumerate(sigs):
                    wav[k, b, :lengths[b]] = x
            item["wav"] = _pinned(torch.from_numpy(wav))
            item["channels"] = 1
        return item

    def __iter__(self):
        q = queue.Queue(maxsize=max(1, self.depth))
        stop = threading.Event()
        # The decode threads run Python between their file reads; with the default 5 ms GIL switch interval one of them
        # can hold the interpreter for milliseconds while the stepping thread waits to enqueue its next kernels.
        import sys
        old_interval = sys.getswitchinterval()
        sys.setswitchinterval(min(old_interval, 2e-4))

        def produce():
            try:
                with ThreadPoolExecutor(max_workers=self.workers) as pool:
                    for names in self.name_batches:
                        if stop.is_set():
                            return
                        q.put(self._stage(pool, names))
                q.put(None)
            except BaseException as exc:          # surface decoding errors on the consumer thread
                q.put(exc)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            sys.setswitchinterval(old_interval)
            stop.set()
            while t.is_alive():               # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    t.join(0.01)
