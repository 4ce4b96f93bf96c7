"""Import the UNMODIFIED reference package (test infrastructure only; see oracle/make_ref.py).

`import_reference()` returns the reference's own `onssen` module: from oracle/_ref/ (the staged copy that travels to
the GPU box) when present, else straight from /root/reference (build container).  `librosa` / `attrdict` are empty
placeholders in both cases.  `patch_stft(onssen)` replaces the one librosa-backed function of the featurizer
(`get_stft`, onssen/data/feature_utils.py:5-21) by a callable that returns the oracle's STFT of an in-memory
waveform, so that the reference's OWN `wsj0_2mix_dataset.get_feature` (crop / tile / log-mag / one-hot / VAD /
cos / phase arithmetic, onssen/data/wsj0_2mix.py:103-158) can run on synthetic signals without wav files.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_STAGED = os.path.join(HERE, "_ref")
REF_LIVE = "/root/reference"


def reference_root():
    if os.path.isdir(os.path.join(REF_STAGED, "onssen")):
        return REF_STAGED
    if os.path.isdir(os.path.join(REF_LIVE, "onssen")):
        return REF_LIVE
    return None


def available():
    return reference_root() is not None


def import_reference():
    """-> the reference `onssen` package (torch modules + losses + data helpers + evaluate.sdr)."""
    if "onssen" in sys.modules and getattr(sys.modules["onssen"], "__onssen_ref_root__", None):
        return sys.modules["onssen"]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not available: run `python oracle/make_ref.py` in the build container")
    stubs = os.path.join(root, "_stubs")
    if os.path.isdir(stubs):
        sys.path.insert(0, stubs)
    else:
        for m in ["librosa", "librosa.core", "librosa.feature", "attrdict"]:
            sys.modules.setdefault(m, types.ModuleType(m))
        sys.modules["librosa"].core = sys.modules["librosa.core"]
        sys.modules["librosa"].feature = sys.modules["librosa.feature"]
        sys.modules["attrdict"].AttrDict = dict
    sys.dont_write_bytecode = True
    sys.path.insert(0, root)
    import onssen
    onssen.__onssen_ref_root__ = root
    return onssen


def patch_stft(onssen, signals):
    """signals: {file name -> float32 waveform}.  After this call the reference's get_stft(fn, sr, n_fft, hop)
    returns oracle.stft(signals[fn], n_fft, hop) -- the only substituted arithmetic (librosa is absent)."""
    from . import onssen_oracle as O

    def get_stft(fn, sampling_rate, window_size, hop_size):
        return O.stft(signals[fn], window_size, hop_size)

    import onssen.data.feature_utils as fu
    import onssen.data.wsj0_2mix as w
    import onssen.data.edinburgh_tts as ed
    import onssen.data.daps_enhance as dp
    for mod in (fu, w, ed, dp):
        mod.get_stft = get_stft
    return get_stft
