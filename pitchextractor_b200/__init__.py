"""pitchextractor_b200 -- B200-native (sm_100a) implementation of the PitchExtractor training hot path."""
__version__ = "0.1.0"

from .mel import LogMel, DEFAULT_MEL_PARAMS  # noqa: E402,F401
from .model import JDCNet, ResBlock, SequenceModel, SinusoidalPositionalEncoding  # noqa: E402,F401
from .meldataset import MelDataset, Collater, build_dataloader, align_length  # noqa: E402,F401
from .optimizers import build_optimizer, FusedAdamW  # noqa: E402,F401
from .trainer import Trainer  # noqa: E402,F401
from .inference import (predict_f0, compute_metrics, rms_cents_error, hz_to_cents,  # noqa: E402,F401
                        estimate_tracking_delay_ms, compute_overshoot_cents)
