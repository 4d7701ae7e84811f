"""B200-native (sm_100a) implementation of the acoustic-model hot path of boji123/pytorch-kaldi-asr
(project/attention-transformer-timit + pytorch/utils): same Python model / training-loop / decode API, hand-written
CUDA kernels underneath, reached through the C ABI of csrc/libpka_b200.so (include/pka_b200.h).

Import name: `pytorch_kaldi_asr_b200` (the directory is `pytorch-kaldi-asr_b200/`; the importable alias package at the
repo root extends its __path__ to this directory).
"""
from . import _lib
from .ops import set_compute_mode, compute_mode
from .transformer.Models import Transformer, Encoder, EncoderTest, Decoder, fold_seq_and_mask
from .transformer.Optim import ScheduledOptim, FusedAdam
from .train import train_epoch, cal_loss, get_performance, GraphedTrainStep

__all__ = ["Transformer", "Encoder", "EncoderTest", "Decoder", "fold_seq_and_mask", "ScheduledOptim", "FusedAdam",
           "train_epoch", "cal_loss", "get_performance", "GraphedTrainStep", "set_compute_mode", "compute_mode"]
