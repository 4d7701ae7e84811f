"""Checkpoints and model averaging -- the step right after the training loop.

The reference pickles the whole `nn.Module` (`torch.save({'model': model, 'model_options': opt, 'epoch': n,
'train_options': opt})`, L/initialize_model.py:90-95, L/train.py:252-260) and averages the last N epoch files in
`combine` (L/train.py:276-322).  Here a checkpoint is plain data:

    {'format': 'pka-b200-checkpoint-v1', 'state_dict': {name: cpu tensor}, 'model_options': {...}, 'epoch': n,
     'train_options': {...}, 'optimizer': {...} | None, 'extra': {...}}

`state_dict` has the reference's keys and shapes, so it loads into either implementation.  `optimizer` carries what a
bit-exact resume needs and the reference never saved: Adam moments and step count (in torch.optim.Adam's own
state-dict layout, so it also loads into / from a stock Adam), the LR-schedule step and the dropout counter.

`load_checkpoint` reads both this format and the reference's pickled-module files: with
`pytorch_kaldi_asr_b200.dropin` installed the pickled class paths (`transformer.Models.Transformer`, `TDNN.*`) resolve
to this package, the weights are taken from the unpickled object's `state_dict()` and a fresh model is built from
`model_options` (an argparse Namespace with the flag names of L/initialize_model.py:24-46).
"""
from __future__ import annotations

import argparse
import os
from typing import Dict, Iterable, Iterator, Optional, Tuple

import numpy as np
import torch

FORMAT = "pka-b200-checkpoint-v1"

# constructor keyword <- name in the reference's model_options Namespace (L/initialize_model.py:70-88)
_OPTION_NAMES = dict(n_src_dim="src_dim", n_tgt_vocab="tgt_vocab_dim", encoder_max_len="encoder_max_len",
                     decoder_max_len="decoder_max_len", src_fold="src_fold", decoder_sub_sequence="decoder_sub_sequence",
                     en_layers="en_layers", de_layers="de_layers", n_head="n_head", en_d_model="en_d_model",
                     de_d_model="de_d_model", d_k="d_k", d_v="d_v", en_dropout="en_dropout", de_dropout="de_dropout",
                     tdnn_contexts="tdnn_contexts")
_EXTRA_KWARGS = ("encoder_type", "cmvn", "seed")
_LDA_W, _LDA_B = "encoder_test.lda_layer.weight", "encoder_test.lda_layer.bias"


def options_dict(options) -> dict:
    """argparse Namespace / dict / None -> plain dict (what gets stored)."""
    if options is None:
        return {}
    if isinstance(options, argparse.Namespace):
        return dict(vars(options))
    return dict(options)


def model_kwargs(model_options) -> dict:
    """Transformer constructor keywords from the reference's `model_options` (flag names) or from a dict that already
    uses the constructor's names.  `encoder_sub_sequence` is what L/initialize_model.py:77 hard-wires, (-100, 0),
    unless the options use the constructor's own name for it."""
    opt = options_dict(model_options)
    kw = {}
    for ctor_name, flag_name in _OPTION_NAMES.items():
        if ctor_name in opt:
            kw[ctor_name] = opt[ctor_name]
        elif flag_name in opt:
            kw[ctor_name] = opt[flag_name]
    missing = [k for k in ("n_src_dim", "n_tgt_vocab", "encoder_max_len", "decoder_max_len") if k not in kw]
    if missing:
        raise ValueError("[ERROR] model_options lacks %s" % ", ".join(missing))
    if "n_src_dim" in opt and "encoder_sub_sequence" in opt:          # constructor-style options: take it as given
        kw["encoder_sub_sequence"] = tuple(opt["encoder_sub_sequence"])
    else:                                                             # reference flags: the flag is parsed but ignored
        kw["encoder_sub_sequence"] = (-100, 0)
    if "decoder_sub_sequence" in kw:
        kw["decoder_sub_sequence"] = tuple(kw["decoder_sub_sequence"])
    for name in _EXTRA_KWARGS:
        if name in opt:
            kw[name] = opt[name]
    return kw


def lda_from_state_dict(state_dict) -> Optional[np.ndarray]:
    """Rebuild the Kaldi lda.mat ([out, in + 1], last column = offset) the TDNN encoder was constructed with."""
    if _LDA_W not in state_dict:
        return None
    w, b = state_dict[_LDA_W].detach().cpu().float(), state_dict[_LDA_B].detach().cpu().float()
    return torch.cat([w.t(), b.view(-1, 1)], dim=1).numpy()


def build_model(model_options, state_dict, device=None):
    """Fresh Transformer from options + weights (strict load: same keys and shapes as the reference)."""
    from .transformer.Models import Transformer
    kw = model_kwargs(model_options)
    lda = lda_from_state_dict(state_dict)
    if lda is None and kw.get("encoder_type", "tdnn") == "tdnn":
        raise ValueError("[ERROR] the state dict has no LDA layer; is this an attention-encoder model? set encoder_type")
    model = Transformer(lda_mat=lda, **kw)
    model.load_state_dict(state_dict, strict=True)
    return model.to(device) if device is not None else model


# ------------------------------------------------------------------------------------------------ optimiser state
def optimizer_state(optimizer, model=None) -> dict:
    """Everything a bit-exact resume needs.  `optimizer` is a ScheduledOptim (or a bare FusedAdam / torch Adam)."""
    inner = getattr(optimizer, "optimizer", optimizer)
    state = dict(adam=_cpu(inner.state_dict()))
    if inner is not optimizer:
        n_steps = int(optimizer.n_current_steps)
        fused = state["adam"].get("fused")
        if fused is not None:
            # A CUDA-graph step advances the schedule on the device only (train.GraphedTrainStep): the device counter
            # is the truth, the host-side mirror and the param-group lr are brought in line before they are written
            n_steps = max(n_steps, int(fused["n_current_steps"]))
            if n_steps != int(optimizer.n_current_steps):
                optimizer.n_current_steps = n_steps
                lr = (optimizer.start_lr * optimizer.soft_coefficient) / (n_steps + optimizer.soft_coefficient)
                for group in inner.param_groups:
                    group["lr"] = lr
                state["adam"] = _cpu(inner.state_dict())
        state["schedule"] = dict(n_current_steps=n_steps, start_lr=float(optimizer.start_lr),
                                 soft_coefficient=float(optimizer.soft_coefficient))
    if model is not None and hasattr(model, "dropout_state"):
        rng = model.dropout_state
        state["dropout"] = dict(seed=int(rng.seed), steps={dev: int(t.item()) for dev, t in rng._step.items()})
    return state


def restore_optimizer(optimizer, state: dict, model=None):
    inner = getattr(optimizer, "optimizer", optimizer)
    inner.load_state_dict(state["adam"])
    sched = state.get("schedule")
    if sched is not None and inner is not optimizer:
        optimizer.n_current_steps = int(sched["n_current_steps"])
        optimizer.start_lr = float(sched["start_lr"])
        optimizer.soft_coefficient = float(sched["soft_coefficient"])
        if hasattr(inner, "set_schedule_step") and "fused" not in state["adam"]:
            # moments written by a stock torch Adam: derive the device-side counter / lr from the host-side schedule
            inner.set_schedule_step(optimizer.n_current_steps, inner.param_groups[0]["lr"])
    if getattr(inner, "flat_shadow", None) is not None:
        inner.flat_shadow.copy_(inner.flat_param)                # bf16 operand copy follows the (re)loaded weights
    drop = state.get("dropout")
    if drop is not None and model is not None and hasattr(model, "dropout_state"):
        rng = model.dropout_state
        rng.seed = int(drop["seed"])
        device = next(model.parameters()).device
        steps = list(drop["steps"].values())
        if steps:
            rng.step_tensor(device).fill_(int(max(steps)))


def _cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu(v) for v in obj)
    return obj


# ------------------------------------------------------------------------------------------------ files
def save_state(path, state_dict, model, model_options, epoch, train_options=None, optimizer=None, extra=None) -> dict:
    """Write `state_dict` (not necessarily the live weights of `model`: a best-epoch snapshot, an average) atomically
    (temporary file + rename): a killed job never leaves a truncated `epoch.N.torch` behind."""
    ckpt = dict(format=FORMAT, state_dict=_cpu(dict(state_dict)), model_options=options_dict(model_options),
                epoch=int(epoch), train_options=options_dict(train_options),
                optimizer=optimizer_state(optimizer, model) if optimizer is not None else None, extra=dict(extra or {}))
    for name in _EXTRA_KWARGS[:2]:                               # what the constructor needs beyond the reference's flags
        if hasattr(model, name):
            ckpt["model_options"].setdefault(name, getattr(model, name))
    if hasattr(model, "dropout_state"):
        ckpt["model_options"].setdefault("seed", model.dropout_state.seed)
    tmp = "%s.tmp.%d" % (path, os.getpid())
    torch.save(ckpt, tmp)
    os.replace(tmp, path)
    return ckpt


def save_checkpoint(path, model, model_options, epoch, train_options=None, optimizer=None, extra=None) -> dict:
    """Checkpoint of the live model (+ optimiser / schedule / dropout state when `optimizer` is given)."""
    return save_state(path, model.state_dict(), model, model_options, epoch, train_options, optimizer, extra)


def _pickle_allowed(allow_pickle) -> bool:
    if allow_pickle is not None:
        return bool(allow_pickle)
    return os.environ.get("PKA_ALLOW_PICKLE", "1") != "0"


def read_checkpoint(path, allow_pickle=None) -> dict:
    """File -> {'state_dict', 'model_options', 'epoch', 'train_options', 'optimizer', 'extra'} without building a model.
    Files of this package are plain data and load with `weights_only=True`.  Reference-format files (a pickled
    nn.Module + argparse.Namespace, L/train.py:252-262) can only be read by unpickling, which executes code from the
    file: that path is taken only when the safe loader rejects the file *as a pickle of non-plain objects*, and only if
    allowed (`allow_pickle=True`, or unset with PKA_ALLOW_PICKLE != "0"; trusted local files only, like the reference).
    A missing, truncated or otherwise unreadable file raises the original error instead of being retried unsafely."""
    import pickle
    try:
        raw = torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError as exc:
        # torch raises UnpicklingError("Weights only load failed ... Unsupported global ...") for pickled classes
        if "Unsupported" not in str(exc) and "weights_only" not in str(exc).lower() and "Weights only" not in str(exc):
            raise
        if not _pickle_allowed(allow_pickle):
            raise RuntimeError("[ERROR] %s holds pickled objects (reference checkpoint format); pass allow_pickle=True "
                               "or set PKA_ALLOW_PICKLE=1 to unpickle it (only for files you trust)" % path) from exc
        from . import dropin                                     # noqa: F401  (registers transformer.*, TDNN, utils.*)
        raw = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(raw, dict):
        raise ValueError("[ERROR] %s is not a checkpoint dictionary" % path)
    if raw.get("format") == FORMAT:
        return raw
    if "model" in raw and isinstance(raw["model"], torch.nn.Module):
        return dict(format="reference-pickled-module", state_dict=_cpu(dict(raw["model"].state_dict())),
                    model_options=options_dict(raw.get("model_options")), epoch=int(raw.get("epoch", 0)),
                    train_options=options_dict(raw.get("train_options")), optimizer=None, extra={})
    raise ValueError("[ERROR] %s: unknown checkpoint layout (keys: %s)" % (path, sorted(raw)))


def load_checkpoint(path, device=None, allow_pickle=None) -> dict:
    """-> the checkpoint dictionary plus 'model': a Transformer of this package with the weights loaded."""
    ckpt = dict(read_checkpoint(path, allow_pickle))
    ckpt["model"] = build_model(ckpt["model_options"], ckpt["state_dict"], device)
    return ckpt


# ------------------------------------------------------------------------------------------------ averaging
def running_average(state_dicts: Iterable[Dict[str, torch.Tensor]]) -> Iterator[Tuple[int, Dict[str, torch.Tensor]]]:
    """Yield (n, mean of the first n state dicts) for n = 1, 2, ... with the reference's arithmetic
    (L/train.py:276-303: avg <- avg * (1 - 1/n) + sd_n * (1/n), in the tensors' own dtype)."""
    avg = None
    for n, sd in enumerate(state_dicts, 1):
        if avg is None:
            avg = {k: v.detach().clone() for k, v in sd.items()}
        else:
            if set(sd) != set(avg):
                raise ValueError("[ERROR] cannot average checkpoints with different parameter names")
            factor = 1 / n
            avg = {k: avg[k].mul(1 - factor).add(sd[k].to(avg[k].device), alpha=factor) for k in avg}
        yield n, avg
