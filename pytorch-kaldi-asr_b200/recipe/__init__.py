"""Command-line entry points with the flags of the reference's recipe scripts:

    python -m pytorch_kaldi_asr_b200.recipe.initialize_model   (L/initialize_model.py)
    python -m pytorch_kaldi_asr_b200.recipe.train              (L/train.py main; `torchrun --nproc-per-node N` for N GPUs)
    python -m pytorch_kaldi_asr_b200.recipe.combine            (L/combine.py)
    python -m pytorch_kaldi_asr_b200.recipe.decode             (L/decode.py main)
    python -m pytorch_kaldi_asr_b200.recipe.rescore            (L/rescore.py)
    python -m pytorch_kaldi_asr_b200.recipe.score              (compute-wer + best_wer.sh of P/run.sh:196-203)

Each module has `main(argv=None)`.  Devices come from LOCAL_RANK (one process per GPU) instead of the reference's
nvidia-smi polling (U/get_gpu.py); `-use_gpu` is accepted and implied -- this path has no CPU mode.
"""
import os


def pick_device():
    """-> (torch.device, rank, world).  Under torchrun: cuda:LOCAL_RANK; otherwise the current CUDA device."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('[ERROR] no cuda device available!')
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', torch.cuda.current_device()))
    torch.cuda.set_device(local)
    print('[INFO] use gpu device {}'.format(local))
    return torch.device('cuda', local), rank, world
