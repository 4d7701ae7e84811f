"""torchrun --nproc-per-node N tools/check_symm.py: the symmetric-memory gradient all-reduce equals NCCL's SUM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import pytorch_kaldi_asr_b200 as pk
from pytorch_kaldi_asr_b200 import parallel
torch.manual_seed(0)
params = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in (1000, 333, 1 << 20, 77)]
opt = pk.FusedAdam(params)
sync = parallel.GradAllReduce(opt, backend="symm")
g = torch.Generator(device="cuda").manual_seed(100 + rank)
for trial in range(3):
    local_grad = torch.randn(opt.numel, device="cuda", generator=g)
    ref = local_grad.clone()
    dist.all_reduce(ref)
    opt.flat_grad.copy_(local_grad)
    torch.cuda.synchronize(); dist.barrier()
    sync.finish()
    torch.cuda.synchronize()
    err = float((opt.flat_grad - ref).abs().max())
    print("rank %d trial %d op %s max abs diff vs NCCL sum: %.3g" % (rank, trial, sync.symm_op.__name__ if hasattr(sync.symm_op, "__name__") else sync.symm_op, err), flush=True)
    assert err < 1e-4 * float(ref.abs().max())
dist.barrier()
if rank == 0:
    print("symm all-reduce OK")
os._exit(0)
