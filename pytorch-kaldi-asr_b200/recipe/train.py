"""Training entry point -- flags and flow of L/train.py:333-402: load the initial model, Adam + the soft LR schedule,
loaders for train / dev / test, `train` for `-epoch` epochs, then `combine` around the best epoch.

One process per GPU under torchrun: every rank loads the same model, the training loader deals its batches round-robin
to the ranks (same seed everywhere), gradients are summed over NCCL inside the step (parallel.GradAllReduce); rank 0
evaluates, writes the checkpoints and runs `combine`."""
import argparse
import os


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('-read_train_dir', required=True)
    parser.add_argument('-read_dev_dir', required=True)
    parser.add_argument('-read_test_dir', required=True)
    parser.add_argument('-read_vocab_file', required=True)
    parser.add_argument('-load_model_file', required=True)
    parser.add_argument('-save_model_dir', required=True)
    parser.add_argument('-seq_error_prob', type=float, default=0)
    parser.add_argument('-epoch', type=int, default=50)
    parser.add_argument('-optim_start_lr', type=float, default=0.001)
    parser.add_argument('-optim_soft_coefficient', type=float, default=1000)
    parser.add_argument('-batch_size', type=int, default=64)
    parser.add_argument('-use_gpu', action='store_true')
    parser.add_argument('-save_interval', type=int, default=10)
    # beyond the reference
    parser.add_argument('-compute_mode', choices=('fp32', 'bf16'), default='fp32', help='bf16 = tensor-core path')
    parser.add_argument('-graphed', action='store_true', help='replay the whole step as one CUDA graph (fixed batch shape)')
    parser.add_argument('-shuffle_seed', type=int, default=None, help='seed of the epoch shuffles (required for >1 GPU)')
    parser.add_argument('-resume', action='store_true', help='continue from the optimiser state in -load_model_file')
    parser.add_argument('-dp_backend', choices=('nccl', 'peer'), default='nccl',
                        help='>1 GPU: NCCL all-reduce + Adam, or the fused peer-memory reduce-scatter/Adam/all-gather kernel')
    return parser


def main(argv=None):
    import torch
    from . import pick_device
    from .. import checkpoint, ops, parallel
    from .. import train as T
    from ..transformer.Optim import FusedAdam, ScheduledOptim
    from ..utils import instances_handler
    opt = build_parser().parse_args(argv)
    device, rank, world = pick_device()
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group('nccl', device_id=device)
        if opt.shuffle_seed is None:
            opt.shuffle_seed = 0
    print('[PROCEDURE] prepare trainning.')
    ops.set_compute_mode(opt.compute_mode)
    loaded = checkpoint.load_checkpoint(opt.load_model_file, device=device)
    model, model_options = loaded['model'], loaded['model_options']
    print('[INFO] loading model with parameter:\n\t{}'.format(model_options))
    crit = T.get_criterion(len(instances_handler.read_vocab(opt.read_vocab_file)))
    print('[INFO] using cross entropy loss.')
    optimizer = ScheduledOptim(FusedAdam(model.parameters(), betas=(0.9, 0.999), eps=1e-08),
                               start_lr=opt.optim_start_lr, soft_coefficient=opt.optim_soft_coefficient)
    print('[INFO] using adam as optimizer.')
    if opt.resume:
        if loaded.get('optimizer') is None:
            raise ValueError('[ERROR] -resume: {} holds no optimiser state'.format(opt.load_model_file))
        checkpoint.restore_optimizer(optimizer, loaded['optimizer'], model)
        opt.start_epoch = int(loaded['epoch']) + 1
        opt.resume_extra = dict(loaded.get('extra') or {})
        best_file = os.path.join(opt.save_model_dir, 'epoch.{}.torch'.format(opt.resume_extra.get('best_epoch', 0)))
        if opt.resume_extra.get('best_epoch', 0) > 0 and os.path.exists(best_file):
            opt.resume_extra['best_state'] = checkpoint.read_checkpoint(best_file)['state_dict']
    grad_sync = None
    if world > 1 and opt.dp_backend == 'peer':
        optimizer.optimizer.enable_peer_step()
    elif world > 1:
        grad_sync = parallel.GradAllReduce(optimizer.optimizer).finish

    def loader(directory, **kw):
        return T.initialize_batch_loader(directory + '/feats.scp', directory + '/text', opt.read_vocab_file, opt.batch_size, **kw)
    print('[INFO] reading training data...')
    train_data = loader(opt.read_train_dir, seed=opt.shuffle_seed, shard=(rank, world))
    print('[INFO] reading dev data...')
    dev_data = loader(opt.read_dev_dir)
    print('[INFO] reading test data...')
    test_data = loader(opt.read_test_dir)
    print('[INFO] batch loader is initialized')
    graphed = None
    if opt.graphed:
        example = next(iter(loader(opt.read_train_dir, seed=0)))
        graphed = T.GraphedTrainStep(model, optimizer, example, grad_sync=grad_sync)
    os.makedirs(opt.save_model_dir, exist_ok=True)
    print('[PROCEDURE] trainning start...')
    best_accu, best_epoch = T.train(model, train_data, dev_data, test_data, crit, optimizer, opt, model_options,
                                    graphed=graphed, grad_sync=grad_sync, writer=(rank == 0),
                                    stats_reduce=parallel.all_reduce_stats if world > 1 else None)
    if rank == 0:
        print('[PROCEDURE] combining start on best epoch {}'.format(best_epoch))
        best_accu = T.combine(opt, best_epoch, crit, dev_data, 30 if opt.epoch > 30 else opt.epoch)
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        torch.distributed.destroy_process_group()
    return best_accu


if __name__ == '__main__':
    main()
