"""CPU (`-m "not gpu"`): host-side logic of the package -- batch contract, symbolic masks, position tables, weight
interchange with the reference, LR schedule, gradient bucketing and the world-size-2 all-reduce path (gloo)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden_state_dict, load_golden
from oracle import acoustic_model as am

import pytorch_kaldi_asr_b200 as pk
from pytorch_kaldi_asr_b200 import parallel
from pytorch_kaldi_asr_b200.transformer import Models
from pytorch_kaldi_asr_b200.utils import constants, synthetic
from pytorch_kaldi_asr_b200.utils.instances_handler import pad_to_longest

SMALL = dict(n_src_dim=8, n_tgt_vocab=11, encoder_max_len=40, decoder_max_len=24, src_fold=1,
             encoder_sub_sequence=(-100, 0), decoder_sub_sequence=(-3, 0), en_layers=2, de_layers=2, n_head=2,
             en_d_model=32, de_d_model=32, d_k=16, d_v=16, en_dropout=0.0, de_dropout=0.0,
             tdnn_contexts=[[-1, 0, 1], [-3, 0, 3]])


def test_constants_and_pad_polarity():
    assert (constants.PAD, constants.UNK, constants.BOS, constants.EOS) == (0, 1, 2, 3)
    data, mask = pad_to_longest([np.ones((3, 2), np.float32), np.ones((5, 2), np.float32)])
    assert data.shape == (2, 5, 2) and mask.dtype == np.uint8
    assert mask.tolist() == [[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]] and not data[0, 3:].any()
    lab, lmask = pad_to_longest([np.array([2, 5, 3]), np.array([2, 3])])
    assert lab.tolist() == [[2, 5, 3], [2, 3, 0]] and lmask.tolist() == [[1, 1, 1], [1, 1, 0]]


def test_synthetic_batches_follow_the_survey_recipe():
    b = synthetic.batches(2, 4, seed=1234)
    keys, src, smask, tgt, tmask = b[0]
    assert src.dtype == np.float32 and src.shape[2] == 40 and tgt.dtype == np.int64
    lens = smask.sum(1)
    assert lens.min() >= 90 and lens.max() <= 499 and src.shape[1] == lens.max()
    assert (tgt[:, 0] == constants.BOS).all()
    for row, m in zip(tgt, tmask):
        n = int(m.sum())
        assert row[n - 1] == constants.EOS and (row[1:n - 1] >= 4).all() and (row[1:n - 1] <= 51).all()
    same = synthetic.batches(2, 4, seed=1234, pad_to="set")
    assert same[0][1].shape[1] == same[1][1].shape[1]
    assert np.array_equal(same[0][1][:, :src.shape[1]], src)


def test_position_table_and_symbolic_masks_match_the_oracle():
    assert torch.equal(Models.position_encoding_init(50, 16), am.sinusoid_table(50, 16))
    kp = torch.tensor([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 0, 0]], dtype=torch.uint8)
    m = Models.get_attn_padding_mask(kp, kp) + Models.get_attn_subsequent_mask(kp, -2, 0)
    assert torch.equal(m.dense(), am.attention_mask(6, kp, (-2, 0)))
    g = load_golden("semantics")
    assert np.array_equal(m.dense().numpy(), g["mask_m2_0"].astype(bool))


@pytest.mark.parametrize("fold", [2, 3])
def test_fold_seq_and_mask_matches_reference(fold):
    g = load_golden("semantics")
    x = torch.from_numpy(g["concat_in"])
    m = torch.tensor([[1] * 7, [1, 1, 1, 1, 0, 0, 0]], dtype=torch.uint8)
    s, fm = Models.fold_seq_and_mask(x, m, fold)
    assert np.array_equal(s.numpy(), g["fold%d_seq" % fold]) and np.array_equal(fm.numpy(), g["fold%d_mask" % fold])


def test_state_dict_keys_shapes_and_seeded_init_equal_the_reference():
    """Weight interchange: same keys/shapes, and -- because the constructors consume torch's RNG in the
    reference's order -- bit-identical initial weights under the same seed (the golden was built with seed 1)."""
    g = load_golden("tdnn_small_fwd_bwd")
    torch.manual_seed(1)
    model = pk.Transformer(lda_mat=g["lda_mat"], **SMALL)
    sd = model.state_dict()
    ref = golden_state_dict(g)
    assert sorted(sd) == sorted(ref)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
        assert torch.equal(sd[k], ref[k]), k
    frozen = [k for k, p in model.named_parameters() if not p.requires_grad]
    assert sorted(frozen) == sorted(k for k in ref if k not in am.trainable_keys(ref))
    # reference quirk: the decoder heads ignore the Transformer's d_k/d_v (T/Models.py:250-251)
    assert sd["decoder.layer_stack.0.slf_attn.w_qs"].shape == (2, 32, 64)


def test_dropout_sites_are_the_25_of_the_example_model():
    model = pk.Transformer(lda_mat=synthetic.lda_matrix(), **{k: v for k, v in am.example_config().items() if k != "encoder_type"})
    assert len(model.dropout_sites) == 25
    plan = am.DropoutPlan("off")
    sd = {k: v for k, v in model.state_dict().items()}
    b = synthetic.batches(1, 2, seed=3, mean_len=100, std_len=5, min_len=90, max_len=110)[0]
    am.transformer_forward(sd, am.example_config(), torch.from_numpy(b[1]), torch.from_numpy(b[2]),
                           torch.from_numpy(b[3][:, :-1]), torch.from_numpy(b[4][:, :-1]), plan)
    assert sorted(set(plan.visited)) == sorted(model.dropout_sites)


def test_no_cpu_fallback_anywhere():
    model = pk.Transformer(lda_mat=synthetic.lda_matrix(), **{k: v for k, v in am.example_config().items() if k != "encoder_type"})
    b = synthetic.batches(1, 2, seed=3, mean_len=100, std_len=5, min_len=90, max_len=110)[0]
    with pytest.raises(RuntimeError):
        model(torch.from_numpy(b[1]), torch.from_numpy(b[2]), torch.from_numpy(b[3][:, :-1]), torch.from_numpy(b[4][:, :-1]))

    class Loader(list):
        mode = "drop"
    with pytest.raises(RuntimeError):
        pk.train_epoch(model, Loader([b]), None, mode="eval")
    from pytorch_kaldi_asr_b200.decode import translate_batch
    import types
    with pytest.raises(RuntimeError):
        translate_batch(model, b, types.SimpleNamespace(beam_size=2, max_token_seq_len=3, nbest=1), None)
    with pytest.raises((RuntimeError, ValueError)):
        pk.FusedAdam(model.parameters())


def test_scheduled_optim_host_schedule_with_a_plain_torch_optimizer():
    p = torch.nn.Parameter(torch.zeros(3))
    opt = pk.ScheduledOptim(torch.optim.Adam([p], betas=(0.9, 0.999), eps=1e-8), start_lr=2e-3, soft_coefficient=10)
    assert opt.optimizer.param_groups[0]["lr"] == 1e-3            # step 1 runs at Adam's constructor default
    for n in range(1, 4):
        p.grad = torch.ones(3)
        opt.step()
        opt.update_learning_rate()
        assert opt.optimizer.param_groups[0]["lr"] == pytest.approx(2e-3 * 10 / (n + 10))
    assert opt.n_current_steps == 3


def test_bucket_plan_covers_the_arena_on_tensor_boundaries():
    sizes = [100, 40, 8, 300, 52, 20]
    offsets = np.cumsum([0] + sizes[:-1]).tolist()
    bounds, owner = parallel.plan_buckets(offsets, sizes, sum(sizes), 3)
    assert bounds[0][0] == 0 and bounds[-1][1] == sum(sizes)
    assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
    assert all(lo in offsets for lo, _ in bounds)
    assert owner == sorted(owner) and len(set(owner)) == len(bounds)
    assert parallel.shard_range(10, 0, 4) == (0, 3) and parallel.shard_range(10, 3, 4) == (8, 10)
    assert sum(hi - lo for lo, hi in (parallel.shard_range(1000, r, 8) for r in range(8))) == 1000


class _FakeArena:
    """Stands in for FusedAdam on CPU: the flat gradient arena + its views (the real one needs CUDA)."""

    def __init__(self, shapes):
        self._train = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
        self._offsets, total = [], 0
        for p in self._train:
            self._offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.flat_grad = torch.zeros(total)
        for p, off in zip(self._train, self._offsets):
            p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)


def _dp_worker(rank, world, port, overlap, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    arena = _FakeArena([(7, 5), (3,), (4, 4), (11,)])
    sync = parallel.GradAllReduce(arena, n_buckets=2, overlap=overlap)
    x = torch.full((5,), float(rank + 1))
    # a loss whose gradient differs per rank; SUM (not mean) semantics expected
    loss = (arena._train[0] @ x).sum() * (rank + 1) + arena._train[1].sum() * 2 + (arena._train[2] ** 2).sum() \
        + arena._train[3].sum() * (rank + 3)
    loss.backward()
    sync.finish()
    stats = parallel.all_reduce_stats(torch.tensor([1.0 + rank, 2.0, 3.0]))
    out[rank] = (arena.flat_grad.clone(), stats)
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
def test_gradient_allreduce_is_a_sum_over_ranks_world2_gloo(overlap):
    world = 2
    port = 29500 + os.getpid() % 2000 + (1 if overlap else 0)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(world, port, overlap, out), nprocs=world, join=True)
    g0, s0 = out[0]
    g1, s1 = out[1]
    assert torch.equal(g0, g1)                                    # both ranks hold the reduced gradient
    arena = _FakeArena([(7, 5), (3,), (4, 4), (11,)])
    want0 = torch.ones(7, 5) * (1.0 * 1 + 2.0 * 2)                # sum_r (r+1) * x_r
    assert torch.allclose(g0[:35].view(7, 5), want0)
    off = arena._offsets
    assert torch.allclose(g0[off[1]:off[1] + 3], torch.full((3,), 4.0))      # 2 + 2, NOT averaged
    assert torch.allclose(g0[off[3]:off[3] + 11], torch.full((11,), 7.0))    # (0+3) + (1+3)
    assert s0.tolist() == [3.0, 4.0, 6.0] and s1.tolist() == s0.tolist()


def test_no_cache_decoder_reads_prefixes_off_the_lattice_arrays():
    """decode.BeamDecoder._slot_prefixes (the no-cache path of a non-causal decoder band): token prefixes of the live beam
    slots walked along the back-pointers of the device lattice arrays.  Host-side index logic, checked on CPU tensors
    against the oracle lattice (T/Lattice.py:84-107) advanced with random scores; idle slots give all-BOS rows."""
    from oracle.lattice import BeamLattice
    from pytorch_kaldi_asr_b200.decode import BeamDecoder
    rng = np.random.RandomState(7)
    beam, V, max_len, n_utt = 4, 9, 6, 3
    lats = [BeamLattice(max_len, beam) for _ in range(n_utt)]
    for step in range(4):
        E = 1 + beam * max_len
        prev = torch.full((n_utt, E), -1, dtype=torch.int32)
        word = torch.zeros((n_utt, E), dtype=torch.int32)
        slot_edge = torch.zeros((n_utt, beam), dtype=torch.int32)
        for u, lat in enumerate(lats):
            prev[u, :len(lat.prev)] = torch.tensor(lat.prev, dtype=torch.int32)
            word[u, :len(lat.word)] = torch.tensor(lat.word, dtype=torch.int32)
            live = lat.active()
            slot_edge[u, :len(live)] = torch.tensor(live, dtype=torch.int32)
        bd = BeamDecoder.__new__(BeamDecoder)                      # index logic only: no model, no device
        bd.n_utt, bd.beam, bd.dev = n_utt, beam, "cpu"
        bd.slot_edge, bd.edge_prev, bd.edge_word = slot_edge, prev, word
        toks = bd._slot_prefixes(step + 1).view(n_utt, beam, step + 1)
        for u, lat in enumerate(lats):
            seqs, _ = lat.get_results("active")
            for k, seq in enumerate(seqs):
                assert toks[u, k].tolist() == seq
            for k in range(len(seqs), beam):
                assert toks[u, k].tolist() == [constants.BOS] * (step + 1)
        for lat in lats:                                           # next step: random scores, EOS made unlikely
            lp = rng.randn(max(1, lat.num_curr_active), V).astype(np.float32)
            lp[:, constants.EOS] -= 3.0
            lat.advance(lp)
