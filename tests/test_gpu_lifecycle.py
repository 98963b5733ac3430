"""run.py's training lifecycle replayed against the drop-in modules (the reference script itself cannot run here: it
hard-codes dataset paths, run.py:48,84): the epoch loop (run.py:457-487), the train -> eval weight hand-off and the
per-epoch evaluation (run.py:493, 501-526), learning-rate and BatchNorm-momentum decay (run.py:548-556), checkpoint save
(run.py:559-569) and resume with optimiser state and generator RNG (run.py:411-417, 436-445). A run interrupted after
epoch 1 and resumed from its checkpoint (fresh model, optimiser and generator objects) must reproduce the loss trajectory
of the uninterrupted run."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from common.loss import mpjpe  # noqa: E402
from common.models.TemporalModel import TemporalModel, TemporalModelBase  # noqa: E402
from vp3d_b200.evaluation import evaluate  # noqa: E402
from vp3d_b200.feeder import DeviceSequenceFeeder, DeviceWindowFeeder  # noqa: E402
from vp3d_b200.optim import FusedAdam  # noqa: E402

FW = [3, 3, 3]
CH = 256
LR, LR_DECAY = 1e-3, 0.95                     # common/arguments.py:37-38
INITIAL_MOMENTUM, FINAL_MOMENTUM = 0.1, 0.001  # run.py:428-429
EPOCHS = 3


def _sequences(seed, n_seq, lo, hi, J=17):
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi, n_seq)
    X = [(rng.normal(0, 0.3, (n, J, 3)) + np.array([0, 0, 4.0])).astype(np.float32) for n in lens]
    Q = []
    for n in lens:
        q = np.array([1, 0, 0, 0], np.float32) + rng.normal(0, 0.05, (n, 4)).astype(np.float32)
        Q.append((q / np.linalg.norm(q, axis=-1, keepdims=True)).astype(np.float32))
    T = [rng.normal(0, 0.1, (n, 3)).astype(np.float32) for n in lens]
    cam = np.tile(np.array([2.29, 2.2876, 0.0251, 0.0289, -0.2071, 0.2478, -0.0031, -0.00098, -0.0014], np.float32),
                  (n_seq, 1))
    return X, Q, T, cam


def _models():
    """run.py:294-300: two instances of the same class, one in train() mode with dropout, one for evaluation."""
    torch.manual_seed(7)
    model_pos_train = TemporalModel(17, 2, 17, FW, causal=False, dropout=0.0, channels=CH).cuda()
    model_pos = TemporalModel(17, 2, 17, FW, causal=False, dropout=0.0, channels=CH).cuda()
    return model_pos_train, model_pos


def _run(tmp_path, stop_after=None, resume_from=None):
    """The body of run.py's train() for `EPOCHS` epochs; returns ({epoch: [losses]}, {epoch: eval MPJPE mm}, checkpoint)."""
    pad = (27 - 1) // 2
    Xtr, Qtr, Ttr, camtr = _sequences(1, 6, 60, 140)
    Xte, Qte, Tte, camte = _sequences(2, 3, 50, 90)
    train_generator = DeviceWindowFeeder(Xtr, Qtr, Ttr, camtr, batch_size=128, chunk_length=1, pad=pad, shuffle=True,
                                         random_seed=1234)
    test_generator = DeviceSequenceFeeder(Xte, Qte, Tte, camte, pad=pad)
    model_pos_train, model_pos = _models()
    lr = LR
    optimizer = FusedAdam(model_pos_train.parameters(), lr=lr, amsgrad=True)          # run.py:662
    epoch = 0
    if resume_from is not None:
        checkpoint = torch.load(resume_from, weights_only=False)                       # run.py:411-417
        model_pos_train.load_state_dict(checkpoint['model_pos'])
        model_pos.load_state_dict(checkpoint['model_pos'])
        epoch = checkpoint['epoch']                                                    # run.py:436-445
        optimizer.load_state_dict(checkpoint['optimizer'])
        train_generator.set_random_state(checkpoint['random_state'])
        lr = checkpoint['lr']
    lr_decay = LR_DECAY
    momentum = INITIAL_MOMENTUM * np.exp(-epoch / EPOCHS * np.log(INITIAL_MOMENTUM / FINAL_MOMENTUM))
    if epoch > 0 and isinstance(model_pos_train, TemporalModelBase):
        model_pos_train.set_bn_momentum(momentum)
    losses, evals, last_ckpt = {}, {}, None
    while epoch < EPOCHS:
        model_pos_train.train()                                                        # run.py:455
        ep_losses = []
        for _cams, batch_3d, batch_2d in train_generator.next_epoch():                 # run.py:457
            optimizer.zero_grad()                                                      # run.py:467
            predicted_3d_pos = model_pos_train(batch_2d)                               # run.py:473
            loss_3d_pos = mpjpe(predicted_3d_pos, batch_3d)                            # run.py:480
            ep_losses.append(loss_3d_pos.item())                                       # run.py:481
            loss_3d_pos.backward()                                                     # run.py:485
            optimizer.step()                                                           # run.py:487
        losses[epoch] = ep_losses
        model_pos.load_state_dict(model_pos_train.state_dict())                        # run.py:493
        evals[epoch] = evaluate(model_pos, test_generator)['e1']                       # run.py:501-526
        lr *= lr_decay                                                                 # run.py:548-550
        for param_group in optimizer.param_groups:
            param_group['lr'] *= lr_decay
        epoch += 1
        momentum = INITIAL_MOMENTUM * np.exp(-epoch / EPOCHS * np.log(INITIAL_MOMENTUM / FINAL_MOMENTUM))
        model_pos_train.set_bn_momentum(momentum)                                      # run.py:554-556
        chk_path = os.path.join(str(tmp_path), 'epoch_{}.bin'.format(epoch))           # run.py:559-569
        torch.save({'epoch': epoch, 'lr': lr, 'random_state': train_generator.random_state(),
                    'optimizer': optimizer.state_dict(), 'model_pos': model_pos_train.state_dict()}, chk_path)
        last_ckpt = chk_path
        if stop_after is not None and epoch >= stop_after:
            break
    return losses, evals, last_ckpt, model_pos_train


def test_train_eval_checkpoint_resume_like_run_py(tmp_path):
    full_losses, full_evals, _, m_full = _run(tmp_path)
    assert sorted(full_losses) == [0, 1, 2]
    flat = [v for e in sorted(full_losses) for v in full_losses[e]]
    assert all(np.isfinite(flat))
    assert np.mean(full_losses[2]) < 0.8 * np.mean(full_losses[0][:3]), 'training does not reduce the loss'
    assert all(np.isfinite(v) and v > 0 for v in full_evals.values())
    assert full_evals[2] < 1.1 * full_evals[0]
    # BatchNorm momentum decayed as run.py:554-556 sets it, on every BatchNorm of the stack
    want = INITIAL_MOMENTUM * np.exp(-3 / EPOCHS * np.log(INITIAL_MOMENTUM / FINAL_MOMENTUM))
    assert abs(m_full.expand_bn.momentum - want) < 1e-12 and all(abs(bn.momentum - want) < 1e-12 for bn in m_full.layers_bn)
    n_steps = sum(len(v) for v in full_losses.values())
    assert int(m_full.expand_bn.num_batches_tracked.item()) == n_steps

    # "interrupted" after epoch 1: a fresh set of objects resumes from the checkpoint the run above wrote at that point
    # (comparing against a second from-scratch run would also measure how two runs drift apart through the order of
    # their floating-point atomics, amplified by 16-bit roundings -- ~1e-3 after a handful of steps)
    ckpt = os.path.join(str(tmp_path), 'epoch_1.bin')
    ck = torch.load(ckpt, weights_only=False)
    assert ck['epoch'] == 1 and abs(ck['lr'] - LR * LR_DECAY) < 1e-12
    assert set(ck['optimizer']['state'][0]) >= {'step', 'exp_avg', 'exp_avg_sq', 'max_exp_avg_sq'}     # amsgrad state
    # the checkpoint is interchangeable with stock torch.optim.Adam (FusedAdam is a subclass with the same state)
    stock = torch.optim.Adam(_models()[0].parameters(), lr=LR, amsgrad=True)
    stock.load_state_dict(ck['optimizer'])
    out_dir = tmp_path / 'resumed'
    out_dir.mkdir()
    resumed_losses, resumed_evals, _, m_res = _run(out_dir, resume_from=ckpt)
    assert sorted(resumed_losses) == [1, 2]
    for e in (1, 2):
        # same samples in the same order (generator RNG restored), same optimiser state: the trajectories agree up to
        # the order of floating-point atomics (statistics, split-K weight gradients)
        np.testing.assert_allclose(resumed_losses[e], full_losses[e], rtol=5e-3)
        assert abs(resumed_evals[e] - full_evals[e]) < 5e-3 * full_evals[e]
    for (k, a), (_, b) in zip(m_full.state_dict().items(), m_res.state_dict().items()):
        if a.dtype.is_floating_point:
            # in norm, not element by element: where a gradient is ~0 Adam's update is +-lr whatever its size, so the
            # order of the fp32 atomics moves single weights by a few lr (1e-3) between two runs
            # (tensors that start at 0 -- BatchNorm biases -- consist of nothing but such +-lr steps: RMS allowance 0.5 lr)
            assert (a - b).norm().item() <= 2e-2 * a.norm().item() + 0.5 * LR * a.numel() ** 0.5, k
            assert (a - b).abs().max().item() <= 8 * LR + 2e-2 * a.abs().max().item(), k
        else:
            assert int(a) == int(b), k
