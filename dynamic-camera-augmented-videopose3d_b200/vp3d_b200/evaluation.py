"""Evaluation loop of run.py on the device: the per-sequence metrics of `evaluate()`
(run.py:677-774: MPJPE, P-MPJPE, N-MPJPE, MPJVE, the pose-motion statistic) and the camera-motion correlation table of
`run_evaluation()` (run.py:946-983) for the TemporalModel path.

The reference moves every prediction to the host (`.cpu().numpy()`, run.py:749-750) to run the Procrustes and velocity
metrics in NumPy and calls `.item()` three times per sequence. Here every metric is a device kernel (common.loss on CUDA
tensors), the frame-weighted sums stay on the device, and the host reads ONE small tensor after the last sequence.

Generators: anything with the UnchunkedGenerator protocol (`next_epoch()` yielding `(cams, batch_3d, batch_2d,
seq_info)`), in particular vp3d_b200.feeder.DeviceSequenceFeeder (device tensors, no host staging at all) -- NumPy
batches as the reference's generator yields them are uploaded first.
"""
import numpy as np
import torch

from common import loss as closs

CAM_KEYS = ('cam_velocity', 'cam_acceleration', 'cam_angular_velocity', 'cam_angular_acceleration')   # run.py:946-953


def _to_device(a, dev):
    if isinstance(a, torch.Tensor):
        return a.to(dev, dtype=torch.float32)
    return torch.from_numpy(np.asarray(a).astype('float32')).to(dev)      # run.py:698-705


@torch.no_grad()
def evaluate(model, generator, device=None):
    """-> dict(e1, e2, e3, ev: frame-weighted means in millimetres like run.py:764-767; e1_per_seq: (n_seq,) metres,
    pose_motion_per_seq: (n_seq,), cam_info_per_seq: list of the generator's seq_info dicts)."""
    model.eval()
    dev = device if device is not None else next(model.parameters()).device
    sums = torch.zeros(4, dtype=torch.float64, device=dev)     # frames x {mpjpe, p_mpjpe, n_mpjpe, velocity}
    n_frames = 0
    e1_seq, motion_seq, infos = [], [], []
    for _cams, batch, batch_2d, seq_info in generator.next_epoch():
        x2d, x3d = _to_device(batch_2d, dev), _to_device(batch, dev)
        pred = model(x2d)                                                     # run.py:711
        frames = x3d.shape[0] * x3d.shape[1]
        e1 = closs.mpjpe(pred, x3d)                                           # run.py:734
        e3 = closs.n_mpjpe(pred, x3d)                                         # run.py:736
        flat_p = pred.reshape(-1, x3d.shape[-2], x3d.shape[-1])               # run.py:749-750, without the .cpu()
        flat_t = x3d.reshape(-1, x3d.shape[-2], x3d.shape[-1])
        e2 = closs.p_mpjpe(flat_p, flat_t)                                    # run.py:752
        ev = closs.mean_velocity_error(flat_p, flat_t)                        # run.py:756
        sums += frames * torch.stack([e1, e2, e3, ev]).double()
        n_frames += frames
        e1_seq.append(e1)
        # run.py:742-745: mean over frames and joints of |x3d[t+1] - x3d[t]| = the velocity "error" against a still pose
        motion_seq.append(closs.mean_velocity_error(flat_t, torch.zeros_like(flat_t)))
        infos.append(seq_info)
    mm = (sums / max(n_frames, 1) * 1000).cpu()                               # the only device -> host read
    return {'e1': float(mm[0]), 'e2': float(mm[1]), 'e3': float(mm[2]), 'ev': float(mm[3]), 'frames': n_frames,
            'e1_per_seq': torch.stack(e1_seq) if e1_seq else torch.zeros(0, device=dev),
            'pose_motion_per_seq': torch.stack(motion_seq) if motion_seq else torch.zeros(0, device=dev),
            'cam_info_per_seq': infos}


def camera_motion_pmcc(e1_per_seq, cam_info_per_seq, pose_motion_per_seq, reference_quirk=False):
    """Pearson correlation of the per-sequence MPJPE with |camera velocity|, |acceleration|, |angular velocity|,
    |angular acceleration| and the pose motion (run.py:946-983). Returns a dict keyed like the reference's printout.

    The reference calls `np.corrcoef(corr_data)` on the (n_seq, 6) table, which correlates ROWS (sequences) rather than
    the six columns, so what it prints as "PMCC (MPJPE and cam velocity)" is the correlation between the 6-vectors of
    sequences 0 and 1 (run.py:966-972). Default here: the column-wise coefficients the printout describes;
    `reference_quirk=True` reproduces the reference's numbers."""
    e1 = torch.as_tensor(e1_per_seq, dtype=torch.float64).reshape(-1)
    dev = e1.device
    cols = [e1]
    for k in CAM_KEYS:
        v = torch.as_tensor(np.array([np.asarray(info[k], dtype=np.float64) for info in cam_info_per_seq]), device=dev)
        cols.append(torch.linalg.norm(v.reshape(len(cam_info_per_seq), -1), dim=1))   # np.linalg.norm(..., axis=1)
    cols.append(torch.as_tensor(pose_motion_per_seq, dtype=torch.float64, device=dev).reshape(-1))
    table = torch.stack(cols, dim=1)                                           # (n_seq, 6), run.py:959-965
    corr = torch.corrcoef(table if reference_quirk else table.T)
    names = ('cam_velocity', 'cam_acceleration', 'cam_angular_velocity', 'cam_angular_acceleration', 'pose_motion')
    return {n: float(corr[0, i + 1]) for i, n in enumerate(names)}


def run_evaluation(model, generators_by_action, reference_quirk=False):
    """run.py:906-987 for a dict {action: generator}: action-wise averages of the four protocols and the PMCC table over
    all sequences of all actions."""
    per_action, e1_all, motion_all, infos_all = {}, [], [], []
    for action, gen in generators_by_action.items():
        r = evaluate(model, gen)
        per_action[action] = r
        e1_all.append(r['e1_per_seq'])
        motion_all.append(r['pose_motion_per_seq'])
        infos_all += r['cam_info_per_seq']
    out = {k: float(np.mean([r[k] for r in per_action.values()])) for k in ('e1', 'e2', 'e3', 'ev')}   # run.py:974-977
    out['per_action'] = per_action
    if infos_all and all(k in infos_all[0] for k in CAM_KEYS):
        out['pmcc'] = camera_motion_pmcc(torch.cat(e1_all), infos_all, torch.cat(motion_all), reference_quirk)
    return out


@torch.no_grad()
def sliding_window(model, inputs_2d, inputs_cam, window_size, max_windows=None):
    """The sliding-window evaluator of the camera-aware sibling models (CamLSTM.py:33-44, CamTransformer.py:72-91; called
    from run.py:511-512,518-519) for ANY model with their `model(win_2d, win_cam) -> (n, J_out, F_out)` protocol: every
    window of `window_size` consecutive frames of the ONE sequence in `inputs_2d` (1, T, J, F) / `inputs_cam` (1, T, 3, 4)
    becomes a batch element, in frame order. Returns (1, T - window_size + 1, J_out, F_out).
    The windows are overlapping strided views of the sequence (no gather kernel: frame i of window w is frame w + i);
    `max_windows` bounds how many are materialised per model call (the reference materialises all of them at once)."""
    _, t, j, _ = inputs_2d.shape
    n_windows = t - window_size + 1
    if n_windows <= 0:
        raise ValueError("window_size larger than sequence length")        # CamLSTM.py:36-37

    def windows(x, lo, hi):
        seq = x[0]
        s = seq.stride()
        return seq.as_strided((hi - lo, window_size) + tuple(seq.shape[1:]), (s[0], s[0]) + tuple(s[1:]),
                              seq.storage_offset() + lo * s[0])

    step = n_windows if not max_windows else int(max_windows)
    outs = []
    for lo in range(0, n_windows, step):
        hi = min(lo + step, n_windows)
        outs.append(model(windows(inputs_2d, lo, hi), windows(inputs_cam, lo, hi)))
    out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
    return out.reshape(1, n_windows, out.shape[-2], out.shape[-1])
