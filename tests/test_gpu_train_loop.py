"""T1 parity: `train.train()` (the mirror of agedb-dir/train.py:234-283) against the oracle's restatement of the
epoch-end FDS refresh (train.py:269-281) over three epochs.

Teacher-forced on the features: the backbone's raw encodings of the collection pass are captured on the device and
replayed through the oracle in the REFERENCE's order --

    for every batch:   feature = FDS.smooth(encoding, targets, epoch)     (train-mode forward, old tables; resnet.py:142-144)
    then               FDS.update_last_epoch_stats(epoch)                  (train.py:280)
    then               FDS.update_running_stats(all features, labels, epoch)   (train.py:281)

-- and after every epoch all eight FDS buffers must agree (rtol 1e-5, the north-star tolerance), as must the smoothed
features each collection batch fed into the accumulators.  A driver that refreshes the last-epoch tables BEFORE the
collection pass (round 1's bug) fails here at epoch 1: its collection features are calibrated with the new tables.
"""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from oracle import dir_oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu

BUFFERS = ("running_mean", "running_var", "running_mean_last_epoch", "running_var_last_epoch",
           "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked")


def _setup(bucket_num, bucket_start, start_update, start_smooth, n=96, batch=32, img=64, momentum=0.9):
    import train
    from resnet import ResNet, Bottleneck
    from parallel import DataParallel, ShardSampler
    from optim import FusedAdam
    args = train.build_parser().parse_args(
        ["--fds", "--lds", "--reweight", "sqrt_inv", "--lds_ks", "5", "--lds_sigma", "2", "--fds_ks", "5",
         "--fds_sigma", "2", "--bucket_num", str(bucket_num), "--bucket_start", str(bucket_start),
         "--start_update", str(start_update), "--start_smooth", str(start_smooth), "--print_freq", "1000",
         "--batch_size", str(batch), "--img_size", str(img), "--lr", "1e-4"])
    torch.manual_seed(0)
    net = ResNet(Bottleneck, [1, 1, 1, 1], fds=True, bucket_num=bucket_num, bucket_start=bucket_start,
                 start_update=start_update, start_smooth=start_smooth, kernel="gaussian", ks=5, sigma=2,
                 momentum=momentum).cuda()
    model = DataParallel(net)
    ds = train.SyntheticAges(n, img, args, seed=3)
    # labels beyond both table edges so that the edge folding of fds.py:94-97 is exercised
    ds.labels[:6] = np.asarray([0., 1., float(bucket_start), float(bucket_num - 1), float(bucket_num + 5), 120.], np.float32)
    loader = DataLoader(ds, batch_size=batch, sampler=ShardSampler(n, 0, 1, shuffle=False, pad=True), num_workers=0)
    opt = FusedAdam([p for p in model.parameters()], lr=args.lr)
    ref = O.FDSState(2048, bucket_num, bucket_start, start_update=start_update, start_smooth=start_smooth,
                     kernel="gaussian", ks=5, sigma=2, momentum=momentum)
    return train, args, net, model, loader, opt, ref


@pytest.mark.parametrize("cfg", [(100, 3, 0, 1), (100, 0, 0, 1), (101, 0, 1, 2)],
                         ids=["agedb_100_3", "imdbwiki_100_0", "late_start_101_0"])
def test_train_loop_fds_refresh_matches_reference_order(cfg):
    bucket_num, bucket_start, start_update, start_smooth = cfg
    train, args, net, model, loader, opt, ref = _setup(*cfg)
    fds = net.FDS
    raw, fed, state = [], [], {"collecting": False}
    run_forward, accumulate, begin = net._run_forward, fds.accumulate_batch, fds.begin_epoch_stats

    def rec_begin(all_labels):                           # the collection pass starts here (train.py:269-272) ...
        state["collecting"] = True
        return begin(all_labels)

    def rec_forward(x, training):
        enc = run_forward(x, training)
        if state["collecting"]:
            raw.append(enc.detach().clone())             # before FDS.smooth touches it in place
        return enc

    def rec_accumulate(features, labels):
        fed.append((features.detach().clone(), labels.detach().clone().reshape(-1)))
        return accumulate(features, labels)

    net._run_forward, fds.accumulate_batch, fds.begin_epoch_stats = rec_forward, rec_accumulate, rec_begin
    for epoch in range(start_update + 3):
        raw.clear()
        fed.clear()
        state["collecting"] = False                      # ... and ends with train()
        loss = train.train(loader, model, opt, epoch, args)
        assert np.isfinite(loss)
        if epoch < start_update:
            assert not raw and not fed                   # train.py:269: no collection before start_update
            continue
        assert len(raw) == len(fed) == len(loader)
        feats, labs = [], []
        for enc, (got, lab) in zip(raw, fed):
            lab_np = lab.cpu().numpy()
            want = ref.smooth(enc.cpu().numpy().copy(), lab_np, epoch) if epoch >= start_smooth else enc.cpu().numpy()
            assert_close(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6, what=f"collection features, epoch {epoch}")
            feats.append(want)
            labs.append(lab_np)
        ref.update_last_epoch_stats(epoch)                                        # train.py:280
        ref.update_running_stats(np.vstack(feats), np.hstack(labs), epoch)        # train.py:281
        assert int(fds.epoch.item()) == ref.epoch == fds._epoch_host
        for k in BUFFERS:
            assert_close(getattr(fds, k).cpu().numpy(), getattr(ref, k), rtol=1e-5, atol=1e-6, what=f"{k}, epoch {epoch}")
        # the alias of fds.py:55-56 holds on both sides once the first transition happened
        if ref.epoch > start_update:
            assert fds.running_mean_last_epoch.data_ptr() == fds.running_mean.data_ptr()
            assert ref.running_mean_last_epoch is ref.running_mean
    assert ref.epoch == start_update + 2                 # two transitions in three collecting epochs


def test_retrain_fc_step_updates_only_the_regressor():
    """--retrain_fc (train.py:156-160): frozen backbone, the optimizer holds linear.{weight,bias} only.  The runner's
    backward never runs, yet the two gradients must be views of the flat gradient buffer for the fused optimizer."""
    train, args, net, model, loader, opt, ref = _setup(100, 3, 0, 1, n=32, batch=32)
    from optim import FusedSGD
    from loss import weighted_l1_loss
    for name, p in model.named_parameters():
        if 'fc' not in name and 'linear' not in name:
            p.requires_grad = False
    params = [p for p in model.parameters() if p.requires_grad]
    assert len(params) == 2
    opt = FusedSGD(params, lr=0.1, momentum=0.9, weight_decay=1e-4)
    before = net.flat_parameters().clone()
    model.train()
    x, t, w = next(iter(loader))
    x, t, w = x.cuda(), t.cuda(), w.cuda()
    out, _ = model(x, t, 0)
    loss = weighted_l1_loss(out, t, w)
    opt.zero_grad()
    loss.backward()
    model.reduce_gradients()
    opt.step()
    after = net.flat_parameters()
    nb = net._flat["backbone"]
    assert torch.equal(before[:nb], after[:nb])
    assert not torch.equal(before[nb:], after[nb:])
    assert net.linear.weight.grad.data_ptr() == net.flat_grads()[nb:].data_ptr()
