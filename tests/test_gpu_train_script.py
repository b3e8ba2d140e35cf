"""train.py (the reference-shaped driver) end to end on synthetic data: 3 tiny epochs with FDS + LDS, i.e. the
whole state machine -- epoch 0 collect, epoch 1 smooth with identity tables, epoch 2 live calibration -- plus
checkpoint writing and the validation pass; once with the datasets' float tensors and once with --device_transform (uint8
images, RandomCrop / flip / ToTensor / Normalize on the GPU per batch)."""
import os
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("extra", [[], ["--device_transform"]], ids=["host_tensors", "device_transform"])
def test_train_script_runs_three_epochs(tmp_path, extra):
    import train
    argv = extra + ["--synthetic", "192", "--batch_size", "64", "--epoch", "3", "--fds", "--lds", "--reweight", "sqrt_inv",
            "--lds_ks", "5", "--lds_sigma", "2", "--fds_ks", "5", "--fds_sigma", "2", "--bucket_num", "101",
            "--bucket_start", "0", "--workers", "0", "--img_size", "64", "--print_freq", "1",
            "--store_root", str(tmp_path), "--lr", "1e-4"]
    train.main(argv)
    runs = os.listdir(tmp_path)
    assert len(runs) == 1
    ckpt = torch.load(os.path.join(tmp_path, runs[0], "ckpt.pth.tar"), map_location="cpu")
    assert ckpt["epoch"] == 3
    sd = ckpt["state_dict"]
    assert "module.FDS.smoothed_mean_last_epoch" in sd and "module.layer4.2.bn3.running_var" in sd
    assert int(sd["module.FDS.epoch"][0]) == 2                    # two update_last_epoch_stats transitions
    assert float(sd["module.FDS.num_samples_tracked"].sum()) == 3 * 192
    assert torch.isfinite(sd["module.conv1.weight"]).all()
    # the tables became non-trivial and smoothing changed them
    assert not torch.allclose(sd["module.FDS.smoothed_mean_last_epoch"], torch.zeros_like(sd["module.FDS.smoothed_mean_last_epoch"]))
