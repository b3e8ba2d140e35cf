"""CPU half of the input-pipeline row: the random draws handed to dirb200_augment_batch reproduce, for the
same seed, the choices of the reference's per-sample torchvision chain (agedb-dir/datasets.py:38-45) -- the GPU test
(tests/test_gpu_input_pipeline.py) then checks the pixels bit for bit."""
import numpy as np
import torch


def test_draw_order_matches_torchvision_compose():
    from PIL import Image
    from torchvision import transforms
    import torchvision.transforms.functional as TF
    import datasets as D
    n, size = 9, 24
    imgs = np.random.RandomState(0).randint(0, 256, size=(n, size, size, 3)).astype(np.uint8)
    chain = transforms.Compose([transforms.RandomCrop(size, padding=16), transforms.RandomHorizontalFlip(),
                                transforms.ToTensor(), transforms.Normalize([.5] * 3, [.5] * 3)])
    torch.manual_seed(7)
    ref = torch.stack([chain(Image.fromarray(im)) for im in imgs])
    torch.manual_seed(7)
    crop, flip = D.draw_augment_params(n, size, 16)
    assert crop.dtype == torch.int32 and flip.dtype == torch.uint8
    assert int(crop.min()) >= 0 and int(crop.max()) <= 32
    out = []
    for k in range(n):
        im = TF.crop(TF.pad(Image.fromarray(imgs[k]), 16), int(crop[k, 0]), int(crop[k, 1]), size, size)
        if flip[k]:
            im = TF.hflip(im)
        out.append(TF.normalize(TF.to_tensor(im), [.5] * 3, [.5] * 3))
    assert torch.equal(torch.stack(out), ref)


def test_dataset_device_transform_yields_the_resized_uint8_image(tmp_path):
    """AgeDB(..., device_transform=True).__getitem__ stops after the reference's Resize((img_size, img_size)) and returns
    the uint8 HWC image (what dirb200_augment_batch takes); label / weight as in the reference (datasets.py:27-36)."""
    import pandas as pd
    from PIL import Image
    from torchvision import transforms
    import datasets as D
    rng = np.random.RandomState(1)
    paths = []
    for i in range(3):
        arr = rng.randint(0, 256, size=(50 + 7 * i, 40 + 5 * i, 3)).astype(np.uint8)
        Image.fromarray(arr).save(tmp_path / f"img{i}.png")
        paths.append(f"img{i}.png")
    df = pd.DataFrame({"path": paths, "age": [31.0, 44.0, 74.0], "split": ["train"] * 3})
    ds = D.AgeDB(df=df, data_dir=str(tmp_path), img_size=32, split="train", device_transform=True)
    img, label, weight = ds[1]
    assert img.dtype == torch.uint8 and tuple(img.shape) == (32, 32, 3)
    want = np.asarray(transforms.Resize((32, 32))(Image.open(tmp_path / "img1.png").convert("RGB")))
    assert np.array_equal(img.numpy(), want)
    assert label.dtype == np.float32 and float(label[0]) == 44.0 and float(weight[0]) == 1.0
    # the host path is untouched: a normalised float tensor
    ds_host = D.AgeDB(df=df, data_dir=str(tmp_path), img_size=32, split="val")
    x, _, _ = ds_host[1]
    assert x.dtype == torch.float32 and tuple(x.shape) == (3, 32, 32) and float(x.min()) >= -1.0 and float(x.max()) <= 1.0
