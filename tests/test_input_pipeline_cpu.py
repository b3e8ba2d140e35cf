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
