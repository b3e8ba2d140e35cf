"""GPU parity of the batched input pipeline against the reference's own per-sample torchvision chain
(agedb-dir/datasets.py:38-53: RandomCrop(224, padding=16) -> RandomHorizontalFlip -> ToTensor -> Normalize): bit-exact
for the same random draws, and the draws themselves reproduce torchvision's for the same seed."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _images(n, size, seed):
    rng = np.random.RandomState(seed)
    return rng.randint(0, 256, size=(n, size, size, 3)).astype(np.uint8)


@pytest.mark.parametrize("n,size", [(5, 32), (16, 224)], ids=["small", "img224"])
def test_train_transform_bit_exact_vs_torchvision(n, size):
    from PIL import Image
    from torchvision import transforms
    import datasets as D
    imgs = _images(n, size, 3)
    chain = transforms.Compose([transforms.RandomCrop(size, padding=16), transforms.RandomHorizontalFlip(),
                                transforms.ToTensor(), transforms.Normalize([.5, .5, .5], [.5, .5, .5])])
    torch.manual_seed(1234)
    ref = torch.stack([chain(Image.fromarray(im)) for im in imgs])          # the reference: one sample at a time
    torch.manual_seed(1234)
    got = D.gpu_transform_batch(torch.from_numpy(imgs).to(DEV), train=True)   # same generator, same draw order
    assert got.shape == (n, 3, size, size) and got.dtype == torch.float32
    assert torch.equal(got.cpu(), ref)
    # a run without flips / with extreme crop origins: padding on every side
    crop = torch.tensor([[0, 0], [32, 32], [0, 32], [16, 16], [5, 27]][:n] + [[16, 16]] * max(0, n - 5), dtype=torch.int32)
    flip = torch.tensor(([1, 0] * n)[:n], dtype=torch.uint8)
    got = D.gpu_transform_batch(torch.from_numpy(imgs).to(DEV), train=True, crop_yx=crop, flip=flip).cpu()
    import torchvision.transforms.functional as TF
    for k in range(n):
        im = TF.pad(Image.fromarray(imgs[k]), 16)
        im = TF.crop(im, int(crop[k, 0]), int(crop[k, 1]), size, size)
        if flip[k]:
            im = TF.hflip(im)
        want = TF.normalize(TF.to_tensor(im), [.5, .5, .5], [.5, .5, .5])
        assert torch.equal(got[k], want), k


def test_val_transform_bit_exact_vs_torchvision():
    from PIL import Image
    from torchvision import transforms
    import datasets as D
    imgs = _images(7, 64, 5)
    chain = transforms.Compose([transforms.ToTensor(), transforms.Normalize([.5, .5, .5], [.5, .5, .5])])
    ref = torch.stack([chain(Image.fromarray(im)) for im in imgs])
    got = D.gpu_transform_batch(torch.from_numpy(imgs).to(DEV), train=False)
    assert torch.equal(got.cpu(), ref)
