"""Drop-in for the reference's src/utils.py: to_var, to_cuda, get_data."""
import os

import torch

from gm_b200.gan_api import to_cuda, to_var  # noqa: F401  (src/utils.py:6-14)


def synthetic_binary_mnist(n, seed=3435, p=0.1307):
    """i.i.d. Bernoulli(p) 28x28 {0,1} images (p = MNIST mean intensity; mirrors the
    torch.bernoulli binarisation of src/utils.py:31) with random labels."""
    g = torch.Generator().manual_seed(seed)
    imgs = (torch.rand(n, 1, 28, 28, generator=g) < p).float()
    labels = torch.randint(0, 10, (n,), generator=g)
    return imgs, labels


def get_data(BATCH_SIZE=100, root="./data/"):
    """ Load data for binarised MNIST (src/utils.py:16-53): seeded Bernoulli binarisation,
    50k/10k/10k split, three shuffling DataLoaders.  MNIST is read from `root` when it is
    already on disk (no download: there is no network); otherwise a synthetic stand-in of
    the same shape and density is used and a note is printed. """
    torch.manual_seed(3435)
    train_img = test_img = None
    try:
        import torchvision.datasets as datasets
        import torchvision.transforms as transforms
        if os.path.isdir(os.path.join(root, "MNIST")):
            tr = datasets.MNIST(root=root, train=True, transform=transforms.ToTensor(), download=False)
            te = datasets.MNIST(root=root, train=False, transform=transforms.ToTensor())
            train_img = torch.stack([torch.bernoulli(d[0]) for d in tr])
            train_label = torch.LongTensor([d[1] for d in tr])
            test_img = torch.stack([torch.bernoulli(d[0]) for d in te])
            test_label = torch.LongTensor([d[1] for d in te])
    except Exception:
        train_img = None
    if train_img is None:
        print("get_data: MNIST not found under %s (no network) -> synthetic Bernoulli(0.1307) 28x28 images" % root)
        train_img, train_label = synthetic_binary_mnist(60000, seed=3435)
        test_img, test_label = synthetic_binary_mnist(10000, seed=3436)
    val_img, val_label = train_img[-10000:].clone(), train_label[-10000:].clone()
    train_img, train_label = train_img[:-10000], train_label[:-10000]
    train = torch.utils.data.TensorDataset(train_img, train_label)
    val = torch.utils.data.TensorDataset(val_img, val_label)
    test = torch.utils.data.TensorDataset(test_img, test_label)
    train_iter = torch.utils.data.DataLoader(train, batch_size=BATCH_SIZE, shuffle=True)
    val_iter = torch.utils.data.DataLoader(val, batch_size=BATCH_SIZE, shuffle=True)
    test_iter = torch.utils.data.DataLoader(test, batch_size=BATCH_SIZE, shuffle=True)
    return train_iter, val_iter, test_iter
