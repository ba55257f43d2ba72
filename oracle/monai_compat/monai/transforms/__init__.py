"""monai.transforms subset (0.7.0): AsDiscrete / Activations / Compose as used by
OARSegmentation/config.py:69-70."""
import torch


class AsDiscrete:
    def __init__(self, argmax=False, to_onehot=False, n_classes=None, threshold_values=False,
                 logit_thresh=0.5):
        self.argmax, self.to_onehot, self.n_classes = argmax, to_onehot, n_classes
        self.threshold_values, self.logit_thresh = threshold_values, logit_thresh

    def __call__(self, img):
        if self.argmax:
            img = torch.argmax(img, dim=0, keepdim=True)
        if self.to_onehot:
            idx = img.long()
            out = torch.zeros((self.n_classes,) + tuple(idx.shape[1:]), dtype=torch.float32, device=img.device)
            img = out.scatter_(0, idx, 1.0)
        if self.threshold_values:
            img = img >= self.logit_thresh
        return img.float()


class Activations:
    def __init__(self, sigmoid=False, softmax=False, other=None):
        self.sigmoid, self.softmax, self.other = sigmoid, softmax, other

    def __call__(self, img):
        if self.sigmoid:
            img = torch.sigmoid(img)
        if self.softmax:
            img = torch.softmax(img, dim=0)
        if self.other is not None:
            img = self.other(img)
        return img


class Compose:
    def __init__(self, transforms=None):
        self.transforms = list(transforms or [])

    def __call__(self, x):
        for t in self.transforms:
            x = t(x)
        return x
