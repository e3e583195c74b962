"""jittor.nn slice (see package docstring)."""
import torch


class Module:
    def __init__(self, *a, **k):
        pass

    def is_training(self):
        return False

    def __call__(self, *a, **k):
        return self.execute(*a, **k)


class Sequential(Module):
    pass


class Linear(Module):
    pass


class ReLU(Module):
    pass


class ModuleList(list):
    pass


def softmax(x, dim=None):
    from . import Var
    return torch.softmax(x, dim=dim).as_subclass(Var)
