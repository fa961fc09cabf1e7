"""Minimal stand-in for pytorch_lightning (absent in this image)."""
import torch


class LightningModule(torch.nn.Module):
    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")


def seed_everything(seed):
    torch.manual_seed(seed)


class Callback:
    pass


class Trainer:
    pass
