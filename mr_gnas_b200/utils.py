"""Host-side helpers the scripts import from utils/utils.py that touch the hot path."""
import torch


def weights_init(m):
    """reference: utils/utils.py:121-131 -- xavier-normal on every nn.Linear (zero bias); BatchNorm1d is
    left at its default init; Conv2d/BatchNorm2d branches serve the ConvE scorer."""
    if isinstance(m, torch.nn.Linear):
        torch.nn.init.xavier_normal_(m.weight)
        if m.bias is not None:
            torch.nn.init.constant_(m.bias, 0)
    elif isinstance(m, torch.nn.Conv2d):
        torch.nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
    elif isinstance(m, torch.nn.BatchNorm2d):
        torch.nn.init.constant_(m.weight, 1)
        if m.bias is not None:
            torch.nn.init.constant_(m.bias, 0)


def count_parameters_in_MB(model):
    """reference: utils/utils.py:36-37"""
    return sum(v.numel() for name, v in model.named_parameters() if "auxiliary" not in name) / 1e6
