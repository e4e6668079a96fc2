"""Step loop mirror of /root/reference/unet3d/train/training_utils.py:20-112 (epoch_training / batch_loss /
_batch_loss), kept call-compatible so the reference's ``run_training`` can drive it, minus the per-step
``torch.cuda.empty_cache()`` (training_utils.py:46-47) and with non-blocking pinned H2D copies
(training_utils.py:89-91 copies synchronously from pageable memory).
"""
from __future__ import annotations

import time

import torch


class AverageMeter(object):
    """training_utils.py:156-178."""

    def __init__(self, name, fmt=":f"):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return ("{name} {val" + self.fmt + "} ({avg" + self.fmt + "})").format(**self.__dict__)


def _to_device(t: torch.Tensor) -> torch.Tensor:
    if t.is_cuda:
        return t
    return t.cuda(non_blocking=t.is_pinned())


def _batch_loss(model, images, target, criterion, inferer=None):
    """training_utils.py:101-112."""
    if inferer is not None:
        output = inferer(images, model).to(images.device)
    else:
        output = model(images)
    batch_size = images.size(0)
    loss = criterion(output, target)
    return loss, batch_size


def batch_loss(model, images, target, criterion, n_gpus=0, use_amp=None, inferer=None):
    """training_utils.py:88-98.  ``use_amp`` selects nothing here: precision is a model property."""
    if n_gpus is not None:
        images = _to_device(images)
        target = _to_device(target)
    return _batch_loss(model, images, target, criterion, inferer=inferer)


def epoch_training(train_loader, model, criterion, optimizer, epoch, n_gpus=None, print_frequency=1,
                   print_gpu_memory=False, scaler=None, samples_per_epoch=None, iteration=1, grad_sync=None):
    """training_utils.py:20-85.  ``grad_sync`` (optional callable) runs between backward and optimizer.step:
    the data-parallel gradient all-reduce of ``parallel.GradAllReduce``."""
    batch_time = AverageMeter("Time", ":6.3f")
    data_time = AverageMeter("Data", ":6.3f")
    losses = AverageMeter("Loss", ":.4e")
    model.train()
    end = time.time()
    for i, item in enumerate(train_loader):
        images, target = item["image"], item["label"]
        data_time.update(time.time() - end)
        optimizer.zero_grad()
        loss, batch_size = batch_loss(model, images, target, criterion, n_gpus=n_gpus, use_amp=scaler is not None)
        loss.backward()
        if grad_sync is not None:
            grad_sync()
        optimizer.step()
        losses.update(loss.item(), batch_size)   # the only host sync of the step, after all work is queued
        del loss
        batch_time.update(time.time() - end)
        end = time.time()
        if print_frequency and i % print_frequency == 0:
            print("Epoch: [{}][{}/{}]\t{}\t{}\t{}".format(epoch, i + 1, len(train_loader), batch_time, data_time, losses))
        if samples_per_epoch and (i + 1) * batch_size >= samples_per_epoch:
            break
    return losses.avg


def epoch_validation(val_loader, model, criterion, n_gpus, print_freq=1, use_amp=False, inferer=None):
    """training_utils.py:115-147."""
    losses = AverageMeter("Loss", ":.4e")
    model.eval()
    with torch.no_grad():
        for i, item in enumerate(val_loader):
            loss, batch_size = batch_loss(model, item["image"], item["label"], criterion, n_gpus=n_gpus, inferer=inferer)
            losses.update(loss.item(), batch_size)
    return losses.avg
