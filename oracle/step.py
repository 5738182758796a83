"""Restatement of the reference training step.  TEST INFRASTRUCTURE ONLY.

``format_output``  <- deeprank_gnn/NeuralNet.py:616-631
``make_loss``      <- deeprank_gnn/NeuralNet.py:239-263
``train_step``     <- deeprank_gnn/NeuralNet.py:490-503 (body of the per-batch loop of ``_epoch``)
``eval_step``      <- deeprank_gnn/NeuralNet.py:432-446
"""
import torch
import torch.nn as nn


def format_output(pred, target=None, task='reg', classes=(0, 1), transform_sigmoid=False):
    if task == 'class':
        if target is not None:
            c2i = {c: i for i, c in enumerate(classes)}
            target = torch.tensor([c2i[int(x)] for x in target])
    elif transform_sigmoid is True:
        pred = torch.sigmoid(pred.reshape(-1))
    else:
        pred = pred.reshape(-1)
    return pred, target


def make_loss(task='reg', weights=None):
    if task == 'reg':
        return nn.MSELoss()
    return nn.CrossEntropyLoss(weight=weights, reduction='mean')


def train_step(model, optimizer, loss_fn, batch, task='reg', classes=(0, 1), transform_sigmoid=False):
    """zero_grad -> forward -> format_output -> loss -> backward -> Adam step."""
    y = batch.y
    optimizer.zero_grad()
    pred = model(batch)
    pred, y = format_output(pred, y, task, classes, transform_sigmoid)
    loss = loss_fn(pred, y)
    loss.backward()
    optimizer.step()
    return loss.detach(), pred.detach()


def eval_step(model, loss_fn, batch, task='reg', classes=(0, 1), transform_sigmoid=False):
    y = batch.y
    pred = model(batch)          # NB the reference does not use no_grad() here (NeuralNet.py:424-435)
    pred, y = format_output(pred, y, task, classes, transform_sigmoid)
    loss = None if y is None else loss_fn(pred, y).detach()
    return loss, pred.detach()
