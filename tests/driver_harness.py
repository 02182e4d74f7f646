"""Runs the reference's OWN training / rollout drivers on a model (TEST INFRASTRUCTURE).

`EGNO/main_simulation_simple_no.py:run_epoch` (+ `rollout_fn`, `prepare_inputs`) and `SEGNO/train_nbody.py:run_epoch`
(+ `rollout_fn`) are imported unmodified through oracle/ref_loader.py and handed either the reference's module (CPU) or
the CUDA drop-in module; everything around the model — dataset classes, featurisation, loss, Adam, Pearson horizon,
numpy energies — is the reference's code in both cases, which is the driver-level drop-in proof of SURVEY.md §8(c).
"""
from __future__ import annotations

import contextlib
import io
from argparse import Namespace

import numpy as np
import torch
from torch import nn
from torch.utils.data import DataLoader


def egno_driver_run(R, model, device, data_dir, n_balls, T, batch, epochs, traj_len, lr=5e-4, dataset="charged"):
    """-> dict(train=[avg loss per epoch], valid=avg loss, test_losses=[per-frame MSE], preds, targets, energies)."""
    args = Namespace(device=device, n_balls=n_balls, num_inputs=1, num_timesteps=T, traj_len=traj_len, batch_size=batch)
    mk = lambda part, tl: R.EgnoDataset(partition=part, data_dir=data_dir, dataset=dataset, dataset_name="nbody_small",
                                        n_balls=n_balls, num_timesteps=T, num_inputs=1, traj_len=tl)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        train, valid, test = mk("train", 1), mk("val", 1), mk("test", traj_len)
        ld = lambda ds: DataLoader(ds, batch_size=batch, shuffle=False, drop_last=True)
        opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=1e-8)
        crit = nn.MSELoss(reduction="none")
        run = R.egno_main.run_epoch
        out["train"] = [float(run(model, opt, crit, ep, ld(train), args, backprop=True, num_timesteps=T)) for ep in range(epochs)]
        with torch.no_grad():
            out["valid"] = float(run(model, opt, crit, epochs, ld(valid), args, backprop=False, num_timesteps=T))
        losses, traj = run(model, opt, crit, epochs, ld(test), args, backprop=False, rollout=True, num_timesteps=T)
    out["test_losses"] = [float(v) for v in losses]
    out["preds"] = traj["preds"].detach().cpu()
    out["targets"] = traj["targets"].detach().cpu()
    out["energies"] = traj["energy_conservation"].detach().cpu()
    out["test_loss"] = float(traj["test_loss"])
    return out


def segno_driver_run(R, model, device, data_dir, n_balls, T, batch, epochs, traj_len, lr=5e-4, dataset="gravity"):
    args = Namespace(device=device, n_balls=n_balls, num_inputs=1, num_timesteps=T, traj_len=traj_len, batch_size=batch,
                     varDT=False)
    mk = lambda part: R.SegnoDataset(root=data_dir, partition=part, dataset=dataset, dataset_size="small", n_balls=n_balls)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        train, valid, test = mk("train"), mk("val"), mk("test")
        ld = lambda ds: DataLoader(ds, batch_size=batch, shuffle=False, drop_last=True)
        opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=1e-12)
        crit = (nn.MSELoss(), nn.MSELoss(reduction="none"))                      # main.py:115
        run = R.segno_train.run_epoch
        out["train"] = [float(run(model, opt, crit, ep, ld(train), args, backprop=True, num_timesteps=T)) for ep in range(epochs)]
        with torch.no_grad():
            out["valid"] = float(run(model, opt, crit, epochs, ld(valid), args, backprop=False, num_timesteps=T))
        loss, traj = run(model, opt, crit, epochs, ld(test), args, backprop=False, rollout=True, num_timesteps=T)
    out["test_loss"] = float(loss)
    out["test_losses"] = [float(v) for v in np.mean(np.asarray(traj["losses"]), axis=0)]
    out["preds"] = traj["preds"].detach().cpu()
    out["targets"] = traj["targets"].detach().cpu()
    out["energies"] = traj["energies"].detach().cpu()
    return out
