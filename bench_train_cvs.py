#!/usr/bin/env python
"""SLODE train-epoch time on the CVS configuration (BASELINE.json configs[0] / the second half of its metric).

One epoch = the reference's: 7 mini-batches of 128 (two SVI steps each) + the four evaluation passes
(val-post, val-prior, train-post, train-prior; training_cvs.py:256-315) on 810 / 90 synthetic CVS series, default
config (midpoint, odeint_adjoint semantics, seed 12).  Two arms on the same data and weights:

  b200       structured_latent_odes_b200.training_cvs over the fused Decoder (this repo, cuda:0)
  reference  the same restated training step over the CPU port of the reference's Decoder / OdeModel / torchdiffeq
             (oracle/), all host threads -- Pyro itself is not installable, see training_cvs.py's header

Prints one JSON line per arm.  A scaled variant (--scale K: K x 810 series, batch K x 128) shows the path with a
batch the GPU can fill.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from structured_latent_odes_b200 import training_cvs as tc  # noqa: E402


def run(arm, cfg, data, epochs, warm):
    dev = "cuda" if arm == "b200" else "cpu"
    torch.manual_seed(cfg.seed)
    times = torch.arange(0.0, cfg.seq_len, 1.0, device=dev)
    if arm == "b200":
        model = tc.MechanisticModel(cfg, dev, times).to(dev)
    else:
        from oracle import slode_port
        torch.set_num_threads(os.cpu_count() or 1)
        model = tc.MechanisticModel(cfg, dev, times, decoder_cls=slode_port.Decoder)
    d = {k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in data.items()}
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate, betas=(0.9, 0.999))
    out = {}
    for evaluate in (True, False):
        ts = []
        for e in range(warm + epochs):
            if dev == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            losses, _ = tc.train_epoch(model, opt, d, cfg, generator=torch.Generator().manual_seed(e), evaluate=evaluate)
            if dev == "cuda":
                torch.cuda.synchronize()
            if e >= warm:
                ts.append(time.perf_counter() - t0)
        out["s_per_epoch_with_eval" if evaluate else "s_per_epoch_train_steps_only"] = min(ts)
    out["last_losses_per_sample"] = [float(x) for x in losses]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    cfg = tc.cvs_config(data_size=1000 * a.scale, mini_batch_size=128 * a.scale)
    data = tc.make_cvs_dataset(cfg, "cuda", generator=torch.Generator(device="cuda").manual_seed(cfg.seed))
    n_train = data["train"]["observations"].shape[0]
    base = {"metric": "slode_cvs_train_epoch_time", "unit": "s", "higher_is_better": False,
            "config": {"workload": "training_cvs.py default config (midpoint, odeint_adjoint), synthetic CVS data",
                       "n_train": n_train, "n_val": data["val"]["observations"].shape[0], "mini_batch": cfg.mini_batch_size,
                       "epoch": "train steps + 4 evaluation passes"}}
    for arm in ["b200"] + ([] if a.no_reference else ["reference"]):
        r = run(arm, cfg, data, a.epochs, a.warmup)
        line = dict(base, impl=arm, value=r["s_per_epoch_with_eval"], **r)
        if arm == "reference":
            line["cores"] = os.cpu_count()
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
