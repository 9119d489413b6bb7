#!/usr/bin/env python
"""Runs the UNMODIFIED reference NSGANTrainer.train (byte-compiled into oracle/_ref by
oracle/make_ref.py) on synthetic data and prints one JSON line with its throughput.
TEST / MEASUREMENT INFRASTRUCTURE, not product code: only bench.py executes this.

    python oracle/ref_runner.py --batch 64 --steps 782 --warmup 10 [--pool 50000] [--device cpu|cuda]

What runs is the reference's own public API on its own code path (src/ns_gan.py:80-170):
`NSGAN(784,400,20)`, `NSGANTrainer(model, loader, loader, loader, viz=False)`,
`trainer.train(num_epochs=1, G_lr=2e-4, D_lr=2e-4, D_steps=1)` over a shuffling
`DataLoader(TensorDataset(images, labels), batch_size=B)` — i.e. including the per-step
`next(iter(DataLoader))` fetch of src/ns_gan.py:222-226.  The harness only (a) stubs matplotlib /
IPython / torchvision imports that are not installed and never used by train(), (b) hides the GPU
for the CPU arm (`to_cuda`, src/utils.py:10-14, would otherwise move everything to cuda:0) and
(c) makes one "epoch" exactly `steps` train steps through the loader's reported length
(`epoch_steps = ceil(len(train_iter) / D_steps)`, src/ns_gan.py:114).  When oracle/_ref is absent
the runner falls back to oracle/torch_port.py and says so (`kind: "port"`).
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--pool", type=int, default=0, help="dataset size N (default 4 x batch, at least 4096)")
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--budget", type=float, default=150.0, help="seconds the timed steps may take; the batch is halved until they fit")
    ap.add_argument("--prefetched", action="store_true", help="compute-only: batches from a list, no DataLoader fetch per step")
    args = ap.parse_args()
    if args.device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""          # before torch initialises CUDA
    import torch
    if args.threads > 0:
        torch.set_num_threads(args.threads)
    threads = torch.get_num_threads()
    ref_dir = os.path.join(HERE, "_ref")
    have_ref = os.path.exists(os.path.join(ref_dir, "ns_gan.pyc"))

    def images(n):
        g = torch.Generator().manual_seed(3435)           # the reference's data seed, src/utils.py:18
        return (torch.rand(n, 1, 28, 28, generator=g) < 0.1307).float()

    if have_ref:
        for name in ["matplotlib", "matplotlib.pyplot", "IPython", "IPython.display", "torchvision", "torchvision.datasets",
                     "torchvision.transforms", "torchvision.utils"]:
            try:
                __import__(name)
            except Exception:     # noqa: BLE001 (not installed, or broken in this image): never used by train()
                m = types.ModuleType(name)
                m.display = lambda *a, **k: None
                sys.modules[name] = m
        sys.path.insert(0, ref_dir)
        import ns_gan as ref                               # oracle/_ref/ns_gan.pyc — the reference's own module

        class StepLoader(torch.utils.data.DataLoader):
            steps = 1

            def __len__(self):                             # epoch_steps = ceil(len(train_iter) / D_steps), src/ns_gan.py:114
                return self.steps

        class ListLoader(list):                            # compute-only variant: next(iter(list)) is its first batch
            pass

        def make(batch):
            n = args.pool or max(4 * batch, 4096)
            x = images(n)
            ds = torch.utils.data.TensorDataset(x, torch.zeros(n, dtype=torch.long))
            torch.manual_seed(1234)
            model = ref.NSGAN(784, 400, 20)
            if args.prefetched:
                loader = ListLoader([(x[:batch], torch.zeros(batch, dtype=torch.long))])
            else:
                loader = StepLoader(ds, batch_size=batch, shuffle=True)
            return ref.NSGANTrainer(model, loader, loader, loader, viz=False), loader

        def run(trainer, loader, steps):
            if args.prefetched:
                loader[:] = [loader[0]] * steps            # len(list) = steps
            else:
                loader.steps = steps
            sink = io.StringIO()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
                trainer.train(num_epochs=1, G_lr=2e-4, D_lr=2e-4, D_steps=1)
            if args.device == "cuda":
                torch.cuda.synchronize()
            return time.perf_counter() - t0
        kind = "reference"
    else:
        sys.path.insert(0, os.path.dirname(HERE))
        from oracle import torch_port as TP

        def make(batch):
            return batch, None

        def run(batch, _loader, steps):
            _, dt, _ = TP.time_cpu_steps(batch, steps=steps, warmup=0, threads=threads, with_loader=not args.prefetched)
            return dt
        kind = "port"

    batch = args.batch
    while True:
        trainer, loader = make(batch)
        t1 = run(trainer, loader, 1)                       # first step: allocations, thread pool start
        t1 = run(trainer, loader, 1)
        if t1 * args.steps <= args.budget or batch <= 64:
            break
        batch //= 2
    if args.warmup > 0:
        run(trainer, loader, args.warmup)
    dt = run(trainer, loader, args.steps)
    dev_name = torch.cuda.get_device_name(0) if args.device == "cuda" else None
    print(json.dumps({"kind": kind, "device": args.device, "device_name": dev_name, "batch": batch, "requested_batch": args.batch,
                      "steps": args.steps, "warmup": args.warmup, "seconds": round(dt, 4), "images_per_s": round(batch * args.steps / dt, 1),
                      "ms_per_step": round(dt / args.steps * 1e3, 4), "threads": threads, "host_cores": os.cpu_count(),
                      "pool": args.pool or max(4 * batch, 4096), "with_loader": not args.prefetched, "torch": torch.__version__}))


if __name__ == "__main__":
    main()
