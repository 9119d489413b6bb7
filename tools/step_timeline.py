"""In-stream per-kernel timing of one workload's train step (gm_prof_enable level 2: CUDA events around EVERY
launch on its stream): what each kernel costs inside the real back-to-back stream, next to the event-timed
step.  usage: python tools/step_timeline.py [ns|wgp|vae] [steps] > profiles/<name>.md"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "generative-models_b200"))
import torch  # noqa: E402
import gm_b200  # noqa: E402
import bench  # noqa: E402

from gm_b200 import parallel as par  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ns"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world, local = par.init_from_env("nccl")       # torchrun: one rank per GPU, rank 0 reports
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
B = bench.DEFAULT_BATCH[name]
N = max(4 * 65536, 2 * B)
bits = bench.make_pool(N, dev, rank)
comm = par.make_peer_comm(330000) if (world > 1 and name != "vae") else None
wl = bench.Workload(name, B, bits, N, rank, world, comm)
for _ in range(5):
    wl.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
t0 = time.perf_counter()
for _ in range(steps):
    wl.step()
host_ms = (time.perf_counter() - t0) * 1e3 / steps
e1.record()
torch.cuda.synchronize()
free_ms = e0.elapsed_time(e1) / steps
gm_b200.prof_enable(2)
for _ in range(steps):
    wl.step()
rep = gm_b200.prof_report()
gm_b200.prof_enable(0)
tot = sum(r[2] for r in rep)
if world > 1:
    torch.distributed.barrier()
    if rank != 0:
        comm and comm.close()
        torch.distributed.destroy_process_group()
        sys.exit(0)
    print("### %d ranks (this is rank 0; the exchange kernel's time includes waiting for the slowest peer)" % world)
print("## %s, B=%d: free-running step %.4f ms (host enqueue %.4f ms/step); sum of in-stream kernel times %.4f ms/step"
      % (name, B, free_ms, host_ms, tot / steps))
print("\n| kernel (launch order) | launches/step | us/step | share |\n|---|---|---|---|")
for n, c, ms in rep:
    print("| `%s` | %.1f | %.1f | %.1f %% |" % (n, c / steps, ms * 1e3 / steps, 100 * ms / tot))
if world > 1:
    comm and comm.close()
    torch.distributed.destroy_process_group()
