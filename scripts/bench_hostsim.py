"""Test harness: bench.py's control flow (argument handling, the e2e arm through ParticleSet.execute, the JSON line, under torchrun
the mode R reduction + the mode_d block, its fall-back to the collectives and its watchdog) on the HOST SIMULATION of the kernel
sources, with gloo instead of NCCL.  The torch.cuda entry points bench.py calls are replaced by no-ops here -- nothing in this file
is product or measurement code, and the numbers such a run prints mean nothing.
  PB_LIB=oracle/_build/hostsim/libparcels_b200_hostsim.so PB_HOSTSIM_TEST=1 python scripts/bench_hostsim.py --workload c2_small ...
  ... python -m torch.distributed.run --nproc-per-node 2 ... scripts/bench_hostsim.py --gpus 2 --workload c2_small ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

assert os.environ.get("PB_HOSTSIM_TEST") == "1", "this harness only drives the host simulation"
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.empty_cache = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self.clone()
_tensor = torch.tensor


def _cpu_tensor(*a, **k):
    if str(k.get("device", "")).startswith("cuda"):
        k["device"] = "cpu"
    return _tensor(*a, **k)


torch.tensor = _cpu_tensor
_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _init("gloo")
os.environ["LOCAL_RANK"] = "0"  # the simulation has one device
import bench  # noqa: E402

sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()
