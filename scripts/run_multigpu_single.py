"""One cosmology over N GPUs under torchrun: k partition -> NCCL all-gather of S -> q partition -> all-reduce of C_l.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
       scripts/run_multigpu_single.py [fixture]
Checks the result against the golden C_l of the fixture and prints device timings."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from classpp_public_b200 import modules as M, multigpu

name = sys.argv[1] if len(sys.argv) > 1 else "planck18"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
inp = M.Inputs.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
a = inp.arrays

class NL:
    nl_corr_density_m = a.get("nl.nl_corr_density_m")

nl = NL if NL.nl_corr_density_m is not None else None
prim = M.TabulatedPrimordial(a["pm.pk_at_transfer_k"])
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    cl, ct = multigpu.compute_cl_distributed(inp, prim, nl, rank, world, local)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
ref = a["ref.cl"].reshape(-1, ct); mine = cl.reshape(-1, ct)
err = {n: float(np.max(np.abs(mine[:, c] / ref[:, c] - 1))) for n, c in (("tt", 0), ("ee", 1), ("pp", 4)) if c < ct}
if rank == 0:
    print(json.dumps({"fixture": name, "world": world, "wall_s": dt, "max_rel_err_vs_reference": err}))
    assert max(err.values()) < 1e-4
dist.destroy_process_group()
