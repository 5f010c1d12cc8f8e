#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_multi_gpu.py : sharded rdi / crdi / uniformised rdi / dynamics coupling == unsharded, bit for bit."""
import os, sys
import numpy as np, pandas as pd, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import recipes
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ["SCRIBE_B200_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import scribe_py_b200 as S
from scribe_py_b200 import distributed as D
E = recipes.rdi_synthetic(12, 700, 5)
frame = pd.DataFrame(E, index=pd.Index(["g%02d" % i for i in range(12)], name="GENE_ID"))
m = S.causal_model(frame)
r = m.rdi([1, 5])
c = m.crdi(L=1)
spl, vel, cgenes = recipes.gv_c()
ad = recipes.FakeAnnData({"spliced": spl.copy(), "velocity": vel.copy()}, cgenes)
S.causal_net_dynamics_coupling(ad)
cp = ad.uns["causal_net"]["RDI"].to_numpy(dtype=float)
u = m.rdi([1], uniformization=True)[1].to_numpy(dtype=float)
# unsharded on this rank alone
real_world = D.world
D.world = lambda: (0, 1)
m1 = S.causal_model(frame)
r1 = m1.rdi([1, 5]); c1 = m1.crdi(L=1)
ad1 = recipes.FakeAnnData({"spliced": spl.copy(), "velocity": vel.copy()}, cgenes)
S.causal_net_dynamics_coupling(ad1)
cp1 = ad1.uns["causal_net"]["RDI"].to_numpy(dtype=float)
u1 = m1.rdi([1], uniformization=True)[1].to_numpy(dtype=float)
D.world = real_world
# a failing rank raises everywhere instead of hanging the others in the all-gather
def failing(lo, hi, out_dev_ptr=None):
    if dist.get_rank() == 1:
        raise ValueError("delay leaves no samples")
    return None if out_dev_ptr is not None else np.zeros(hi - lo)
try:
    D.sharded_evaluate(64, failing, S._lib.default_handle()); raised = False
except ValueError:
    raised = dist.get_rank() == 1
except D.ShardError:
    raised = dist.get_rank() != 1
ok = raised and all(np.array_equal(r[d].to_numpy(), r1[d].to_numpy(), equal_nan=True) for d in (1, 5, "MAX")) and \
    np.array_equal(c.to_numpy(), c1.to_numpy(), equal_nan=True) and np.array_equal(cp, cp1) and np.array_equal(u, u1, equal_nan=True)
t = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if dist.get_rank() == 0:
    print("multi-gpu invariance (world=%d): %s" % (dist.get_world_size(), "OK bit-identical" if t.item() == 1.0 else "MISMATCH"))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
