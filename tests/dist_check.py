"""torchrun check of the library-owned NCCL communicator: ShardedCTCLoss (lib comm) vs a c10d all-reduce."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import oracle
from end2end_b200 import CTCLoss
from end2end_b200.distributed import ShardedCTCLoss, LossComm
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 8
x, tg, ll, tl = oracle.make_inputs(B, 100, 29, 10, 40, 100 + rank)
x = x.cuda().requires_grad_(); tg, ll, tl = tg.cuda(), ll.cuda(), tl.cuda()
crit = ShardedCTCLoss(reduce=True, size_average=True, global_batch=B * world)
loss = crit(x, tg, ll, tl); loss.backward()
g1 = x.grad.clone(); x.grad = None
ref = CTCLoss(reduce=True, size_average=False)(x, tg, ll, tl)
ref.backward()
tot = ref.detach().clone(); dist.all_reduce(tot); tot = tot / (B * world)
g2 = x.grad / (B * world)
ok = abs(loss.item() - tot.item()) < 1e-5 * abs(tot.item()) + 1e-5 and torch.allclose(g1, g2, rtol=1e-5, atol=1e-6)
print("rank %d lib comm %s loss %.6f ref %.6f grads close %s -> %s" % (rank, LossComm.get() is not None, loss.item(), tot.item(), torch.allclose(g1, g2, rtol=1e-5, atol=1e-6), "OK" if ok else "MISMATCH"), flush=True)
dist.barrier()
dist.destroy_process_group()
