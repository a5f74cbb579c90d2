import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from passion_b200.engine import Trainer
from passion_b200.models import rfnet
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x, target, mask, _ = synth.make_batch(4, 32, seed=77, labels="U", mask_ids=[10, 7, 12, 5])
sd = synth.make_state_dict(1037)
m = rfnet.Model(4).to(dev); m.load_state_dict(sd); m.compute_dtype = torch.float32
tr = Trainer(m, use_graph=False)
sl = slice(rank * 2, rank * 2 + 2)
loss, parts = tr.forward_loss(x[sl].to(dev), target[sl].to(dev), mask[sl].to(dev))
tr.optimizer.zero_grad(set_to_none=True)
tr.reducer.prepare()
p = m.flair_encoder.e1_c1.conv.weight
b, i = tr.reducer._slot[id(p)]
v = b["views"][i]
print(rank, "before backward: grad is view", p.grad is v, p.grad.data_ptr() == v.data_ptr(), flush=True)
loss.backward()
torch.cuda.synchronize()
print(rank, "after backward: grad ptr==view ptr", p.grad.data_ptr() == v.data_ptr(), "|grad|", float(p.grad.norm()), "|view|", float(v.norm()),
      "pending", [bb["pending"] for bb in tr.reducer.buckets], "work set", [bb["work"] is not None for bb in tr.reducer.buckets], flush=True)
tr.reducer.finish()
torch.cuda.synchronize()
print(rank, "after finish: |grad|", float(p.grad.norm()), "|view|", float(v.norm()), flush=True)
dist.barrier(); os._exit(0)
