"""Multi-GPU equivalence on hardware (SURVEY.md §8e "Equivalence test"): N ranks x B = 2 under NCCL must give the same
loss and gradients as ONE process that loops over the same chunks, applies the global rp_mask and sums the gradients.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ddp_equivalence.py [--size 32]

fp32 check mode, eager steps.  Checked on rank 0 (every rank computes the chunk loop for its own comparison of rp_mask):
  * the all-reduced gradient buckets == sum over chunks of the chunk gradients (rel-L2 <= 1e-6: only the reduction order of the
    cross-rank sum differs; per-chunk kernels use float64 atomics whose order varies run to run, hence not bit-exact);
  * the step loss summed over ranks == the chunk loop's;
  * rp_mask is the GLOBAL-batch statistic (train.py:265-268): every rank sees the all-reduced rp_iter (printed next to the
    per-chunk values, with whether a chunk-local gate would have decided differently);
  * the prototype loss's class gate (criterions.py:157 `.all()` over the LOCAL batch, i.e. the DataParallel chunk) is evaluated
    per rank: sample 1 of chunk 0 lacks class 3, so chunk 0 drops that class while chunk 1 keeps it.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32)
    args = ap.parse_args()
    from oracle import synth                       # deterministic weights / batches (test infrastructure; this is a test)
    from passion_b200.engine import Trainer
    from passion_b200.models import rfnet
    from passion_b200.train_step import loss_mix

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, S = 2, args.size
    mask_ids = [10, 7, 12, 5, 14, 3, 9, 13][: B * world]
    x, target, mask, y = synth.make_batch(B * world, S, seed=77, labels="U", mask_ids=mask_ids)
    # class gate: remove class 3 from sample 1 (chunk 0) only
    y = target.argmax(1)
    y[1][y[1] == 3] = 0
    target = torch.from_numpy(np.ascontiguousarray(np.eye(4)[y.numpy()].transpose(0, 4, 1, 2, 3)))
    sd = synth.make_state_dict(1037)
    beta = torch.tensor([1.1, 0.9, 1.3, 0.7])
    mw = torch.tensor([219 / 90.0, 219 / 135.0, 219 / 184.0, 219 / 43.0])

    def fresh():
        m = rfnet.Model(4).to(dev)
        m.load_state_dict(sd)
        m.compute_dtype = torch.float32
        return m

    # ---- distributed step: this rank's chunk, NCCL exchanges
    model = fresh()
    tr = Trainer(model, lr=2e-4, weight_decay=1e-4, temp=4.0, mask_type="idt", use_passion=True, modal_weight=mw, imb_beta=beta,
                 use_graph=False)
    sl = slice(rank * B, (rank + 1) * B)
    loss, parts = tr.forward_loss(x[sl].to(dev), target[sl].to(dev), mask[sl].to(dev))
    tr.optimizer.zero_grad(set_to_none=True)
    tr.reducer.prepare()
    loss.backward()
    tr.reducer.finish()
    g_ddp = torch.cat([p.grad.flatten().double() for p in model.parameters()])
    loss_sum = loss.detach().clone()
    dist.all_reduce(loss_sum)
    rp_global = parts["rp_iter"].detach().cpu()

    # ---- single-process chunk loop with the same semantics
    ref = fresh()
    ref.is_training, ref.use_passion, ref.mask_type = True, True, "idt"
    outs_c, rp_c, present_c = [], [], []
    for c in range(world):
        sc = slice(c * B, (c + 1) * B)
        outs = ref(x[sc].to(dev), mask[sc].to(dev), target=target[sc].to(dev), temp=4.0)
        outs_c.append(outs)
        _, p = loss_mix(outs, target[sc].to(dev), mask[sc].to(dev), beta.to(dev), mw.to(dev))
        rp_c.append(p["rp_iter"].detach())
        present_c.append((target[sc].sum((2, 3, 4)) > 0).all(0))
    rp_sum = torch.stack(rp_c).sum(0)
    total = 0.0
    for c in range(world):
        sc = slice(c * B, (c + 1) * B)
        l, _ = loss_mix(outs_c[c], target[sc].to(dev), mask[sc].to(dev), beta.to(dev), mw.to(dev), rp_allreduce=lambda t: rp_sum)
        total = total + l
    total.backward()
    g_ref = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten().double() for p in ref.parameters()])
    torch.cuda.synchronize()

    r_g = float((g_ddp - g_ref).norm() / g_ref.norm())
    if rank == 0 and r_g > 1e-4:                       # diagnosis: which parameters disagree
        off, bad = 0, []
        for name, p_ in ref.named_parameters():
            k = p_.numel()
            a, b = g_ddp[off:off + k], g_ref[off:off + k]
            r = float((a - b).norm() / b.norm().clamp_min(1e-30))
            if r > 1e-4 and float(b.norm()) > 1e-9:
                bad.append((name, round(r, 4), round(float(a.norm()), 5), round(float(b.norm()), 5)))
            off += k
        print(f"  {len(bad)} parameters disagree; first: {bad[:10]}")
    r_l = abs(float(loss_sum) - float(total)) / abs(float(total))
    flips = [bool(((rp_c[c] > 0) != (rp_sum > 0)).any()) for c in range(world)]
    if rank == 0:
        print(f"ddp_equivalence: world {world}, B={B}/rank, 4x{S}^3, fp32 check mode")
        print(f"  gradient rel-L2 (NCCL-reduced buckets vs chunk loop): {r_g:.3e}")
        print(f"  loss: sum over ranks {float(loss_sum):.6f} vs chunk loop {float(total):.6f} (rel {r_l:.2e})")
        print(f"  rp_iter global {rp_sum.tolist()} == rank view {rp_global.tolist()}")
        print(f"  per-chunk rp_iter {[r.tolist() for r in rp_c]}; a chunk-local gate would differ from the global one: {flips}")
        print(f"  class present in all samples of chunk: {[p.tolist() for p in present_c]}")
    ok = r_g < 1e-6 and r_l < 1e-6 and torch.allclose(rp_global, rp_sum.cpu(), atol=1e-5, equal_nan=True)
    ok = ok and (not bool(present_c[0][3])) and bool(present_c[min(1, world - 1)][3] or world == 1)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("ddp_equivalence:", "PASS" if int(flag) == 0 else "FAIL")
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if int(flag) == 0 else 1)


if __name__ == "__main__":
    main()
