"""Developer timing probe (not the contract bench): one RFNet+PASSION step at B=2, 80^3, eager."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth                                   # noqa: E402  (inputs only)
from passion_b200 import _lib                              # noqa: E402
from passion_b200.models import rfnet                      # noqa: E402
from passion_b200.train_step import loss_mix               # noqa: E402


def main():
    S = int(os.environ.get("S", 80))
    B = int(os.environ.get("B", 2))
    dev = "cuda"
    sd = synth.make_state_dict(1037)
    x, target, mask, _ = synth.make_batch(B, S, seed=1037, labels="U", mask_ids=[10, 14][:B])
    x, target, mask = x.to(dev), target.to(dev), mask.to(dev)
    beta = torch.ones(4, device=dev)
    mw = torch.tensor([2.4, 1.6, 1.2, 5.1], device=dev)
    for dt in (torch.bfloat16, torch.float32):
        model = rfnet.Model(4).to(dev)
        model.load_state_dict(sd)
        model.is_training, model.use_passion, model.mask_type, model.compute_dtype = True, True, "idt", dt
        opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=1e-4, amsgrad=True)
        times = []
        for it in range(4):
            torch.cuda.synchronize()
            t0 = time.time()
            l0 = _lib.launch_count()
            outs = model(x, mask, target=target, temp=4.0)
            loss, _ = loss_mix(outs, target, mask, beta, mw)
            torch.cuda.synchronize()
            t1 = time.time()
            opt.zero_grad()
            loss.backward()
            opt.step()
            torch.cuda.synchronize()
            t2 = time.time()
            times.append((t1 - t0, t2 - t1))
            print(f"{dt} it{it}: fwd {t1 - t0:.3f}s bwd+opt {t2 - t1:.3f}s loss {float(loss):.5f} "
                  f"launches {_lib.launch_count() - l0} mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        del model, opt
        torch.cuda.empty_cache()
    # per-op profile of one bf16 step
    from torch.profiler import ProfilerActivity, profile
    model = rfnet.Model(4).to(dev)
    model.load_state_dict(sd)
    model.is_training, model.use_passion, model.mask_type, model.compute_dtype = True, True, "idt", torch.bfloat16
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        outs = model(x, mask, target=target, temp=4.0)
        loss, _ = loss_mix(outs, target, mask, beta, mw)
        loss.backward()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))


if __name__ == "__main__":
    main()
