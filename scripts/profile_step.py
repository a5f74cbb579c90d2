"""One eager RFNet+PASSION training step (B=2, 4x80^3, bf16) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` (see profiles/README.md).  Two un-profiled warm-up steps come first."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                   # noqa: E402  (synthetic batches, modal weights)
from passion_b200.engine import Trainer                         # noqa: E402
from passion_b200.models import rfnet                           # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1037)
    model = rfnet.Model(num_cls=4).to(dev)
    model.compute_dtype = torch.bfloat16
    trainer = Trainer(model, lr=2e-4, weight_decay=1e-4, temp=4.0, mask_type="idt", use_passion=True,
                      modal_weight=bench.modal_weight(), use_graph=False)
    batch = tuple(t.to(dev) for t in bench.synth_host_batches(0, 1, 2, 80)[0])
    for _ in range(2):
        trainer.step(*batch)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    trainer.step(*batch)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
