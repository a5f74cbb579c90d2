#!/usr/bin/env python
"""Training entry point with the reference's flags and loop shape (code/train.py:69-373), on the B200-native path.

    python train.py --use_passion --model rfnet --batch_size 2 --synthetic --num_epochs 1
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 train.py --use_passion --batch_size 2 ...

One process per GPU (torchrun) replaces torch.nn.DataParallel (train.py:90).  Per iteration the work is
passion_b200.engine.Trainer.step (train.py:198-289); per epoch the LR schedule (lr_scheduler.py:15-17), the
relative-preference update of imb_beta (train.py:325-335) and the reference-format checkpoint (train.py:358-364).
Data: `<datasetPath>/vol/*_vol.npy` + `seg/*_seg.npy` with a random 80^3 crop on the host, or --synthetic; with
--device_aug the cases stay resident in HBM and the reference's whole transform chain (options.py:50: crop, rotation,
intensity change, flips) plus the label encoding run as one kernel per batch, bit-exact against the numpy / scipy
original (passion_b200/data.py).
"""
import csv
import logging
import os
import random
import time

import numpy as np
import torch
import torch.distributed as dist

from options import args_parser
from passion_b200 import ops
from passion_b200.data import AugmentSampler, DeviceAugment, ResidentCases
from passion_b200.engine import DevicePrefetcher, Trainer
from passion_b200.models import build_model
from passion_b200.train_step import poly_lr, preference_update

MASK_ARRAY = np.array([[False, False, False, True], [False, True, False, False], [False, False, True, False],
                       [True, False, False, False], [False, True, False, True], [False, True, True, False],
                       [True, False, True, False], [False, False, True, True], [True, False, False, True],
                       [True, True, False, False], [True, True, True, False], [True, False, True, True],
                       [True, True, False, True], [False, True, True, True], [True, True, True, True]])   # train.py:42-45


def read_split(path):
    with open(path) as f:
        return list(csv.DictReader(f))


class Source:
    """Yields (x f32 [B,4,80,80,80], one-hot target f64 [B,4,80,80,80], mask bool [B,4]) like the reference loader."""

    def __init__(self, args, rows, rank, world):
        self.args, self.rows, self.rank, self.world = args, rows, rank, world
        self.rs = np.random.RandomState(args.seed + rank)
        per_step = args.batch_size * world
        self.iters = args.iters_per_epoch or max(1, len(rows) // per_step)
        self._perm_epoch, self._perm = -1, None

    def row_index(self, it, b):
        """Case index of sample b of this rank at global iteration `it`: the rows are re-shuffled every epoch (the
        reference's DataLoader(shuffle=True), train.py:122-128) with a generator seeded by (seed, epoch) — the same
        permutation on every rank, each rank takes its own slice of it."""
        per_step = self.args.batch_size * self.world
        epoch, i = divmod(it, self.iters)
        if epoch != self._perm_epoch:
            self._perm = np.random.RandomState((self.args.seed * 1000003 + epoch) % (2 ** 32)).permutation(len(self.rows))
            self._perm_epoch = epoch
        return int(self._perm[(i * per_step + self.rank * self.args.batch_size + b) % len(self.rows)])

    def _mask_id(self, row):
        if self.args.mask_type == 'idt':
            return int(row['mask_id'])
        if self.args.mask_type == 'idt_drop':
            return int(self.rs.choice(eval(row['pos_mask_ids']), 1)[0])
        return int(self.rs.choice(15, 1)[0])

    def batch(self, it):
        B, S = self.args.batch_size, self.args.crop_size
        xs, ys, ms = [], [], []
        for b in range(B):
            row = self.rows[self.row_index(it, b)]
            if self.args.synthetic:
                x = self.rs.standard_normal((4, S, S, S)).astype(np.float32)
                y = self.rs.randint(0, 4, (S, S, S))
            else:
                vol = np.load(os.path.join(self.args.datasetPath, 'vol', row['data_name'] + '_vol.npy'))   # [H,W,Z,4]
                seg = np.load(os.path.join(self.args.datasetPath, 'seg', row['data_name'] + '_seg.npy'))
                o = [self.rs.randint(0, max(1, vol.shape[i] - S + 1)) for i in range(3)]
                x = np.ascontiguousarray(vol[o[0]:o[0] + S, o[1]:o[1] + S, o[2]:o[2] + S].transpose(3, 0, 1, 2)).astype(np.float32)
                y = seg[o[0]:o[0] + S, o[1]:o[1] + S, o[2]:o[2] + S].astype(np.int64)
            xs.append(x)
            ys.append(np.eye(4)[y].transpose(3, 0, 1, 2))                       # datasets_nii.py:150-153
            ms.append(MASK_ARRAY[self._mask_id(row)])
        return (torch.from_numpy(np.stack(xs)), torch.from_numpy(np.ascontiguousarray(np.stack(ys))),
                torch.from_numpy(np.stack(ms)))


class ResidentSource(Source):
    """--device_aug: yields DEVICE batches (x f32 [B,4,S,S,S], labels uint8 [B,S,S,S], mask bool [B,4]).  The cases live in
    HBM; per step the host draws the transform parameters in the reference's order and sends ~10 KB."""

    def __init__(self, args, rows, rank, world, dev):
        super().__init__(args, rows, rank, world)
        self.dev = dev
        S = args.crop_size
        self.cases = ResidentCases(dev)
        if args.synthetic:
            for i in range(4):
                self.cases.add(self.rs.standard_normal((S + 40, S + 40, S + 24, 4)).astype(np.float32),
                               self.rs.randint(0, 4, (S + 40, S + 40, S + 24)).astype(np.uint8))
        else:
            self.cases.add_files(args.datasetPath, [r['data_name'] for r in rows])
        logging.info('%d cases resident in HBM (%.1f GB)', len(self.cases), self.cases.nbytes() / 1e9)
        self.sampler = AugmentSampler((S, S, S), py_rng=random.Random(args.seed + rank), np_rng=np.random.RandomState(args.seed + rank))
        self.aug = DeviceAugment(dev, size=(S, S, S), batch=args.batch_size)

    def batch(self, it):
        B = self.args.batch_size
        ids, ps, ms = [], [], []
        for b in range(B):
            k = self.row_index(it, b)
            cid = k % len(self.cases) if self.args.synthetic else k
            ids.append(cid)
            ps.append(self.sampler.sample(tuple(self.cases.vols[cid].shape[:3])))
            ms.append(MASK_ARRAY[self._mask_id(self.rows[k])])
        x, labels, _ = self.aug(self.cases, ids, ps)
        return x, labels, torch.from_numpy(np.stack(ms)).to(self.dev)


def main():
    args = args_parser()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.model not in ('rfnet', 'mmformer'):
        raise SystemExit('--model rfnet | mmformer are implemented on the B200-native path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    os.makedirs(args.savepath, exist_ok=True)
    logging.basicConfig(level=logging.INFO if rank == 0 else logging.WARNING, format='%(asctime)s %(message)s',
                        handlers=[logging.StreamHandler(), logging.FileHandler(os.path.join(args.savepath, f'{args.mask_type}_training.txt'))])
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)

    csv_path = os.path.join(args.datarootPath, args.imbmrpath)
    if not os.path.exists(csv_path):
        csv_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tests', 'golden', os.path.basename(args.imbmrpath))
    rows = read_split(csv_path)
    src = ResidentSource(args, rows, rank, world, dev) if args.device_aug else Source(args, rows, rank, world)
    if not args.device_aug and not args.synthetic:
        logging.warning('--host_crop_only: the host loader applies the random crop only; RandomRotion / RandomIntensityChange / '
                        'RandomFlip of args.train_transforms are SKIPPED (use --device_aug for the reference chain)')
    modal_num = torch.tensor(np.sum([eval(r['mask']) for r in rows], 0), dtype=torch.float32)      # train.py:163-166
    logging.info('Training Imperfect Datasets with Mod.Flair-%d, Mod.T1c-%d, Mod.T1-%d, Mod.T2-%d', *modal_num.int().tolist())
    iter_per_epoch = src.iters
    modal_weight = iter_per_epoch / modal_num                                                       # train.py:171

    model = build_model(args.model, num_cls=4, crop=args.crop_size).to(dev)
    model.compute_dtype = torch.float32 if args.dtype == 'f32' else torch.bfloat16
    if args.resume is not None and args.use_pretrain:                                               # train.py:144-152
        sd = torch.load(args.resume, map_location=dev)['state_dict']
        sd = {k[len('module.'):] if k.startswith('module.') else k: v for k, v in sd.items()}
        model.load_state_dict({**model.state_dict(), **{k: v for k, v in sd.items() if k in model.state_dict()}})
        logging.info('load ok')
    trainer = Trainer(model, lr=args.lr, weight_decay=args.weight_decay, temp=args.temp, mask_type=args.mask_type,
                      use_passion=args.use_passion, modal_weight=modal_weight, use_graph=not args.no_graph)
    eta, eta_ext = 0.01, 1.5
    for epoch in range(args.num_epochs):
        lr = poly_lr(args.lr, epoch, args.num_epochs)
        trainer.set_lr(lr)
        trainer.warmup = epoch < args.region_fusion_start_epoch
        acc = torch.zeros(4, device=dev)
        t0 = time.time()
        if args.device_aug:
            feed = (src.batch(epoch * iter_per_epoch + i) for i in range(iter_per_epoch))
        else:
            feed = DevicePrefetcher((tuple(t.pin_memory() for t in src.batch(epoch * iter_per_epoch + i))
                                     for i in range(iter_per_epoch)), dev)        # H2D of batch i+1 overlaps step i
        for i, batch in enumerate(feed):
            loss, parts = trainer.step(*batch)
            if not args.device_aug:
                feed.release(batch)
            dist_m = parts['dist_m'].clone()
            if world > 1:
                dist.all_reduce(dist_m)
            acc += dist_m / (modal_num.to(dev) if args.mask_type == 'idt' else iter_per_epoch)      # train.py:299-308
            if rank == 0 and (i % 10 == 0 or i == iter_per_epoch - 1):
                vals = torch.stack([loss, parts['fuse'], parts['prm'], parts['sep'], parts['kl'], parts['proto']]).tolist()   # one D2H
                logging.info('Epoch %d/%d, Iter %d/%d, Loss %.4f, fuse_loss:%.4f, prm_loss:%.4f, sep_loss:%.4f, kl_loss:%.4f, proto_loss:%.4f',
                             epoch + 1, args.num_epochs, i + 1, iter_per_epoch, *vals)
        torch.cuda.synchronize()
        ops.check_tc_errors()                            # a tcgen05 pipeline time-out must not corrupt training silently
        logging.info('train time per epoch: %.2f s (%.2f samples/s)', time.time() - t0,
                     iter_per_epoch * args.batch_size * world / (time.time() - t0))
        if args.use_passion and epoch >= args.region_fusion_start_epoch:                            # train.py:325-335
            beta, eta, rp_epoch = preference_update(trainer.imb_beta.cpu(), acc.cpu(), eta, epoch, eta_ext)
            trainer.imb_beta.copy_(beta.to(dev))
            logging.info('rp_epoch:%s imb_beta:%s', [round(v, 4) for v in rp_epoch.tolist()], [round(v, 4) for v in beta.tolist()])
        if rank == 0:                                                                               # train.py:358-364
            torch.save({'epoch': epoch, 'state_dict': {'module.' + k: v for k, v in model.state_dict().items()},
                        'optim_dict': trainer.optimizer.state_dict()}, os.path.join(args.savepath, 'model_last.pth'))
    if world > 1:                     # captured graphs still hold the NCCL communicator: leave without tearing it down
        torch.cuda.synchronize()
        dist.barrier()
        logging.shutdown()
        os._exit(0)


if __name__ == '__main__':
    main()
