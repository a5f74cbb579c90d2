"""Command-line flags of the training entry point — same names, defaults and derived fields as the reference
`code/options.py:4-52` (argparse is the reference's only config mechanism), plus a few additions that do not exist
there: --synthetic (no dataset on disk), --dtype, --crop_size, --no_graph and --device_aug."""
import argparse
import os


def args_parser(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--model', default='mmformer', type=str, help='model name: rfnet | mmformer (both on the B200-native kernels); the reference default')
    parser.add_argument('-batch_size', '--batch_size', default=1, type=int, help='Batch size (per GPU)')
    parser.add_argument('--lr', default=2e-4, type=float, help='base learning rate')
    parser.add_argument('--weight_decay', default=1e-4, type=float)
    parser.add_argument('--num_epochs', default=300, type=int, help='training epochs')
    parser.add_argument('--temp', default=4.0, type=float, help='knowledge-distillation temperature')
    parser.add_argument('--region_fusion_start_epoch', default=0, type=int, help='warm-up epochs used in rfnet')
    # system
    parser.add_argument('--seed', default=1037, type=int, help='random seed')
    parser.add_argument('--gpu', type=str, default='0', help="GPU to use (ignored under torchrun: one rank per GPU).  The one default that differs from the reference ('3', its author's box): it must exist on a one-GPU machine")
    # options
    parser.add_argument('--mask_type', default='idt', type=str, help='training settings: pdt idt or idt_drop')
    parser.add_argument('--use_pretrain', action='store_true', default=False, help='whether use pretrained model')
    parser.add_argument('--use_passion', action='store_true', default=False, help='whether use passion')
    parser.add_argument('--use_valid', action='store_true', default=False, help='whether use validation')
    # paths
    parser.add_argument('--dataname', default='BraTS/BRATS2020', type=str)
    parser.add_argument('--datapath', default='BraTS/BRATS2020_Training_none_npy', type=str)
    parser.add_argument('--imbmrpath', default='BraTS/brats_split/Brats2020_imb_split_mr2468.csv', type=str, help='csv path')
    parser.add_argument('--savepath', default='outputs/idt_mr2468_mmformer_passion_bs1_epoch300_lr2e-4_temp4', type=str, help='output path')
    parser.add_argument('--resume', default=None, type=str, help='pretrained model path')
    parser.add_argument('--datarootPath', default=None, type=str, help='dataset root (default: ./datasets)')
    # additions (not in the reference)
    parser.add_argument('--synthetic', action='store_true', help='train on synthetic BraTS-shaped batches (no dataset needed)')
    parser.add_argument('--iters_per_epoch', default=0, type=int, help='with --synthetic: iterations per epoch (default 219 / global batch)')
    parser.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'], help='activation storage (f32 = check mode)')
    parser.add_argument('--crop_size', default=80, type=int, help='edge of the random training crop (reference: 80; mmformer needs a multiple of 16)')
    parser.add_argument('--no_graph', action='store_true', help='do not replay the step as a CUDA graph (N = 1)')
    parser.add_argument('--device_aug', action='store_true',
                        help='keep the preprocessed cases resident in HBM and run train_transforms (crop, rotation, intensity, flip) '
                             'and the label encoding as one kernel per batch (passion_b200/data.py) instead of on the host')
    parser.add_argument('--host_crop_only', action='store_true',
                        help='real data without --device_aug: accept the host loader, which applies the random crop ONLY '
                             '(no rotation / intensity change / flip); without this flag real-data training uses --device_aug')
    args = parser.parse_args(argv)
    if not args.synthetic and not args.device_aug and not args.host_crop_only:
        # the reference trains with the full train_transforms chain (options.py:50); only the device pipeline implements it
        args.device_aug = True

    root = args.datarootPath or os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'datasets'))
    args.datarootPath = root
    args.datasetPath = os.path.abspath(os.path.join(root, args.datapath))
    # kept for interface parity (options.py:50-51): --device_aug implements exactly this chain on the device
    # (passion_b200/data.py); the host loader of train.py without it does the random crop only
    args.train_transforms = 'Compose([RandCrop3D((80,80,80)), RandomRotion(10), RandomIntensityChange((0.1,0.1)), RandomFlip(0), NumpyType((np.float32, np.int64)),])'
    args.test_transforms = 'Compose([NumpyType((np.float32, np.int64)),])'
    return args
