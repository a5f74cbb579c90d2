"""Backbones of the reference's code/models on the passion_b200 kernels."""


def build_model(name, num_cls=4, crop=80):
    """`--model` of options.py -> module instance.  mmFormer's token grid is crop/16 per axis: the reference hard-codes
    patch_size = 5 for its 80^3 crops (mmformer.py:21); here it follows the crop size."""
    if name == "rfnet":
        from . import rfnet
        return rfnet.Model(num_cls=num_cls)
    if name == "mmformer":
        from . import mmformer
        if crop % 16:
            raise ValueError("mmformer needs a crop size that is a multiple of 16")
        old = mmformer.patch_size
        mmformer.patch_size = crop // 16
        try:
            return mmformer.Model(num_cls=num_cls)
        finally:
            mmformer.patch_size = old
    raise ValueError(f"model {name!r} is not implemented on the B200-native path (rfnet, mmformer)")
