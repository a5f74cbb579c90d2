"""Stage the UNMODIFIED reference modules that `bench.py --impl reference` times into baseline/_ref/ (git-ignored, but it
travels to the GPU box with the gpurun snapshot).  The reference is a script tree without setup.py / pyproject.toml, so
`pip install /root/reference` does not apply; what is staged is exactly the import closure of the training hot path:
code/models/*.py and code/utils/{criterions,lr_scheduler}.py, byte for byte (sha256 in MANIFEST.json).  Nothing here is
product code and nothing under baseline/_ref is ever committed.  Run by __graft_entry__.build() when /root/reference exists."""
import hashlib
import json
import os
import shutil

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "code")
FILES = ["models/blocks.py", "models/rfnet.py", "models/mmformer.py", "models/mask.py", "models/m2ftrans.py",
         "utils/criterions.py", "utils/lr_scheduler.py"]


def install():
    if not os.path.isdir(REF):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(src, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    print("installed" if install() else "no /root/reference here")
