"""Install the UNMODIFIED reference into baseline/_ref/ so that it travels to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/combat_oracle.py header): the product path never touches baseline/_ref.

The reference is a flat directory of Python scripts with no setup.py / pyproject.toml, so `pip install --target baseline/_ref
/root/reference` has nothing to build.  The equivalent install is a verbatim copy of its .py files, WHERE THEY LIE under
/root/reference, into baseline/_ref/ -- a directory that is git-ignored (no reference source enters the history) but not
gpurun-ignored (it ships with the snapshot like the built .so).  `bench.py --impl reference` and the `cpu_baseline` leg then run
the reference's own `train_generator.train()` on the box's host cores (oracle/ref_loader.py supplies the sys.modules stand-ins
for the absent kornia / vit_pytorch; nothing is patched).  Run by `__graft_entry__.build()` whenever /root/reference exists;
on the GPU box (no /root/reference) the prebuilt copy is used as it arrived.
"""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
# only what train_generator.py's import chain needs (train_generator.py:1-18)
WANTED_DIRS = ("classifier_models", "networks", "utils", os.path.join("defenses", "frequency_based"))
WANTED_TOP = ("config.py", "train_generator.py")


def installed() -> bool:
    return os.path.isfile(os.path.join(DST, "train_generator.py"))


def install(src: str = SRC, dst: str = DST) -> bool:
    """Copies the files; returns False (and leaves dst alone) when the reference tree is absent."""
    if not os.path.isfile(os.path.join(src, "train_generator.py")):
        return False
    os.makedirs(dst, exist_ok=True)
    for f in WANTED_TOP:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    for d in WANTED_DIRS:
        os.makedirs(os.path.join(dst, d), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(src, d))):
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(src, d, f), os.path.join(dst, d, f))
    with open(os.path.join(dst, "INSTALLED_FROM"), "w") as fh:
        fh.write("verbatim copy of the .py files of %s needed by train_generator.py (oracle/install_reference.py)\n" % src)
    return True


if __name__ == "__main__":
    print("installed" if install() else "reference tree absent: nothing installed", "->", DST)
