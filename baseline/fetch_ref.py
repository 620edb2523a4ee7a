"""TEST / BENCH INFRASTRUCTURE ONLY -- makes the UNMODIFIED reference available to the GPU box.

    python baseline/fetch_ref.py            # /root/reference -> baseline/_ref  (byte-for-byte copies)

The reference (FedeMont/collision_handling_in_instantNGP) is five Python files plus three images; it is not a
pip-installable package (no setup.py / pyproject.toml), so "installing" it is copying those files.  The copy
lives in `baseline/_ref/`, which is git-ignored (reference sources never enter this repository's history) but NOT
gpurun-ignored, so it travels to the GPU box next to the built `.so`.  There it is what

  * `bench.py --impl reference` and the `cpu_baseline` / `gpu_eager_reference` legs execute (the reference's own
    `GeneralNeuralGaugeFields` + `Loss` + Adam, unmodified), and
  * `tests/test_reference_main_gpu.py` runs `main.py` from -- once on top of the drop-in `models` module and once on
    the reference's own CUDA-eager path (PSNR parity).

A manifest with the sha256 of every copied file is written next to the copy so that a run can state exactly which
reference it executed.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["main.py", "functions.py", "models.py", "utils.py", "params.py", "README.md"]


def fetch(src: str = "/root/reference", dest: str = DEST) -> bool:
    """Copies the reference files; returns False (and leaves `dest` alone) when `src` does not exist."""
    if not os.path.isdir(src):
        return False
    os.makedirs(os.path.join(dest, "images"), exist_ok=True)
    manifest = {}
    for name in FILES:
        shutil.copyfile(os.path.join(src, name), os.path.join(dest, name))
    for name in sorted(os.listdir(os.path.join(src, "images"))):
        shutil.copyfile(os.path.join(src, "images", name), os.path.join(dest, "images", name))
    for root, _, names in os.walk(dest):
        for name in sorted(names):
            if name == "MANIFEST.json" or name.endswith(".pyc"):
                continue
            path = os.path.join(root, name)
            manifest[os.path.relpath(path, dest)] = hashlib.sha256(open(path, "rb").read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1, sort_keys=True)
    return True


def available(dest: str = DEST) -> bool:
    return all(os.path.isfile(os.path.join(dest, n)) for n in FILES[:5]) and \
        os.path.isfile(os.path.join(dest, "images", "strawberry.jpeg"))


if __name__ == "__main__":
    ok = fetch(*(sys.argv[1:2]))
    print("copied the reference to", DEST if ok else "(nothing: source tree not found)")
