"""Recipe for ``oracle/_ref``: the reference's own Torch update path, byte-compiled so that it can travel to the GPU box.

TEST INFRASTRUCTURE ONLY (checker + the ``--impl reference`` / ``cpu_baseline`` legs of bench.py).

    python -m oracle.build_ref          # needs /root/reference (the build container)

The reference is pure Python, so "compiling it from the sources where they lie" means ``py_compile``: every module
``import src.torchrl.reppo`` pulls in is compiled from ``/root/reference`` into a SOURCELESS code object under
``oracle/_ref/`` (git-ignored, not gpurun-ignored -- it ships like our own ``.so``; the files are named ``*.pyc.bin``
because the gpurun snapshot skips ``*.pyc``, and ``oracle/ref_torch_shim.py`` installs a small importer for them).  No reference source is copied, and
nothing under ``reppo_b200/`` ever imports it.  ``oracle/ref_torch_shim.py`` imports the modules from
``/root/reference`` when that exists and from ``oracle/_ref`` otherwise; ``oracle/_stubs`` stands in for the four
packages the reference imports that are not installed here (tensordict, hydra, omegaconf, torchinfo).
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"
OUT = os.path.join(HERE, "_ref")
# what `import src.torchrl.reppo` loads (checked with sys.modules after the import)
MODULES = [
    "src/__init__.py",
    "src/networks/__init__.py",
    "src/networks/torch_models.py",      # FCNN / Critic / Actor, torch_models.py:82-335
    "src/torchrl/__init__.py",
    "src/torchrl/envs.py",               # imported by reppo.py at module level; simulators are imported lazily inside
    "src/torchrl/reppo.py",              # make_postprocess_fn / make_critic_update_fn / make_actor_update_fn, :173-360
    "src/torchrl/reppo_util.py",         # hl_gauss, EmpiricalNormalization, :394-471,756-777
]


def build(force: bool = False) -> str | None:
    """Returns the output directory, or None when the reference is not present (GPU box: the prebuilt files are used)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "torchrl")):
        return None
    stamp = os.path.join(OUT, "STAMP")
    tag = f"{sys.version_info[:3]}"
    if not force and os.path.exists(stamp) and open(stamp).read() == tag:
        return OUT
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    for rel in MODULES:
        dst = os.path.join(OUT, rel[:-3] + ".pyc.bin")  # a .pyc (16-byte header + marshalled code object) under another suffix
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(os.path.join(REFERENCE_ROOT, rel), cfile=dst, dfile=rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(stamp, "w") as f:
        f.write(tag)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
