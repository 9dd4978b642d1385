"""TEST / BASELINE INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference into baseline/_ref.

The reference (unslothai/unsloth) is pure Python + Triton: there is nothing to compile into oracle/_ref.
What the GPU parity tests (tests/test_gpu_vs_reference.py) and the `gpu_reference` / `--impl gpu-reference`
legs of bench.py load is the reference package itself, installed offline from a /tmp copy of /root/reference
(the source tree is read-only and the build writes egg-info next to it):

    cp -r /root/reference /tmp/refcopy
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /tmp/refcopy

`--no-deps`: unsloth_zoo / peft / trl / bitsandbytes are not in the wheelhouse.
baseline/_ref is git-ignored, NOT gpurun-ignored: it travels to the GPU box with the snapshot.  Nothing under
/root/reference is read at run time on the GPU box.  `ensure()` is called by __graft_entry__.build(); it is a
no-op when baseline/_ref already holds the package or when /root/reference does not exist (the GPU box)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
REFERENCE = "/root/reference"


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "unsloth", "kernels", "fast_lora.py"))


def ensure(verbose: bool = False) -> bool:
    """Returns True when baseline/_ref holds the reference package after the call."""
    if installed():
        return True
    if not os.path.isdir(REFERENCE):
        return False
    tmp = tempfile.mkdtemp(prefix="refcopy_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, src, symlinks=True, ignore=shutil.ignore_patterns(".git"))
        os.makedirs(TARGET, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout[-2000:] + r.stderr[-2000:])
        return r.returncode == 0 and installed()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print("baseline/_ref installed:", ensure(verbose=True))
