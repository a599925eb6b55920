"""Install the UNMODIFIED reference into baseline/_ref/ (git-ignored, travels to the GPU box with the gpurun snapshot).

The reference is two plain Python files without packaging metadata, so the contract's command
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
fails ("Neither 'setup.py' nor 'pyproject.toml' found").  /root/reference is read-only, so the files are copied to a
scratch directory under /tmp, a three-line setup.py (py_modules=[...]) is generated NEXT TO them, and the same pip
command installs that copy (--no-deps: `diffusers` & co. are absent offline and are shimmed by oracle/ref_shim.py).
The installed module files are byte-identical to the reference's (checked below).  Nothing here is product code; only
bench.py's reference legs and tests import the result, through oracle/ref_shim.py.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
MODULES = ["elastic_diffusion", "elastic_diffusion_w_controlnet"]


def install(src="/root/reference", force=False):
    """Returns a one-line outcome string."""
    have = all(os.path.isfile(os.path.join(TARGET, m + ".py")) for m in MODULES)
    if have and not force:
        return f"present: {TARGET}"
    if not os.path.isdir(src):
        return f"skipped: {src} does not exist here (the GPU box uses the files installed in the build container)"
    os.makedirs(TARGET, exist_ok=True)
    pip = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
           "--upgrade", "--target", TARGET]
    r = subprocess.run(pip + [src], capture_output=True, text=True)
    if r.returncode == 0:
        return "installed straight from " + src
    tmp = tempfile.mkdtemp(prefix="ed_ref_src_")
    try:
        for m in MODULES:
            shutil.copy(os.path.join(src, m + ".py"), tmp)
        with open(os.path.join(tmp, "setup.py"), "w") as f:
            f.write("from setuptools import setup\n"
                    f"setup(name='elasticdiffusion-reference', version='0', py_modules={MODULES!r})\n")
        r = subprocess.run(pip + ["--no-deps", tmp], capture_output=True, text=True)
        if r.returncode != 0:
            return "failed: " + (r.stderr.strip().splitlines() or ["pip error"])[-1]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for m in MODULES:
        if not filecmp.cmp(os.path.join(src, m + ".py"), os.path.join(TARGET, m + ".py"), shallow=False):
            return f"failed: installed {m}.py differs from the reference"
    return f"installed from a /tmp copy with a generated setup.py (pip --target {os.path.relpath(TARGET, ROOT)}); files identical"


if __name__ == "__main__":
    print("[install_reference]", install(force="--force" in sys.argv))
