"""TEST / BENCH INFRASTRUCTURE — makes the UNMODIFIED reference importable on the GPU box.

    python -m oracle.vendor_reference          (run in the build container; __graft_entry__.build() calls it)

`/root/reference` does not exist on the GPU box, and the reference is pure Python, so the honest CPU / GPU-eager
baselines of bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`) need the package to travel.  The
sanctioned route, `pip install --target <dir> /root/reference`, fails on the reference's own packaging metadata
(pyproject.toml declares `[project]` without the fields setup.cfg supplies; setuptools >= 61 rejects it:
"AttributeError: 'NoneType' object has no attribute 'get'" in _long_description), so this script does what that
install would have done for a pure-Python package: it copies the package tree `src/tacorl/**/*.py` into
`oracle/_ref/tacorl/` — git-ignored (never part of the repo's history), not gpurun-ignored (it travels with the
snapshot like the built .so).  Its third-party dependencies that are absent from the image (hydra, omegaconf,
pytorch_lightning, gym, ...) are satisfied by the import stubs in oracle/stubs/ (oracle/ref_loader.py).

Nothing under tacorl_b200/ imports the vendored tree; bench.py's reference arm and oracle/make_golden.py do."""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/tacorl"
DST = os.path.join(HERE, "_ref")


def try_pip():
    """The contract's install command, on a scratch copy (the source tree is read-only).  Returns (ok, last line)."""
    tmp = "/tmp/_tacorl_ref_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree("/root/reference", tmp, ignore=shutil.ignore_patterns(".git"))
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", DST, tmp]
    p = subprocess.run(cmd, capture_output=True, text=True, cwd="/tmp")
    tail = [l for l in (p.stdout + p.stderr).splitlines() if l.strip()]
    return p.returncode == 0 and os.path.isdir(os.path.join(DST, "tacorl")), (tail[-1] if tail else "")


def vendor(force=False, use_pip=False):
    if not os.path.isdir(SRC):
        return None                                   # GPU box / fresh clone: use what travelled, or the oracle port
    stamp = os.path.join(DST, "VENDORED.json")
    if os.path.exists(stamp) and not force:
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    how, note = "copied src/tacorl/**/*.py", ""
    if use_pip:
        ok, note = try_pip()
        if ok:
            how = "pip install --target"
        else:
            shutil.rmtree(DST, ignore_errors=True)
    files = {}
    if not os.path.isdir(os.path.join(DST, "tacorl")):
        for dp, _, fs in os.walk(SRC):
            for f in fs:
                if f.endswith(".py"):
                    src = os.path.join(dp, f)
                    rel = os.path.relpath(src, os.path.dirname(SRC))
                    dst = os.path.join(DST, rel)
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copyfile(src, dst)
                    files[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()[:16]
    json.dump({"source": SRC, "how": how, "pip_note": note, "files": files}, open(stamp, "w"), indent=1)
    return DST


if __name__ == "__main__":
    print(vendor(force=True, use_pip="--pip" in sys.argv))
