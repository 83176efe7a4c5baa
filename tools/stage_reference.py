"""Stages the reference's own hot-path modules for the `--impl reference` arm of bench.py on the GPU box.

/root/reference exists only in the build container, so the UNMODIFIED files of the path (mdqm9/thermo/**, adw/thermo/**,
mdqm9/analysis/utils/*.py - pure Python) are copied, byte for byte, into baseline/_ref/ (git-ignored, NOT gpurun-ignored:
it travels with the snapshot like a built .so).  The reference has no setup.py / pyproject.toml, so `pip install --target
baseline/_ref /root/reference` is not possible (recorded in DESIGN.md); its un-vendored dependencies
(torch_geometric / torch_scatter / torchdiffeq) are the ~150-line stand-ins under oracle/stubs.

    python tools/stage_reference.py            # called by __graft_entry__.build() when /root/reference is present
"""
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("TI_REFERENCE_SRC", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")
TREES = ["mdqm9/thermo", "adw/thermo", "mdqm9/analysis/utils"]


def stage() -> bool:
    if not os.path.isdir(os.path.join(SRC, "mdqm9", "thermo")):
        return False
    manifest = {}
    for tree in TREES:
        for root, _dirs, files in os.walk(os.path.join(SRC, tree)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                src = os.path.join(root, f)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                with open(src, "rb") as fh:
                    manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    for pkg in ("mdqm9", "mdqm9/analysis"):        # namespace packages in the reference; explicit here for importlib
        os.makedirs(os.path.join(DST, pkg), exist_ok=True)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump(dict(source=SRC, files=manifest), fh, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else f"reference not found under {SRC}", DST)
    sys.exit(0 if ok else 1)
