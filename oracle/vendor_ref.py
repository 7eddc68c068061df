"""Stage the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored).

TEST / BENCH INFRASTRUCTURE ONLY.  /root/reference exists in the build container but not on the GPU
box; baseline/_ref/ is git-ignored (no reference source ever enters this repository's history) but
NOT gpurun-ignored, so a byte-identical copy of the seven files the reference's render path imports
travels with the snapshot.  There it serves two purposes, both as the thing the product is compared
WITH, never as something the product calls:

  * `bench.py --impl reference`  times the reference's own PyTorch implementation on the host cores
    (and, as `reference_gpu`, on the B200) -- SURVEY.md 8(d);
  * tests/test_gpu_vs_reference.py runs the reference on the same GPU, same rays, same parameters,
    and compares every ray of whole batches.

    python oracle/vendor_ref.py          # run by __graft_entry__.build() when /root/reference exists

Files are copied verbatim (sha256 recorded in baseline/_ref/MANIFEST.json); nothing is patched.  The
unused third-party imports they carry are satisfied by the empty stub modules of oracle/ref_loader.py.
"""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LOCALRF_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "LICENSE",
    "localTensoRF/local_tensorfs.py",
    "localTensoRF/renderer.py",
    "localTensoRF/models/__init__.py",
    "localTensoRF/models/tensorBase.py",
    "localTensoRF/models/tensoRF.py",
    "localTensoRF/utils/__init__.py",
    "localTensoRF/utils/ray_utils.py",
    "localTensoRF/utils/utils.py",
]


def vendor(verbose=True):
    if not os.path.isfile(os.path.join(SRC, FILES[1])):
        if verbose:
            print(f"vendor_ref: no reference at {SRC}; keeping whatever is in {DST}")
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "facebookresearch/localrf @ 3905e39 (unmodified copies)", "sha256": manifest},
                  f, indent=1)
    if verbose:
        print(f"vendor_ref: {len(FILES)} files -> {DST}")
    return True


if __name__ == "__main__":
    vendor()
