"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that vendors the reference's hot-path sources, UNMODIFIED, into
`oracle/_ref/` (git-ignored: no reference source enters this repository's history; NOT gpurun-ignored: the folder
travels to the GPU box, where /root/reference does not exist).

    python oracle/build_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python: "building" it is copying the three files of the path --
mp_rgcn_layer.py, model.py, main.py -- byte for byte, plus a manifest with their SHA-256.  They only import behind the
dependency stand-ins of oracle/ref_shims.py (torch_geometric / mpi4py / plotting libraries are not installed here or on
the GPU box).  Users: `bench.py --impl reference` and bench.py's `cpu_baseline` leg (kind "reference"), nothing else.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("mp_rgcn_layer.py", "model.py", "main.py")


def build(reference_root="/root/reference"):
    if not os.path.isfile(os.path.join(reference_root, FILES[0])):
        return None                                    # GPU box: use what travelled with the snapshot
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(reference_root, f), os.path.join(DEST, f)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": reference_root, "sha256": manifest}, fh, indent=1)
    return DEST


def available():
    return all(os.path.isfile(os.path.join(DEST, f)) for f in FILES)


if __name__ == "__main__":
    print(build() or ("no reference at /root/reference; oracle/_ref %s" % ("present" if available() else "absent")))
