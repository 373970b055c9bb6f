"""TEST / BENCH INFRASTRUCTURE ONLY.  Stages the UNMODIFIED reference modules of the sampling path into oracle/_ref/ so that they
travel to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored) and `bench.py --impl reference` can time the reference's own
CPU implementation there.  Nothing is edited: the files go byte-for-byte into ONE archive, oracle/_ref/reference_model.zip (+ a
manifest with their sha256); oracle/ref_loader.py unpacks it into a temporary directory at import time.  The reference is pure
Python; its two absent third-party imports are served by oracle/shims.py exactly as for the golden fixtures.

    python -m oracle.stage_ref          (run by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DIP_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
ARCHIVE = "reference_model.zip"
FILES = ["model/diffusion.py", "model/decoder.py", "model/cmsp.py", "model/encoder.py", "model/modules.py", "model/utils.py",
         "model/base_model.py", "gnn_model/gnn_model.py", "gnn_model/basic_module.py"]


def stage():
    if not os.path.isfile(os.path.join(SRC, "model", "diffusion.py")):
        return None
    manifest = {}
    os.makedirs(DST, exist_ok=True)
    with zipfile.ZipFile(os.path.join(DST, ARCHIVE), "w", zipfile.ZIP_DEFLATED) as z:
        for rel in FILES:
            s = os.path.join(SRC, rel)
            if not os.path.isfile(s):
                continue
            data = open(s, "rb").read()
            z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), data)
            manifest[rel] = hashlib.sha256(data).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    return os.path.join(DST, ARCHIVE)


if __name__ == "__main__":
    print(stage())
    sys.exit(0)
