"""Recipe for ``oracle/_ref/``: the reference's own Python files of the hot path, taken verbatim from
/root/reference at build time (TEST INFRASTRUCTURE; ``oracle/_ref/`` is git-ignored output, like a
compiled reference binary would be, and travels to the GPU box with the snapshot).

    python oracle/make_ref.py            # called by __graft_entry__.build() when /root/reference exists

What it is for (and the only places that may use it): ``tests/`` (the reference's GnnNet / gnnnet_copy /
DampNet classes driven on the GPU over this repo's ``methods/`` overlay and over the reference's own
``methods/gnn.py``), and the ``--impl reference`` / ``cpu_baseline`` legs of ``bench.py``
(``cpu_baseline.kind = "reference"``).  Nothing under ``meta-fine-tuning_b200/`` imports it.
Nothing is edited: files are copied byte for byte, a MANIFEST with their sha256 is written next to them.
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

FILES = ["backbone.py", "utils.py", "configs.py", "io_utils.py"]
DIRS = ["methods", "datasets"]          # *.py only


def main() -> int:
    if not os.path.isdir(REF):
        print("oracle/make_ref.py: /root/reference not present; keeping whatever oracle/_ref holds", file=sys.stderr)
        return 0
    todo = list(FILES)
    for d in DIRS:
        todo += [os.path.join(d, f) for f in sorted(os.listdir(os.path.join(REF, d))) if f.endswith(".py")]
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for rel in todo:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    print(f"oracle/_ref: {len(manifest)} files")
    return 0


if __name__ == "__main__":
    sys.exit(main())
