#!/usr/bin/env python
"""Build recipe for oracle/_ref — TEST / MEASUREMENT INFRASTRUCTURE, not product code.

The reference (shayneobrien/generative-models) is pure Python, so its "compiled" form is
byte code: this script byte-compiles the reference modules the benchmark's reference arm runs
FROM WHERE THEY LIE under /root/reference/src into oracle/_ref/*.pyc (sourceless imports; the
directory is git-ignored but travels to the GPU box, which has the same interpreter).  No
reference source text is copied into the repository.  `__graft_entry__.build()` calls this
when /root/reference exists; on the GPU box the prebuilt files are used as they are.

    python oracle/make_ref.py          # -> oracle/_ref/{ns_gan,utils,w_gp_gan,vae}.pyc + MANIFEST.json
"""
import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT = os.path.join(HERE, "_ref")
MODULES = ["utils", "ns_gan", "w_gp_gan", "vae"]


def build(force=False):
    """Returns the output directory, or None when the reference tree is not present."""
    if not os.path.isdir(REF_SRC):
        return OUT if os.path.exists(os.path.join(OUT, "ns_gan.pyc")) else None
    os.makedirs(OUT, exist_ok=True)
    manifest = {"python": sys.version.split()[0], "magic": __import__("importlib.util").util.MAGIC_NUMBER.hex(), "modules": {}}
    for m in MODULES:
        src = os.path.join(REF_SRC, m + ".py")
        dst = os.path.join(OUT, m + ".pyc")
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile="reference/src/%s.py" % m, doraise=True)
        manifest["modules"][m] = {"source": "reference/src/%s.py" % m,
                                  "sha256_of_source": hashlib.sha256(open(src, "rb").read()).hexdigest()}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return OUT


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print("oracle/_ref:", out)
