"""Per-kernel signature of the SASS in libdiskann_b200.so: sha256 over the instruction stream (addresses and encodings
stripped).  Used to prove that a source change behind a template flag leaves the measured kernels untouched:

    python tools/sass_signature.py > /tmp/now.txt && diff profiles/r01_sass_signature.txt /tmp/now.txt
"""
import hashlib
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "pgvectorscale_b200", "libdiskann_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    ins = [re.sub(r"/\*[0-9a-f]{4,6}\*/", "", l).split("/*")[0].strip() for l in f.split("\n")
           if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
    rows.append((name, len(ins), hashlib.sha256("\n".join(ins).encode()).hexdigest()[:16]))
for name, n, h in sorted(rows):
    print(f"{h} {n:6d} {name}")
