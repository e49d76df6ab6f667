#!/usr/bin/env python
"""Run the two example executables on the shipped VGP (GPU needed) and show their results."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import plugin_binding as pb
pb.lib()  # builds build/*
work = tempfile.mkdtemp()
xml = pb.write_reference_xml(os.path.join(work, "ocp.xml"))
for exe in ("etol_ecuda_example1", "etol_ecuda_example2", "etol_ecuda_example3", "etol_ecuda_example4"):
    r = subprocess.run([os.path.join(ROOT, "build", exe), xml], capture_output=True, text=True, cwd=work, timeout=600)
    keep = [l for l in r.stdout.splitlines() if any(w in l for w in ("Score", "saved", "failed", "iter  ", "device model", "mesh:", "path rows"))]
    print(exe, "rc", r.returncode)
    print("\n".join(keep[-6:]))
    print(r.stderr[-400:])
for f in sorted(os.listdir(work)):
    if f.endswith(".csv"):
        print(f, open(os.path.join(work, f)).read().splitlines()[:3])
