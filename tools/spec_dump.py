"""Write out the CUDA C++ text the engine generates for a BASELINE configuration's constraint set (spec_codegen.cpp) and
check it against the interpreted netlist on synthetic objects -- runs on the TEST-ONLY host emulation, no GPU needed.

    python tools/spec_dump.py [--config 2] [--objects 3000] [--out /tmp/spec/config2.cu]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--objects", type=int, default=3000)
ap.add_argument("--out", default="/tmp/spec/config2.cu")
a = ap.parse_args()
os.environ["GK_SPEC_DUMP"] = a.out
os.environ["GK_SPEC_CHECK"] = "1"
os.environ["GK_SPEC_TRACE"] = "1"

from gatekeeper_b200 import build, driver as D, workloads as Wl  # noqa: E402

build.build()
drv = D.Driver(lib_path=os.path.join(ROOT, "tests", "_hostemu", "libgk_hostemu.so"))
if a.config == 2:
    tmpls, cons = Wl.config2()
    mode = 0
elif a.config == 4:
    tmpls, cons = Wl.config4()
    mode = 1
else:
    tmpls, cons = Wl.config5()
    mode = 0
for kind, rego in tmpls:
    drv.add_template(kind, rego)
for c in cons:
    drv.AddConstraint(c)
for ns in Wl.synth_namespaces():
    drv.AddData("admission.k8s.gatekeeper.sh", ["cluster", "v1", "Namespace", ns["metadata"]["name"]], ns)
blob = Wl.synth_objects(0, a.objects, mode=mode)
r = drv.ReviewBlob(blob, with_results=False)
print("reviewed", a.objects, "objects; text in", a.out, os.path.getsize(a.out), "bytes")
