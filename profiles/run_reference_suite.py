"""Acceptance harness: the REFERENCE's own test-suite (tests/*.py of biotite-dev/springcraft 0.3.0) against this
repository's implementation.

Two steps, because /root/reference does not exist on the GPU box:
  1. in the build container:  python profiles/run_reference_suite.py stage
       copies /root/reference/tests (+ tests/data, pyproject.toml) into baseline/_ref/reference_suite/ -- a git-ignored
       directory that still travels with `gpurun` -- and writes an alias package `springcraft` next to it that
       re-exports springcraft_b200 (plus the test-only biotite stub of oracle/biotite_stub).  Nothing of the reference
       enters the repository history.
  2. on the GPU box:          python profiles/run_reference_suite.py run
       runs pytest on the staged suite with the alias first on PYTHONPATH and writes the per-case outcome to
       gpurun_out/reference_suite.json (copied to profiles/ by hand)."""
import json
import os
import shutil
import subprocess
import sys
from os.path import dirname, exists, join, realpath

ROOT = dirname(dirname(realpath(__file__)))
STAGE = join(ROOT, "baseline", "_ref", "reference_suite")

ALIAS = '''"""Alias: the reference's import name on top of springcraft_b200 (acceptance harness only)."""
import sys
import springcraft_b200 as _impl
from springcraft_b200 import *  # noqa: F401,F403
from springcraft_b200 import anm, forcefield, gnm, interaction, nma  # noqa: F401
for _name in ("anm", "forcefield", "gnm", "interaction", "nma"):
    sys.modules[__name__ + "." + _name] = getattr(_impl, _name)
__version__ = _impl.__reference_version__
'''


def stage():
    src = "/root/reference"
    if exists(STAGE):
        shutil.rmtree(STAGE)
    os.makedirs(STAGE)
    shutil.copytree(join(src, "tests"), join(STAGE, "tests"))
    shutil.copy(join(src, "pyproject.toml"), join(STAGE, "pyproject.toml"))
    os.makedirs(join(STAGE, "alias", "springcraft"))
    with open(join(STAGE, "alias", "springcraft", "__init__.py"), "w") as fh:
        fh.write(ALIAS)
    n = sum(len(f) for _, _, f in os.walk(join(STAGE, "tests")))
    print(f"staged {n} files under {STAGE}")


def run():
    if not exists(join(STAGE, "tests")):
        print("reference suite not staged (run `stage` in the build container first)")
        return 2
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([join(STAGE, "alias"), join(ROOT, "oracle", "biotite_stub"), ROOT,
                                         env.get("PYTHONPATH", "")])
    out_dir = join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    xml = join(out_dir, "reference_suite.xml")
    proc = subprocess.run([sys.executable, "-m", "pytest", join(STAGE, "tests"), "-q", "-p", "no:cacheprovider",
                           f"--junitxml={xml}", "-x" if "-x" in sys.argv else "-q"], env=env, cwd=STAGE,
                          capture_output=True, text=True)
    tail = proc.stdout[-3000:]
    print(tail)
    import xml.etree.ElementTree as ET
    cases = []
    for tc in ET.parse(xml).getroot().iter("testcase"):
        status = "passed"
        detail = ""
        for child in tc:
            if child.tag in ("failure", "error"):
                status = "failed"
                detail = (child.attrib.get("message") or "")[:300]
            elif child.tag == "skipped":
                status = "skipped"
        cases.append({"case": f"{tc.attrib.get('classname', '')}::{tc.attrib.get('name', '')}", "status": status,
                      "seconds": float(tc.attrib.get("time", 0.0)), **({"detail": detail} if detail else {})})
    summary = {s: sum(c["status"] == s for c in cases) for s in ("passed", "failed", "skipped")}
    failed = [c for c in cases if c["status"] == "failed"]
    report = {"suite": "biotite-dev/springcraft 0.3.0 tests/ (unmodified), import name aliased to springcraft_b200, "
                       "biotite = test-only stub", "summary": summary,
              "seconds": sum(c["seconds"] for c in cases), "failed": failed, "cases": cases}
    with open(join(out_dir, "reference_suite.json"), "w") as fh:
        json.dump(report, fh, indent=1)
    print(json.dumps({"summary": summary, "failed": [c["case"] for c in failed]}, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(stage() if sys.argv[1:2] == ["stage"] else run())
