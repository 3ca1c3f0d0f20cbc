"""ORACLE / test infrastructure: stage the UNMODIFIED reference where tests and the benchmark's CPU arm can run it.

The reference is pure Python, so "building" it is copying its files out of /root/reference (which does not exist on the GPU
box) - by this committed recipe, at build() time, into git-ignored directories that travel with the built artefacts:

  oracle/_ref/reference/{PPO,AsyncTools}/*.py            the reference's hot path, verbatim - imported behind a stub `gymnasium`
                                                         whose envs are oracle/envs.py by oracle/ref_runner.py (bench.py
                                                         --impl reference, kind "reference")
  oracle/_ref/dropin/unittests/*.py                      the reference's own unittests, verbatim; each does
                                                         sys.path.insert(0, <its directory>/..) - an empty directory here - and then
                                                         imports `PPO` / `AsyncTools`, which PYTHONPATH resolves to the drop-in
                                                         packages of the B200 build
                                                         (tests/test_gpu_api.py::test_reference_unittests_run_unmodified)

Nothing here is product code and nothing staged is ever committed (.gitignore)."""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
REF_STAGE = os.path.join(HERE, "_ref", "reference")
UNITTEST_STAGE = os.path.join(HERE, "_ref", "dropin", "unittests")
PLAN = (("PPO", os.path.join(REF_STAGE, "PPO")), ("AsyncTools", os.path.join(REF_STAGE, "AsyncTools")), ("unittests", UNITTEST_STAGE))


def stage(ref: str = REF) -> bool:
    """Copy the files if `ref` exists; returns whether the staged copies are present afterwards."""
    if os.path.isdir(ref):
        for sub, dst in PLAN:
            os.makedirs(dst, exist_ok=True)
            for f in sorted(os.listdir(os.path.join(ref, sub))):
                if f.endswith(".py"):
                    shutil.copyfile(os.path.join(ref, sub, f), os.path.join(dst, f))
    return staged()


def staged() -> bool:
    return all(os.path.isdir(dst) and any(f.endswith(".py") for f in os.listdir(dst)) for _, dst in PLAN)


def verify(ref: str = REF) -> list[str]:
    """Names of staged files that differ from the reference's (empty list = byte-identical).  Needs `ref`."""
    bad = []
    for sub, dst in PLAN:
        for f in sorted(os.listdir(os.path.join(ref, sub))):
            if f.endswith(".py"):
                a, b = os.path.join(ref, sub, f), os.path.join(dst, f)
                if not os.path.exists(b) or open(a, "rb").read() != open(b, "rb").read():
                    bad.append(os.path.join(sub, f))
    return bad


if __name__ == "__main__":
    print("staged" if stage() else "reference not available and nothing staged")
