#!/usr/bin/env python3
"""Install the UNMODIFIED reference into the git-ignored ``baseline/_ref`` (it travels to the GPU box with gpurun).

Used (a) by ``bench.py --impl reference`` to time the reference's own CPU implementation of the step
(``cpu_baseline.kind = "reference"``) and (b) by ``tests/test_gpu_reference_callers.py`` to drive the reference's
own callers (``Greedy``, ``MLP``, ``SimpleGaussianES.get_fitness``, its unit tests) through the CUDA drop-in.
Nothing under ``therldaisyworld_b200/`` imports it.

Recipe: ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>``
(``--no-deps``: the pins numpy==1.24.2 / matplotlib / mpi4py are not installable offline; the copy is needed because the
build writes into the source tree and /root/reference is read-only).  The reference's ``setup.py`` lists only
``packages=["daisy"]``, so the wheel lacks the sub-packages ``daisy.nn`` (imported by ``daisy_world_rl``),
``daisy.agents`` and ``daisy.evo``: they are completed from the same source tree, together with the reference's
``tests/`` and the stored trained network ``results/cmaes_exp_002/*gen127.json`` the MLP tests load.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")


def install(ref="/root/reference", force=False):
    """Returns the install path, or None when the reference source tree is absent (GPU box: uses what travelled)."""
    marker = os.path.join(DEST, "daisy", "nn", "functional.py")
    if os.path.exists(marker) and not force:
        return DEST
    if not os.path.isdir(os.path.join(ref, "daisy")):
        return None
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(ref, src)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DEST, src]
        subprocess.run(cmd, check=True)
    for sub in ("nn", "agents", "evo"):
        dst = os.path.join(DEST, "daisy", sub)
        if not os.path.isdir(dst):
            shutil.copytree(os.path.join(ref, "daisy", sub), dst)
    shutil.copytree(os.path.join(ref, "tests"), os.path.join(DEST, "ref_tests"), dirs_exist_ok=True)
    os.makedirs(os.path.join(DEST, "results", "cmaes_exp_002"), exist_ok=True)
    for f in glob.glob(os.path.join(ref, "results", "cmaes_exp_002", "*gen127.json")):
        shutil.copy(f, os.path.join(DEST, "results", "cmaes_exp_002"))
    return DEST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
