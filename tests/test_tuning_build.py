"""CPU: the A/B-measurement build of the RoI unit (-DRSDET_TUNING: csrc/roi_align_tuning.cuh and the environment switches the
profiles/ notes quote) still compiles for sm_100a.  Compile only; the object is thrown away."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.mark.skipif(not (os.path.exists(NVCC) or shutil.which("nvcc")), reason="nvcc not available")
def test_roi_unit_compiles_with_tuning_switches():
    src = os.path.join(ROOT, "rs_detection_b200", "csrc", "roi_align.cu")
    with tempfile.TemporaryDirectory() as d:
        cmd = [NVCC if os.path.exists(NVCC) else "nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17",
               "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-DRSDET_TUNING", "-DRSDET_BULK_ROWS=0",
               "-I", os.path.join(ROOT, "include"), "-c", src, "-o", os.path.join(d, "roi_tuning.o")]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-3000:]
