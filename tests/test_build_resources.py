"""Register budget of the hot kernels, checked at build time (no GPU needed).

The pooling kernels are compiled against a 128-register cap (4 CTAs of 128 threads per SM).  In
round 2 a cold feature compiled into the hot backward kernel (the deterministic variant's
bookkeeping behind a runtime flag) raised its spills from 8 to 28 bytes and cost the ATOMIC path
6 % on configs[1] without failing any test (profiles/r02_experiments.log, r03p).  This test reads
ptxas' resource usage and fails when a change pushes the hot kernels over their budget again."""
import os
import re
import subprocess
import tempfile

import pytest

from chainer_maskrcnn_b200 import _build

# kernel (substring of the mangled name) -> (max registers, max spill bytes)
BUDGET = {
    "20rpool_forward_kernel": (128, 0),
    "21rpool_backward_kernel": (128, 16),
    "32rpool_backward_det_window_kernel": (128, 16),
}


def _ptxas_report():
    if not _build.sources_present():
        pytest.skip("sources not present")
    try:
        nvcc = _build._nvcc()
    except RuntimeError:
        pytest.skip("nvcc not available")
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xptxas", "-v", "-cubin", "-o", os.path.join(tmp, "rpool.cubin"),
               os.path.join(_build.CSRC, "rpool_api.cu")]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 0, proc.stdout[-2000:]
    return proc.stdout


def test_hot_kernels_stay_inside_their_register_budget():
    out = _ptxas_report()
    # "Function properties for <name>\n  N bytes stack frame, A bytes spill stores, B bytes spill loads\n
    #  ptxas info : Used R registers, ..."
    blocks = re.findall(r"Function properties for (\S+)\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                        r"(\d+) bytes spill loads\s*\n[^\n]*Used (\d+) registers", out)
    seen = {}
    for name, _stack, st, ld, regs in blocks:
        for key in BUDGET:
            if key in name:
                seen[key] = (int(regs), max(int(st), int(ld)))
    assert set(seen) == set(BUDGET), (sorted(seen), "ptxas output format changed?")
    for key, (regs, spill) in seen.items():
        max_regs, max_spill = BUDGET[key]
        assert regs <= max_regs, (key, regs)
        assert spill <= max_spill, (key, "spills %d bytes > %d: something cold leaked into the hot kernel" %
                                    (spill, max_spill))
