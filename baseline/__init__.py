"""Baselines that bench.py times beside the product (never imported by the package).

refgpu -- the reference's CuPy kernels (roi_align_2d.py:100-144, :196-279) restated
          in CUDA and dispatched per RoI like the reference's FPN heads: the same-box
          GPU baseline ("gpu_baseline" in bench.py) and a cross-check in tests/.
_ref/  -- reserved for an installed copy of the reference (git-ignored; the reference
          has no setup.py/pyproject and its runtime is absent).
"""
