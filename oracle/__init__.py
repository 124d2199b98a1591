"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's FPN multi-level RoIAlign path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or
as the timed CPU baseline.  The product (``chainer-maskrcnn_b200/``) never does.

Layout
------
roialign_oracle.c     C restatement of the op (chainer/NumPy path and caffe2 semantics)
ref_caffe2_wrapper.cpp  extern "C" door onto the reference's own C++ forward (-> _ref/)
Makefile              builds both shared objects
oracle.py             ctypes front-end + NumPy restatement of the level mapper and
                      of the heads' per-RoI level dispatch
reference_loader.py   imports the UNMODIFIED reference module from /root/reference
                      under a stub ``chainer`` (only where that tree exists)
"""
from .oracle import *  # noqa: F401,F403
