"""Overlay of the reference's `jdet` package for the rotated-box hot path.

Module paths, function / class names, argument order and error behaviour mirror
`/root/reference/python/jdet/...` so that call sites (configs, heads, runner) need no change; the
bodies call librsdet.so.  Tensors are torch CUDA tensors here (Jittor is not installable in this
image); `rs_detection_b200/jittor_adapter.py` carries the `jt.code` binding for a Jittor host.
Host inputs (numpy / CPU tensors) are accepted: they are copied to the device, processed there and
the result is copied back -- the computation itself never runs on the CPU.
"""
