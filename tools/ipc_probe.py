#!/usr/bin/env python
"""Development probe: what torch's storage._share_cuda_() hands out on this torch version,
next to cudaIpcGetMemHandle of the same allocation (for tools/tp_check.py's reference arm)."""
import ctypes

import torch

t = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
d = t.untyped_storage()._share_cuda_()
h = bytes(d[1])
print("share_cuda tuple lens:", [len(x) if isinstance(x, (bytes, bytearray)) else x for x in d])
print("handle len", len(h), "first bytes", h[:4].hex(), "last", h[-4:].hex())
rt = ctypes.CDLL("libcudart.so.12")
buf = ctypes.create_string_buffer(64)
base = t.data_ptr() - d[3]
rc = rt.cudaIpcGetMemHandle(buf, ctypes.c_void_p(base))
print("cudaIpcGetMemHandle rc", rc, "equal to torch[1:]", buf.raw == h[1:65], "equal to torch[:64]",
      buf.raw == h[:64])
