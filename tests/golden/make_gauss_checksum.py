"""Writes tests/golden/gauss_table.sha256: SHA-256 of the 10100 doubles of the REFERENCE's Gauss-Legendre table
(/root/reference/cpp/qugar/quadrature_lib.cpp, Gauss::get_quad_data, n = 1..100 on [0,1], (x,w) pairs) as
little-endian IEEE doubles.  Run in the build container (the reference tree is not on the GPU box)."""
import hashlib
import os
import re
import struct

src = open("/root/reference/cpp/qugar/quadrature_lib.cpp").read()
i = src.index("Gauss::get_quad_data")
j = src.index("};", i)
nums = [float(x) for x in re.findall(r"[-+]?\d\.\d+e[-+]\d+", src[i:j])]
assert len(nums) == 10100
digest = hashlib.sha256(struct.pack("<%dd" % len(nums), *nums)).hexdigest()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gauss_table.sha256")
open(out, "w").write(digest + "  reference quadrature_lib.cpp Gauss table, 10100 doubles\n")
print(digest)
