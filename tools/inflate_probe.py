import sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np
from nextpolish_b200 import engine as E
from tests.test_bgzf_inflate import synthetic_streams, zlib_inflate
L = E.lib()
for name, comp in sorted(synthetic_streams().items()):
    buf = np.frombuffer(comp, dtype=np.uint8)
    n, nb = C.c_int64(0), C.c_int32(0)
    assert L.np_bgzf_inflate(0, buf.ctypes.data, len(comp), None, 0, C.byref(n), C.byref(nb), None) == 0
    out = np.zeros(max(n.value, 1), np.uint8)
    assert L.np_bgzf_inflate(0, buf.ctypes.data, len(comp), out.ctypes.data, n.value, C.byref(n), C.byref(nb), None) == 0
    assert out[:n.value].tobytes() == zlib_inflate(comp), name
print("inflate probe ok")
