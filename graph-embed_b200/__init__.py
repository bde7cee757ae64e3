"""graph-embed_b200: B200-native ForceAtlas hot path of LLNL/graph-embed.

The directory name carries a hyphen, so import it through `__graft_entry__.load_package()`
(module name `graph_embed_b200`).  Contents: csrc/ (CUDA kernels + C ABI), host/ (C++ drop-in
headers mirroring the reference interface), capi.py (ctypes binding of the C ABI used by tests and
bench.py), graphs.py (synthetic input generators), build.py (nvcc driver).
"""
